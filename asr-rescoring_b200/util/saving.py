"""json_saving with the reference's exact formatting (util/saving.py:14-16: indent=4,
ensure_ascii=False).  model_saving (checkpointing during fine-tuning) is out of scope."""
import json
from typing import Dict


def json_saving(file_path: str, json_data: Dict):
    with open(file_path, "w", encoding="utf8") as f:
        json.dump(json_data, f, ensure_ascii=False, indent=4)
