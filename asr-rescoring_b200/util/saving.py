"""Output plumbing with the reference's exact formatting: json_saving (util/saving.py:14-16: indent=4,
ensure_ascii=False) and model_saving (util/saving.py:7-11: torch.save of the bare state_dict to
<file_path>/checkpoint_<n>.pth — the file pll_bert_scoring loads back, MLM_PLL/main.py:185-187)."""
import json
import os
from typing import Dict


def model_saving(file_path: str, model_dict, checkpoint_num: int):
    import torch
    torch.save(model_dict, os.path.join(file_path, 'checkpoint_{}.pth'.format(checkpoint_num)))


def json_saving(file_path: str, json_data: Dict):
    with open(file_path, "w", encoding="utf8") as f:
        json.dump(json_data, f, ensure_ascii=False, indent=4)
