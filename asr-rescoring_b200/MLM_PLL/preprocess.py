"""Drop-in for the reference's ``MLM_PLL/preprocess.py`` (the for_training and for_scoring jobs).

The reference materialises every masked copy as a JSON row (O(sum L^2) integers,
MLM_PLL/preprocess.py:9-30); here the expansion happens on the GPU (stage 1 of
libpllb200), so preprocessing only tokenises each hypothesis once and writes a compact
packed file.  ``do_job`` is kept with the reference's signature and row schema for callers
that still want rows (it is what the parity tests compare the device expansion against).
The two ``for_training`` jobs (preprocess.py:36-44, over ref_text.json) do write the reference's row
list: the fine-tuning loss needs every masked copy as a batch row (MLM_PLL/main.py:117-161).
"""
from __future__ import annotations

import json
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_PARENT = os.path.dirname(os.path.dirname(_HERE))
if _PKG_PARENT not in sys.path:
    sys.path.insert(0, _PKG_PARENT)

from asr_rescoring_b200.synth import CLS_ID, MASK_ID, SEP_ID  # noqa: E402
from asr_rescoring_b200.tokenizer import BertCharTokenizer, SyntheticCharTokenizer  # noqa: E402
from asr_rescoring_b200.util.saving import json_saving  # noqa: E402

bert_tokenizer = None


def _special(tokenizer, token, default):
    vocab = getattr(tokenizer, "vocab", None)
    return vocab.get(token, default) if vocab else default


def do_job(sentence, utt_id, hyp_id, task_type, output_json, tokenizer=None):
    """Row schema of MLM_PLL/preprocess.py:9-30 (one row per masked position)."""
    tk = tokenizer or bert_tokenizer
    ids = tk.encode(sentence)
    cls_id, sep_id, mask_id = _special(tk, "[CLS]", CLS_ID), _special(tk, "[SEP]", SEP_ID), _special(tk, "[MASK]", MASK_ID)
    for mask_pos in range(len(ids)):
        output_json.append({
            "utt_id": utt_id,
            "hyp_id": hyp_id,
            "input_ids": [cls_id] + ids[:mask_pos] + [mask_id] + ids[mask_pos + 1:] + [sep_id],
            "attention_masks": [1] * (len(ids) + 2),
            "mask_pos": mask_pos + 1,
            "labels": [cls_id] + ids + [sep_id],
        })
    return output_json


def pack_hyps_text(hyps_text: dict, tokenizer) -> dict:
    """{utt: {hyp: str}} -> packed JSON consumed by main.py (format pllb-packed-v1)."""
    utt, hyp, tokens, offsets = [], [], [], [0]
    for utt_id, hyps in hyps_text.items():
        for hyp_id, sentence in hyps.items():
            utt.append(utt_id)
            hyp.append(hyp_id)
            tokens.extend(tokenizer.encode(sentence))
            offsets.append(len(tokens))
    kind = "synthetic" if isinstance(tokenizer, SyntheticCharTokenizer) else "bert-vocab"
    return {"format": "pllb-packed-v1", "tokenizer": kind, "utt_id": utt, "hyp_id": hyp, "tokens": tokens,
            "offsets": offsets}


if __name__ == "__main__":
    # The reference fetches the bert-base-chinese vocabulary from the hub (preprocess.py:34); offline
    # it has to be given.  The synthetic char->id map is an explicit opt-in for random-init
    # experiments (PLLB_SYNTHETIC_TOKENIZER=1) and is recorded in the packed file, so main.py can
    # refuse to score a real checkpoint with it.
    vocab = os.environ.get("PLLB_VOCAB")
    if vocab:
        bert_tokenizer = BertCharTokenizer(vocab)
    elif os.environ.get("PLLB_SYNTHETIC_TOKENIZER") == "1":
        bert_tokenizer = SyntheticCharTokenizer()
    else:
        raise SystemExit("preprocess.py: set PLLB_VOCAB=<path to bert-base-chinese vocab.txt> "
                         "(or PLLB_SYNTHETIC_TOKENIZER=1 for random-init experiments)")
    jobs = [
        {"task": "for_training", "in": "../espnet_data/alfred/train/ref_text.json", "out": "preprocessed_data/for_training/train.json"},
        {"task": "for_training", "in": "../espnet_data/alfred/dev/ref_text.json", "out": "preprocessed_data/for_training/dev.json"},
        {"task": "for_scoring", "in": "../espnet_data/alfred/train/hyps_text.json", "out": "preprocessed_data/for_scoring/train.json"},
        {"task": "for_scoring", "in": "../espnet_data/alfred/dev/hyps_text.json", "out": "preprocessed_data/for_scoring/dev.json"},
        {"task": "for_scoring", "in": "../espnet_data/alfred/test/hyps_text.json", "out": "preprocessed_data/for_scoring/test.json"},
    ]
    for job in jobs:
        json_data = json.load(open(job["in"], "r", encoding="utf-8"))
        os.makedirs(os.path.dirname(job["out"]), exist_ok=True)
        if job["task"] == "for_training":            # preprocess.py:58-60: rows of the reference sentences, hyp_id None
            output_json = []
            for utt_id, sentence in json_data.items():
                output_json = do_job(sentence, utt_id, None, job["task"], output_json)
            json_saving(job["out"], output_json)
        else:
            json_saving(job["out"], pack_hyps_text(json_data, bert_tokenizer))
