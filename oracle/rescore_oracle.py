"""CPU restatement (numpy / pure Python) of the reference combiner, rescore.py.

TEST INFRASTRUCTURE ONLY — see oracle/pll_oracle.py for the rule.

Restates (file:line relative to the reference tree):
  dict_to_list ............ rescore.py:13-23
  find_best_weight ........ rescore.py:25-45
  rescore ................. rescore.py:47-53   (variant "B"; "A"/"C" are the
                            logged formulas)
  get_highest_score_hyp ... rescore.py:55-58
  jiwer.cer ............... third-party `jiwer` (unpinned, not installed): corpus
                            CER = sum Levenshtein(ref_i, hyp_i) / sum len(ref_i)
                            over code points of the stripped strings; raises on
                            an empty reference.  Pinned by Nbest_Align/cer.json
                            (7 176 triples) and the integer identities of the
                            logged CERs (tests/test_oracle.py).
"""
from __future__ import annotations

import sys
from types import SimpleNamespace
from typing import Dict, List

import numpy as np


def levenshtein(a: str, b: str) -> int:
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j - 1] + (ca != cb), prev[j] + 1, cur[j - 1] + 1))
        prev = cur
    return prev[-1]


def cer(ref, hyp) -> float:
    """jiwer.cer(reference, hypothesis) for str or list[str] arguments."""
    if isinstance(ref, str):
        ref = [ref]
    if isinstance(hyp, str):
        hyp = [hyp]
    if len(ref) != len(hyp):
        raise ValueError("reference and hypothesis lists differ in length")
    edits = 0
    total = 0
    for r, h in zip(ref, hyp):
        r, h = r.strip(), h.strip()
        if len(r) == 0:
            raise ValueError("one or more references are empty strings")
        edits += levenshtein(r, h)
        total += len(r)
    return float(edits) / float(total)


def dict_to_list(d):
    out = []
    for _, hyps in d.items():
        if isinstance(hyps, Dict):
            out.append([hyp for _, hyp in hyps.items()])
        else:
            out.append(hyps)
    return out


def rescore(weight, hyps_len, am, lm, config, variant: str = "B"):
    am = np.array(am)[:, :config.n_best]
    lm = np.array(lm)
    hyps_len = np.array(hyps_len)
    if variant == "B":      # rescore.py:51 (current source)
        return (1 - weight) * (am) / hyps_len + weight * (lm) / hyps_len
    if variant == "A":      # rescore_result/MLM_PLL/rescore.log:28
        return (1 - weight) * (am) + weight * (lm)
    if variant == "C":      # rescore_result/RMBR/BertScore/rescore_mbr_normalize.log:29
        return (1 - weight) * (am) / hyps_len + weight * (lm)
    raise ValueError(variant)


def get_highest_score_hyp(final_score, hyps):
    idx = np.argmax(final_score, axis=-1)
    return [ht[i] for ht, i in zip(hyps, idx)]


def hyps_len_of(hyps: List[List[str]], n_best: int):
    return [[len(h) for h in utt[:n_best]] for utt in hyps]


def find_best_weight(am, lm, hyps, ref, config, variant: str = "B", weights=None):
    best_cer = sys.float_info.max
    best_weight = None
    hyps_len = hyps_len_of(hyps, config.n_best)
    if weights is None:
        weights = np.arange(0.0, 1.01, 0.01)
    with np.errstate(all="ignore"):
        for w in weights:
            final = rescore(w, hyps_len, am, lm, config, variant)
            pred = get_highest_score_hyp(final, hyps)
            err = cer(ref, pred)
            if err < best_cer:
                best_cer, best_weight = err, w
    return best_weight, best_cer


def config(n_best: int) -> SimpleNamespace:
    return SimpleNamespace(n_best=n_best)
