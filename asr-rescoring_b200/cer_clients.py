"""Other reference call sites of ``jiwer.cer`` served by the warp-per-pair Levenshtein kernel
(SURVEY.md §8f rank 3):

* per-hypothesis CER of the upstream data prep — ``espnet_data/preprocess/main.py:59-60``
  (``hyps_cer.json``: {utt: {hyp_k: cer(ref, hyp_k)}});
* the CER utility of minimum-Bayes-risk decoding — ``RMBR/utility_functions.py:24-33``
  (``CerScoreFunction``: 1 - cer(ref, cand) for n*(n-1) pairs per utterance) and
  ``RMBR/mbr.py:5-27`` (``mbr_decode``).

Every pair of a call goes through ONE kernel launch; jiwer's per-pair semantics are kept:
strings stripped, characters = code points, CER = distance / len(reference), an empty
reference raises.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from . import engine


def pair_cer(refs: Sequence[str], hyps: Sequence[str]) -> np.ndarray:
    """float64 [n]: jiwer.cer(refs[i], hyps[i]) for every i."""
    if len(refs) != len(hyps):
        raise ValueError("reference and hypothesis lists differ in length")
    lens = np.array([len(r.strip()) for r in refs], np.int64)
    if (lens == 0).any():
        raise ValueError("one or more references are empty strings")
    if len(refs) == 0:
        return np.zeros(0, np.float64)
    dist = engine.levenshtein(refs, hyps)
    return dist.astype(np.float64) / lens.astype(np.float64)


def hyps_cer(hyps_text: Dict[str, Dict[str, str]], ref_text: Dict[str, str]) -> Dict[str, Dict[str, float]]:
    """espnet_data/preprocess/main.py:51-60 for a whole split."""
    refs, hyps = [], []
    for utt, hs in hyps_text.items():
        for h in hs.values():
            refs.append(ref_text[utt])
            hyps.append(h)
    c = pair_cer(refs, hyps)
    out, i = {}, 0
    for utt, hs in hyps_text.items():
        out[utt] = {}
        for k in hs:
            out[utt][k] = float(c[i])
            i += 1
    return out


class BaseFunction():
    def __init__(self, config) -> None:
        self.config = config

    def score(self):
        raise NotImplementedError


class CerScoreFunction(BaseFunction):
    """RMBR/utility_functions.py:24-33: similarity = 1 - cer(ref, cand)."""

    def score(self, cands, refs):
        return [1 - e for e in pair_cer(refs, cands).tolist()]


def mbr_decode(n_best: int, all_hyps: List[List[str]], utility_function: BaseFunction):
    """RMBR/mbr.py:5-27: expected-utility argmax over the top n_best hypotheses."""
    import torch
    cands, refs = [], []
    for utt_hyps in all_hyps:
        for hyp_i_pos in range(n_best):
            cands += [utt_hyps[hyp_i_pos]] * (n_best - 1)
            refs += utt_hyps[:hyp_i_pos] + utt_hyps[hyp_i_pos + 1:n_best]
    scores = utility_function.score(cands, refs)
    if not isinstance(scores, torch.Tensor):
        scores = torch.tensor(scores, dtype=torch.float)
    scores = scores.reshape(len(all_hyps), n_best, n_best - 1)
    scores = scores.sum(dim=-1)
    max_value_index = scores.argmax(dim=-1).cpu()
    predictions = [all_hyps[utt_id][v_index] for utt_id, v_index in enumerate(max_value_index)]
    return predictions, scores
