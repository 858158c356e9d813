"""Drop-in for the reference's ``rescore.py`` (AM/LM interpolation over a weight grid,
per-utterance argmax, corpus CER).

    python rescore.py --config rescore.yaml

Same flag, YAML keys (rescore.yaml:1-14), log file (``<output_path>/rescore.log`` with the
config, the source of find_best_weight/rescore, best weight, dev and test CER) and
importable functions as the reference (rescore.py:13-58).  The arithmetic runs in
libpllb200.so: one warp-per-pair Levenshtein launch over all (ref, hyp_k) pairs, then one
fp64 sweep kernel for the whole weight grid — instead of 101 x (numpy + jiwer.cer).

Optional keys (defaulted, absent from the reference YAML): ``formula`` = "B" (current
source, rescore.py:51) | "A" (rescore_result/MLM_PLL/rescore.log:28) | "C"
(rescore_result/RMBR/BertScore/rescore_mbr_normalize.log:29); ``weight_grid`` =
[start, stop, step] for np.arange (default [0.0, 1.01, 0.01], rescore.py:37).
"""
from __future__ import annotations

import inspect
import json
import logging
import os
import sys
from typing import Dict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_PARENT = os.path.dirname(_HERE)
if _PKG_PARENT not in sys.path:
    sys.path.insert(0, _PKG_PARENT)

from asr_rescoring_b200 import engine, shard  # noqa: E402
from asr_rescoring_b200.util.arg_parser import ArgParser  # noqa: E402


def dict_to_list(dict):
    """rescore.py:13-23."""
    all_scores = []
    for utt_id, hyps in dict.items():
        if isinstance(hyps, Dict):
            all_scores.append([hyp for hyp_id, hyp in hyps.items()])
        else:
            all_scores.append(hyps)
    return all_scores


def _formula(config) -> str:
    return getattr(config, "formula", "B")


def _grid(config) -> np.ndarray:
    g = getattr(config, "weight_grid", None)
    return np.arange(*g) if g else np.arange(0.0, 1.01, 0.01)


def _hyps_len(hyps, n_best):
    # rescore.py:28-35: Python len() of the raw string, first n_best hypotheses
    return [[len(hyp) for hyp in utt_hyps[:n_best]] for utt_hyps in hyps]


def cer(ref, hyp) -> float:
    """jiwer.cer stand-in (rescore.py:8,40,118): sum of character edit distances over the
    total reference length; strings stripped; an empty reference raises, as jiwer does."""
    if isinstance(ref, str):
        ref = [ref]
    if isinstance(hyp, str):
        hyp = [hyp]
    if len(ref) != len(hyp):
        raise ValueError("reference and hypothesis lists differ in length")
    total = 0
    for r in ref:
        if len(r.strip()) == 0:
            raise ValueError("one or more references are empty strings")
        total += len(r.strip())
    return float(int(engine.levenshtein(ref, hyp).sum())) / float(total)


def _pack_candidates(hyps, n_best):
    """First n_best candidates of every utterance, flattened: (code points, offsets) of the stripped
    strings for the Levenshtein kernel and the RAW lengths [N, n_best] that rescore.py:28-35 uses."""
    counts = {len(utt[:n_best]) for utt in hyps}
    if len(counts) > 1:
        raise ValueError("every utterance needs the same number of hypotheses (rectangular N-best), as np.array "
                         "in rescore.py:48-50 requires")
    flat = [h for utt in hyps for h in utt[:n_best]]
    hc, ho, raw_len = engine.pack_stripped(flat)
    width = counts.pop() if counts else 0
    return hc, ho, raw_len.reshape(len(hyps), width), width


def pair_distances(hyps, ref, n_best, packed=None) -> np.ndarray:
    """int32 [N, n_best] edit distance of every candidate to its reference (one launch)."""
    hc, ho, _, width = packed if packed is not None else _pack_candidates(hyps, n_best)
    rc, ro = engine.pack_strings([r.strip() for r in ref])
    pair_ref = np.repeat(np.arange(len(ref), dtype=np.int32), width)
    return engine.levenshtein_packed(rc, ro, hc, ho, pair_ref).reshape(len(ref), -1)


def sweep(am, lm, hyps, ref, config, weights=None):
    """(weights, cer per weight, argmax [W, N]) for the whole grid.

    Under torchrun (an initialised process group, one process per GPU) the utterances are split
    into contiguous blocks, one per rank: Levenshtein distances, scores and the per-utterance
    argmax are local; the only exchanges are ONE integer all_reduce of the per-weight edit sums
    plus the reference length (the corpus CER of rescore.py:40 is a global sum) and a gather of
    the argmax blocks.  Integer sums are associative, so every world size gives the same CERs."""
    n_best = config.n_best
    for r in ref:
        if len(r.strip()) == 0:
            raise ValueError("one or more references are empty strings")
    weights = _grid(config) if weights is None else np.asarray(weights, np.float64)
    rank, world = shard.active_world()
    N = len(ref)
    lo, hi = shard.block_range(N, rank, world)
    hyps_l, ref_l = hyps[lo:hi], ref[lo:hi]
    packed = _pack_candidates(hyps_l, n_best)          # one pass over the strings: lengths + code points
    hyps_len = packed[2]
    am_ = np.array(am[lo:hi], np.float64).reshape(hi - lo, -1)[:, :n_best]
    lm_ = np.array(lm[lo:hi], np.float64).reshape(hi - lo, -1)
    counts = np.zeros(len(weights) + 1, np.int64)
    argmax = np.zeros((len(weights), hi - lo), np.int32)
    if hi > lo:
        dist = pair_distances(hyps_l, ref_l, n_best, packed)
        argmax, edits = engine.rescore_sweep(am_, lm_, hyps_len, dist, weights, _formula(config))
        counts[:-1] = edits
        counts[-1] = sum(len(r.strip()) for r in ref_l)
    if world > 1:
        counts = shard.reduce_counts(counts)
        argmax = shard.gather_blocks(argmax, N, axis=1)
    return weights, counts[:-1].astype(np.float64) / float(counts[-1]), argmax


def find_best_weight(am, lm, hyps, ref, config):
    """rescore.py:25-45: first weight of the grid with the strictly smallest CER."""
    best_cer = sys.float_info.max
    weights, cers, _ = sweep(am, lm, hyps, ref, config)
    for weight, error in zip(weights, cers):
        error = float(error)
        if error < best_cer:
            best_cer = error
            best_weight = weight
    return best_weight, best_cer


def rescore(weight, hyps_len, am, lm, config):
    """rescore.py:47-53 — float64 [N, n_best], bit-identical to the numpy expression."""
    am = np.array(am, np.float64)[:, :config.n_best]
    lm = np.array(lm, np.float64)
    hyps_len = np.array(hyps_len, np.int64)
    return engine.rescore_scores(am, lm, hyps_len, float(weight), _formula(config))


def get_highest_score_hyp(final_score, hyps):
    """rescore.py:55-58."""
    max_score_hyp_index = np.argmax(final_score, axis=-1)
    return [ht[index] for ht, index in zip(hyps, max_score_hyp_index)]


if __name__ == "__main__":
    arg_parser = ArgParser()
    config = arg_parser.parse()
    rank, world, _local = shard.init_process_group() if shard.dist_env()[1] > 1 else (0, 1, 0)
    if world > 1:
        import torch
        torch.cuda.set_device(_local)

    logging.basicConfig(
        filename=(config.output_path + "/rescore.log") if rank == 0 else os.devnull,
        filemode='w',
        format='%(asctime)s,%(msecs)d %(name)s %(levelname)s %(message)s',
        datefmt='%H:%M:%S',
        level=logging.INFO
    )
    logging.info(config)
    logging.info("\n" + inspect.getsource(find_best_weight))
    logging.info("\n" + inspect.getsource(rescore))

    def _load(path):
        return dict_to_list(json.load(open(path, "r", encoding="utf-8")))

    dev_am, dev_lm = _load(config.dev_am_path), _load(config.dev_lm_path)
    dev_hyps, dev_ref = _load(config.dev_hyps_text_path), _load(config.dev_ref_text_path)
    best_weight, best_cer = find_best_weight(dev_am, dev_lm, dev_hyps, dev_ref, config)
    logging.info("best_weight: " + str(best_weight))
    logging.info("dev cer: " + str(best_cer))
    if rank == 0:
        print("best_weight: ", best_weight)
        print("dev cer: ", best_cer)

    test_am, test_lm = _load(config.test_am_path), _load(config.test_lm_path)
    test_hyps, test_ref = _load(config.test_hyps_text_path), _load(config.test_ref_text_path)
    _, test_cers, _ = sweep(test_am, test_lm, test_hyps, test_ref, config, weights=[best_weight])
    test_cer = float(test_cers[0])
    logging.info("test cer: " + str(test_cer))
    if rank == 0:
        print("test cer: ", test_cer)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
