"""oracle/ — CPU restatements of the reference path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package.  The product package
(asr-rescoring_b200/) never does; it fails loudly without its CUDA library.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        f64p = ctypes.POINTER(ctypes.c_double)
        L.oracle_levenshtein.restype = ctypes.c_int32
        L.oracle_levenshtein.argtypes = [i32p, ctypes.c_int32, i32p, ctypes.c_int32]
        L.oracle_levenshtein_batch.restype = None
        L.oracle_levenshtein_batch.argtypes = [i32p, i64p, i32p, i64p, i32p, ctypes.c_int32, i32p]
        L.oracle_rescore_scores.restype = None
        L.oracle_rescore_scores.argtypes = [f64p, f64p, i64p, ctypes.c_int32, ctypes.c_int32,
                                            ctypes.c_double, ctypes.c_int32, f64p]
        L.oracle_rescore_sweep.restype = None
        L.oracle_rescore_sweep.argtypes = [f64p, f64p, i64p, i32p, ctypes.c_int32, ctypes.c_int32,
                                           f64p, ctypes.c_int32, ctypes.c_int32, i32p, i64p]
        L.oracle_expand.restype = None
        L.oracle_expand.argtypes = [i32p, i64p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                    ctypes.c_int32, i32p, i32p, i32p]
        _LIB = L
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def pack_strings(strings):
    """list[str] -> (int32 code points, int64 offsets)."""
    off = np.zeros(len(strings) + 1, dtype=np.int64)
    for i, s in enumerate(strings):
        off[i + 1] = off[i] + len(s)
    cp = np.fromiter((ord(c) for s in strings for c in s), dtype=np.int32, count=int(off[-1]))
    return cp, off


def levenshtein_batch(ref_cp, ref_off, hyp_cp, hyp_off, pair_ref) -> np.ndarray:
    ref_cp = np.ascontiguousarray(ref_cp, np.int32)
    hyp_cp = np.ascontiguousarray(hyp_cp, np.int32)
    ref_off = np.ascontiguousarray(ref_off, np.int64)
    hyp_off = np.ascontiguousarray(hyp_off, np.int64)
    pair_ref = np.ascontiguousarray(pair_ref, np.int32)
    out = np.zeros(len(pair_ref), np.int32)
    lib().oracle_levenshtein_batch(_p(ref_cp, ctypes.c_int32), _p(ref_off, ctypes.c_int64),
                                   _p(hyp_cp, ctypes.c_int32), _p(hyp_off, ctypes.c_int64),
                                   _p(pair_ref, ctypes.c_int32), len(pair_ref), _p(out, ctypes.c_int32))
    return out


def levenshtein_strings(refs, hyps) -> np.ndarray:
    rc, ro = pack_strings(refs)
    hc, ho = pack_strings(hyps)
    return levenshtein_batch(rc, ro, hc, ho, np.arange(len(hyps), dtype=np.int32))


def rescore_scores(am, lm, length, weight, variant=0) -> np.ndarray:
    am = np.ascontiguousarray(am, np.float64)
    lm = np.ascontiguousarray(lm, np.float64)
    length = np.ascontiguousarray(length, np.int64)
    N, nb = am.shape
    out = np.zeros((N, nb), np.float64)
    lib().oracle_rescore_scores(_p(am, ctypes.c_double), _p(lm, ctypes.c_double), _p(length, ctypes.c_int64),
                                N, nb, float(weight), variant, _p(out, ctypes.c_double))
    return out


def rescore_sweep(am, lm, length, dist, weights, variant=0):
    am = np.ascontiguousarray(am, np.float64)
    lm = np.ascontiguousarray(lm, np.float64)
    length = np.ascontiguousarray(length, np.int64)
    dist = np.ascontiguousarray(dist, np.int32)
    weights = np.ascontiguousarray(weights, np.float64)
    N, nb = am.shape
    W = len(weights)
    arg = np.zeros((W, N), np.int32)
    es = np.zeros(W, np.int64)
    lib().oracle_rescore_sweep(_p(am, ctypes.c_double), _p(lm, ctypes.c_double), _p(length, ctypes.c_int64),
                               _p(dist, ctypes.c_int32), N, nb, _p(weights, ctypes.c_double), W, variant,
                               _p(arg, ctypes.c_int32), _p(es, ctypes.c_int64))
    return arg, es


def expand(hyp_tokens, hyp_off, cls_id=101, sep_id=102, mask_id=103):
    hyp_tokens = np.ascontiguousarray(hyp_tokens, np.int32)
    hyp_off = np.ascontiguousarray(hyp_off, np.int64)
    L = np.diff(hyp_off)
    ids = np.zeros(int((L * (L + 2)).sum()), np.int32)
    mp = np.zeros(int(L.sum()), np.int32)
    lab = np.zeros(int(L.sum()), np.int32)
    lib().oracle_expand(_p(hyp_tokens, ctypes.c_int32), _p(hyp_off, ctypes.c_int64), len(L), cls_id, sep_id,
                        mask_id, _p(ids, ctypes.c_int32), _p(mp, ctypes.c_int32), _p(lab, ctypes.c_int32))
    return ids, mp, lab
