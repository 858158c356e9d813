"""The C-ABI library builds, loads and exports every symbol include/pllb.h declares; with
no GPU every compute entry point fails loudly (there is no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from asr_rescoring_b200 import _lib, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "pllb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pllb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pllb.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert lib.pllb_abi_version() == 1


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.ModelDesc) == 44
    assert ctypes.sizeof(_lib.LayerWeights) == 16 * 8
    assert ctypes.sizeof(_lib.Weights) == 12 * 8
    assert ctypes.sizeof(_lib.Stats) == 8 * 5 + 8 + 4 + 4 + 8
    assert ctypes.sizeof(_lib.TrainDesc) == 56 and _lib.TrainDesc.seed.offset == 32 and _lib.TrainDesc.max_seq.offset == 48


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "asr-rescoring_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)
                assert "liboracle" not in src


@pytest.mark.skipif(_lib.load().pllb_device_count() > 0, reason="a GPU is present")
def test_compute_calls_fail_loudly_without_gpu():
    with pytest.raises(_lib.PllbError):
        engine.levenshtein(["ab"], ["ac"])
    with pytest.raises(_lib.PllbError):
        engine.rescore_sweep(np.zeros((1, 2)), np.zeros((1, 2)), np.ones((1, 2), np.int64), None, [0.5])
    from asr_rescoring_b200 import synth
    with pytest.raises(_lib.PllbError):
        engine.PllScorer(synth.random_init_state_dict(synth.BERT_TINY, 1), synth.BERT_TINY)
    with pytest.raises(_lib.PllbError):
        engine.MlmTrainer(synth.random_init_state_dict(synth.BERT_TINY, 1), synth.BERT_TINY)
