"""Large PLL golden for BASELINE.json configs[1]/[2] (C2/C3), from the UNMODIFIED reference.

    python oracle/make_golden_c2.py [--utts 2000] [--part 250] [--threads 6]
    (needs /root/reference; build container only — hours of CPU, resumable)

Runs the reference's own set_dataloader + run_one_epoch(train_mode=False, do_scoring=True)
(/root/reference/MLM_PLL/main.py, imported by path exactly as oracle/make_golden.py does) on
transformers.BertForMaskedLM carrying synth.random_init_state_dict(BERT_BASE_CHINESE, 10), over the
first --utts utterances x 10-best of synth.make_nbest(7176, 10, seed=0) — the workload bench.py
times.  Parts of --part utterances are cached under oracle/_c2_parts/ (git-ignored) so the run can
be stopped and resumed (`--c4` writes tests/golden/c4_pll_golden.json instead: 48 hypotheses of
the config-4 shape, 24-layer / H 1024, L up to 64); the merged result is tests/golden/c2_pll_golden.npz:
    pll float64[n_utts*10]   per-hypothesis PLL in (utterance, k) order
    tok_crc / off_crc        CRC32 of the packed token ids / offsets the PLLs belong to
The -m gpu tests compare the CUDA path against it (|dPLL| <= 0.05 nats per hypothesis, rescored
1-best identical on >= 99.9 % of utterances at the weight the reference picks).
"""
from __future__ import annotations

import argparse
import os
import sys
import time
import zlib
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402  (shims + import-by-path helpers)
from oracle import pll_oracle  # noqa: E402
from asr_rescoring_b200 import synth  # noqa: E402

PARTS = os.path.join(HERE, "_c2_parts")


def c4_hyps():
    """48 hypotheses of the config-4 shape (50-best lists of 8..64 tokens): the synthetic lists of
    synth.make_nbest(12, 4, seed=4, min_len=8, max_len=64) with four lengths pinned to the extremes."""
    nb = synth.make_nbest(12, 4, seed=4, min_len=8, max_len=64)
    tok, off = nb.packed_tokens()
    lists = [[int(t) for t in tok[off[i]:off[i + 1]]] for i in range(len(off) - 1)]
    filler = [t for l in lists for t in l]
    for i, L in ((0, 64), (13, 64), (26, 64), (5, 8)):
        lists[i] = (lists[i] + filler[100 * i:100 * i + 64])[:L]
    return lists


def golden_c4(ref_main, threads):
    """BASELINE.json configs[3] shape: bert-large-shaped (24 layers, H 1024), L up to 64."""
    cfg = synth.BERT_LARGE_SHAPED
    sd = synth.random_init_state_dict(cfg, 10)
    hf = mg._hf_model(cfg, sd)
    lists = c4_hyps()
    rows, skel = [], {"u": {}}
    for i, toks in enumerate(lists):
        skel["u"][f"hyp_{i + 1}"] = 0
        rows += pll_oracle.expand_rows(toks, "u", f"hyp_{i + 1}")
    t0 = time.time()
    loader = ref_main.set_dataloader(SimpleNamespace(batch_size=32, num_worker=0), ref_main.MyDataset(rows), True)
    with torch.no_grad():
        out = ref_main.run_one_epoch(config=SimpleNamespace(device="cpu"), model=hf, dataloader=loader,
                                     output_score=skel, train_mode=False, do_scoring=True)
    pll = [out["u"][f"hyp_{i + 1}"] for i in range(len(lists))]
    import json
    json.dump(dict(generator="oracle/make_golden_c2.py --c4: /root/reference/MLM_PLL/main.py run_one_epoch (unmodified), "
                             f"transformers {__import__('transformers').__version__}, torch {torch.__version__}",
                   cfg=cfg, seed=10, tokens=lists, pll=pll),
              open(os.path.join(mg.GOLD, "c4_pll_golden.json"), "w"))
    print(f"c4: {len(lists)} hyps, L {min(map(len, lists))}..{max(map(len, lists))}, {len(rows)} copies in "
          f"{time.time() - t0:.0f} s", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=2000)
    ap.add_argument("--part", type=int, default=250)
    ap.add_argument("--threads", type=int, default=6)
    ap.add_argument("--c4", action="store_true", help="generate tests/golden/c4_pll_golden.json instead")
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    os.makedirs(PARTS, exist_ok=True)
    mg._install_shims()
    ref_main = mg._import_by_path("ref_mlm_pll_main", os.path.join(mg.REF, "MLM_PLL", "main.py"),
                                  os.path.join(mg.REF, "MLM_PLL"))
    if args.c4:
        return golden_c4(ref_main, args.threads)
    cfg = synth.BERT_BASE_CHINESE
    sd = synth.random_init_state_dict(cfg, 10)
    hf = mg._hf_model(cfg, sd)
    nb = synth.make_nbest(7176, 10, seed=0)
    tok, off = nb.packed_tokens()
    n_best = 10
    done = []
    for a in range(0, args.utts, args.part):
        b = min(a + args.part, args.utts)
        path = os.path.join(PARTS, f"pll_{a:05d}_{b:05d}.npy")
        if not os.path.exists(path):
            rows, skel = [], {}
            for u in range(a, b):
                skel[nb.utt_ids[u]] = {}
                for k in range(n_best):
                    i = u * n_best + k
                    toks = [int(t) for t in tok[off[i]:off[i + 1]]]
                    skel[nb.utt_ids[u]][f"hyp_{k + 1}"] = 0
                    rows += pll_oracle.expand_rows(toks, nb.utt_ids[u], f"hyp_{k + 1}")
            t0 = time.time()
            loader = ref_main.set_dataloader(SimpleNamespace(batch_size=32, num_worker=0), ref_main.MyDataset(rows), True)
            with torch.no_grad():
                out = ref_main.run_one_epoch(config=SimpleNamespace(device="cpu"), model=hf, dataloader=loader,
                                             output_score=skel, train_mode=False, do_scoring=True)
            pll = np.array([out[nb.utt_ids[u]][f"hyp_{k + 1}"] for u in range(a, b) for k in range(n_best)], np.float64)
            np.save(path, pll)
            print(f"utts [{a},{b}): {len(rows)} copies in {time.time() - t0:.0f} s", flush=True)
        done.append(np.load(path))
        n = b
        pll = np.concatenate(done)
        np.savez_compressed(
            os.path.join(mg.GOLD, "c2_pll_golden.npz"), pll=pll, n_utts=np.int64(n), n_best=np.int64(n_best),
            tok_crc=np.uint32(zlib.crc32(tok[:off[n * n_best]].tobytes())),
            off_crc=np.uint32(zlib.crc32(off[:n * n_best + 1].tobytes())),
            generator=np.array("oracle/make_golden_c2.py: /root/reference/MLM_PLL/main.py run_one_epoch (unmodified), "
                               f"transformers {__import__('transformers').__version__}, torch {torch.__version__}, "
                               "synth.make_nbest(7176, 10, seed=0), synth.random_init_state_dict(BERT_BASE_CHINESE, 10)"))


if __name__ == "__main__":
    main()
