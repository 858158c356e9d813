"""GPU bring-up diagnostics: each step runs in its own process (a trapped kernel must
not poison the next step) and prints verbose numbers instead of asserting.

    python tools/gpu_diag.py            # all steps, each under `timeout`
    python tools/gpu_diag.py gemm       # one step
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def step_lev():
    import numpy as np
    from asr_rescoring_b200 import engine
    z = np.load(os.path.join(GOLD, "levenshtein_kat.npz"))
    t = time.time()
    got = engine.levenshtein_packed(z["ref_cp"], z["ref_off"], z["hyp_cp"], z["hyp_off"], np.arange(len(z["dist"]), dtype=np.int32))
    print("lev KAT mismatches:", int((got != z["dist"]).sum()), "of", len(got), "sum", int(got.sum()), f"{time.time()-t:.3f}s")
    # long strings across the KB template range
    import oracle
    rng = np.random.default_rng(0)
    for n in (0, 1, 31, 32, 33, 64, 65, 200, 700):
        refs = ["".join(chr(0x4E00 + int(c)) for c in rng.integers(0, 30, size=max(n + int(rng.integers(-3, 4)), 0))) for _ in range(50)]
        hyps = ["".join(chr(0x4E00 + int(c)) for c in rng.integers(0, 30, size=n)) for _ in range(50)]
        exp = oracle.levenshtein_strings(refs, hyps)
        got = engine.levenshtein(refs, hyps)
        print(f"lev len~{n}: mismatches {int((exp != got).sum())}")


def step_combiner():
    import numpy as np
    from asr_rescoring_b200 import engine
    z = np.load(os.path.join(GOLD, "combiner_golden.npz"))
    import oracle
    refs_cp, ref_off, hyp_cp, hyp_off = z["ref_cp"], z["ref_off"], z["hyp_cp"], z["hyp_off"]
    N, nb = z["am"].shape
    pair_ref = np.repeat(np.arange(N, dtype=np.int32), nb)
    dist = engine.levenshtein_packed(refs_cp, ref_off, hyp_cp, hyp_off, pair_ref).reshape(N, nb)
    dist_o = oracle.levenshtein_batch(refs_cp, ref_off, hyp_cp, hyp_off, pair_ref).reshape(N, nb)
    print("dist mismatches", int((dist != dist_o).sum()))
    arg, es = engine.rescore_sweep(z["am"], z["lm"], z["lens"], dist, z["weights"], "B")
    print("argmax mismatches vs reference golden:", int((arg != z["argmax"]).sum()), "of", arg.size)
    arg_o, es_o = oracle.rescore_sweep(z["am"], z["lm"], z["lens"], dist_o, z["weights"], 0)
    print("edit_sum equal to oracle:", bool((es == es_o).all()), es[:5], es_o[:5])
    for i, wi in enumerate(z["score_idx"]):
        with np.errstate(all="ignore"):
            s = engine.rescore_scores(z["am"], z["lm"], z["lens"], z["weights"][wi], "B")
        same = (s.view(np.uint64) == z["scores"][i].view(np.uint64)) | (np.isnan(s) & np.isnan(z["scores"][i]))
        print(f"scores w[{wi}] bit-exact:", bool(same.all()), "n diff", int((~same).sum()))


def _mk(cfg_name="tiny", perturb=True, seed=10, chunk=0):
    from asr_rescoring_b200 import engine, synth
    cfg = {"tiny": synth.BERT_TINY, "base": synth.BERT_BASE_CHINESE}[cfg_name]
    sd = synth.random_init_state_dict(cfg, seed, perturb)
    return cfg, sd, engine.PllScorer(sd, cfg, max_chunk_tokens=chunk)


def step_expand():
    import numpy as np
    import oracle
    from asr_rescoring_b200 import synth
    cfg, sd, sc = _mk()
    nb = synth.make_nbest(20, 5, seed=1)
    nb.hyps[0][0] = ""
    nb.hyps[3][2] = nb.hyps[3][2][:1]
    tok, off = nb.packed_tokens(cfg["vocab"])
    ids, mp, lab = sc.expand(tok, off)
    ids_o, mp_o, lab_o = oracle.expand(tok, off)
    print("expand ids equal", bool((ids == ids_o).all()), "mask_pos", bool((mp == mp_o).all()), "labels", bool((lab == lab_o).all()), len(ids))


def step_gemm_simt():
    import torch
    from asr_rescoring_b200 import engine
    torch.manual_seed(0)
    for (M, N, K) in ((300, 256, 128), (77, 512, 64)):
        A = torch.randn(M, K, device="cuda").bfloat16()
        W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        b = torch.randn(N, device="cuda")
        ref = A.float() @ W.float().T + b
        for epi in (0, 1, 2, 3):
            out = engine.debug_gemm(A, W, b, epi, simt=True).float()
            r = torch.nn.functional.gelu(ref) if epi in (1, 3) else ref
            torch.cuda.synchronize()
            print(f"simt M{M} N{N} K{K} epi{epi}: max abs err {(out - r).abs().max().item():.3e}")


def step_gemm():
    import torch
    from asr_rescoring_b200 import engine
    torch.manual_seed(0)
    shapes = ((128, 256, 64), (128, 256, 128), (256, 512, 256), (300, 768, 768), (5, 256, 64), (1000, 2304, 768),
              (4096, 3072, 768), (4099, 768, 3072), (40000, 768, 768))
    for (M, N, K) in shapes:
        A = torch.randn(M, K, device="cuda").bfloat16()
        W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        b = torch.randn(N, device="cuda")
        ref = A.float() @ W.float().T + b
        for epi in (2, 0, 1, 3):
            out = engine.debug_gemm(A, W, b, epi).float()
            torch.cuda.synchronize()
            r = torch.nn.functional.gelu(ref) if epi in (1, 3) else ref
            err = (out - r).abs()
            tol = 2e-2 * r.abs().max().item() if epi in (0, 1) else 2e-3
            bad = err > tol
            msg = f"tcgen05 M{M} N{N} K{K} epi{epi}: max abs err {err.max().item():.3e} (ref max {r.abs().max().item():.2f}) bad {int(bad.sum())}/{bad.numel()}"
            if bad.any():
                rows = bad.any(1).nonzero().flatten()
                cols = bad.any(0).nonzero().flatten()
                msg += f" | bad rows {rows[:8].tolist()}..{rows[-3:].tolist()} n{len(rows)} cols {cols[:8].tolist()}..{cols[-3:].tolist()} n{len(cols)}"
                msg += f" | out[0,:4] {out[0,:4].tolist()} ref {r[0,:4].tolist()}"
            print(msg)
    # throughput
    for (M, N, K, epi) in ((65536, 2304, 768, 0), (65536, 768, 768, 2), (65536, 3072, 768, 1), (65536, 768, 3072, 2),
                           (262144, 3072, 768, 1), (262144, 768, 3072, 2)):
        A = torch.randn(M, K, device="cuda").bfloat16()
        W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        b = torch.randn(N, device="cuda")
        for _ in range(3):
            engine.debug_gemm(A, W, b, epi)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        it = 10
        for _ in range(it):
            engine.debug_gemm(A, W, b, epi)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        print(f"tcgen05 perf M{M} N{N} K{K} epi{epi}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s")
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        Wt = W.T.contiguous()
        for _ in range(3):
            torch.matmul(A, Wt)
        t0.record()
        for _ in range(it):
            torch.matmul(A, Wt)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / it
        print(f"   cuBLAS same shape: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s")


def step_hidden(cfg_name="tiny"):
    import numpy as np
    import torch
    from oracle import pll_oracle
    from asr_rescoring_b200 import synth
    cfg, sd, sc = _mk(cfg_name)
    nb = synth.make_nbest(4, 3, seed=2)
    tok, off = nb.packed_tokens(cfg["vocab"])
    L = np.diff(off)
    for upto in (0, 1, cfg["num_layers"]):
        got = sc.hidden(tok, off, upto)
        exp = []
        for h in range(len(L)):
            toks = [int(t) for t in tok[off[h]:off[h + 1]]]
            for r in pll_oracle.expand_rows(toks, "u", "h"):
                ids = torch.tensor([r["input_ids"]])
                exp.append(pll_oracle.bert_mlm_logits(sd, cfg, ids, torch.ones_like(ids), upto_layer=upto, return_hidden=True)[0])
        exp = torch.cat(exp)
        err = (got - exp).abs()
        print(f"hidden[{cfg_name}] upto {upto}: max abs err {err.max().item():.3e} mean {err.mean().item():.3e} (ref absmax {exp.abs().max().item():.2f}) worst row {int(err.max(1).values.argmax())}")


def step_pll():
    import numpy as np
    from asr_rescoring_b200 import engine, synth
    gold = json.load(open(os.path.join(GOLD, "pll_golden.json")))
    for case in gold["cases"]:
        sd = synth.random_init_state_dict(case["cfg"], case["seed"], case["perturb"])
        sc = engine.PllScorer(sd, case["cfg"])
        t = time.time()
        got = sc.score_hyps(case["hyps"])
        dt = time.time() - t
        diffs = [abs(got[u][h] - case["pll"][u][h]) for u in case["hyps"] for h in case["hyps"][u]]
        print(f"pll[{case['name']}]: max |dPLL| {max(diffs):.4f} mean {np.mean(diffs):.4f} n {len(diffs)} ({dt:.2f}s) sample {list(got[next(iter(got))].values())[:3]} vs {list(case['pll'][next(iter(got))].values())[:3]}")
        print("   stats", {k: v for k, v in sc.stats().items() if k in ("kernel_launches", "chunks", "copies_scored", "tokens_expanded")})
        sc.close()


def step_bench():
    import numpy as np
    from asr_rescoring_b200 import engine, synth
    cfg = synth.BERT_BASE_CHINESE
    sd = synth.random_init_state_dict(cfg, 10, False)
    sc = engine.PllScorer(sd, cfg, max_chunk_tokens=1 << 20)
    nb = synth.make_nbest(1000, 10, seed=0)
    tok, off = nb.packed_tokens()
    sc.set_timing(True)
    for i in range(3):
        t = time.time()
        pll = sc.score_packed(tok, off)
        dt = time.time() - t
        st = sc.stats()
        print(f"bench 1000x10: {dt:.3f}s {len(off)-1} hyps -> {(len(off)-1)/dt:.0f} hyps/s; total_ms {st['last_total_ms']:.1f} gemm_ms {st['last_gemm_ms']:.1f} by kind { {k: round(v,1) for k,v in st['gemm_ms_by_kind'].items()} }")
    print("pll sample", pll[:4], "finite", bool(np.isfinite(pll).all()))


def step_numerics():
    """Error distribution of the GPU path vs the fp32 oracle, bf16 and fp16 operands."""
    import numpy as np
    from oracle import pll_oracle
    from asr_rescoring_b200 import engine, synth
    cfg = synth.BERT_BASE_CHINESE
    sd = synth.random_init_state_dict(cfg, 10, False)
    nb = synth.make_nbest(int(os.environ.get("NUM_UTTS", "30")), 5, seed=321)
    tok, off = nb.packed_tokens()
    L = np.diff(off)
    hyps = {"u": {f"hyp_{i + 1}": [int(t) for t in tok[off[i]:off[i + 1]]] for i in range(len(L))}}
    t = time.time()
    exp = pll_oracle.score_hyps(sd, cfg, hyps)
    ref = np.array([exp["u"][f"hyp_{i + 1}"] for i in range(len(L))])
    print(f"oracle: {len(L)} hyps in {time.time()-t:.1f}s")
    for dt in ("bf16", "fp16"):
        with engine.PllScorer(sd, cfg, operand_dtype=dt) as sc:
            got = sc.score_packed(tok, off)
        e = got - ref
        print(f"{dt}: mean|e| {np.abs(e).mean():.4f} max|e| {np.abs(e).max():.4f} std {e.std():.4f} bias {e.mean():+.4f} "
              f"per-sqrt(L) std {(e/np.sqrt(L)).std():.4f} frac>0.05 {np.mean(np.abs(e)>0.05):.4f} "
              f"p99 {np.percentile(np.abs(e),99):.4f}")


STEPS = {"lev": step_lev, "combiner": step_combiner, "expand": step_expand, "gemm_simt": step_gemm_simt,
         "gemm": step_gemm, "hidden": step_hidden, "hidden_base": lambda: step_hidden("base"), "pll": step_pll,
         "bench": step_bench, "numerics": step_numerics}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        STEPS[sys.argv[1]]()
    else:
        for name in STEPS:
            print(f"===== {name}", flush=True)
            t = time.time()
            r = subprocess.run(["timeout", "300", sys.executable, os.path.abspath(__file__), name], cwd=ROOT)
            print(f"===== {name} exit {r.returncode} ({time.time()-t:.1f}s)", flush=True)
