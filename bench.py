"""bench.py — N-best PLL hypotheses/sec of the MLM-PLL scoring path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step is one pass of the whole hot path over the workload: stage 1-3 (on-device masked
copy expansion, BERT encoder, masked-row MLM head -> per-hypothesis PLL) and stage 4 (all
(ref, hyp) Levenshtein distances + 101-point weight sweep with per-utterance argmax).
Workload (config.workload = "c2"): AISHELL-1-test-shaped 7 176 utterances x 10-best,
random-init bert-base-chinese (BASELINE.json configs[1]); synthetic data (synth.py).

value  : device-resident inputs, whole job, max over ranks, CUDA-event timed.
e2e    : the same through the host-buffer C-ABI calls (pllb_score_host,
         pllb_levenshtein_host, pllb_rescore_sweep_host): H2D and D2H inside the timing.
roofline: the GEMM kernels (all launches of the last timed step), EXECUTED GEMM FLOPs (sum of
         2*M*N*K over the launches: the layer-0 projection runs on unique rows and the last
         layer on the consumed rows only, so this is below the SURVEY formula, which is also
         reported) / summed launch durations from CUDA events recorded on the launch stream.
stdout : exactly one JSON line (library banners are redirected to stderr).
cpu_baseline / --impl reference: the oracle port of the reference's torch CPU path
         (oracle/pll_oracle.py) on a bounded sample, all host threads.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "N-best PLL hypotheses/sec"
UNIT = "hyps/s"

WORKLOADS = {
    # name: (n_utts, n_best, model, min_len, max_len)
    "c1": (100, 10, "bert-base-chinese", None, None),
    "c2": (7176, 10, "bert-base-chinese", None, None),
    "c4": (7176, 50, "bert-large-shaped", 8, 64),
}


WORKLOAD_DESC = {
    "c1": "c1: 100 synthetic AISHELL-1-shaped utterances x 10-best, random-init bert-base-chinese, PLL + weight sweep",
    "c2": "c2: AISHELL-1-test-shaped 7176 utterances x 10-best, random-init bert-base-chinese, PLL + 101-point weight sweep",
    "c4": "c4: 7176 utterances x 50-best, length 8..64, bert-large-shaped encoder (24L/1024H), PLL + weight sweep",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--utts", type=int, default=0, help="override the number of utterances (debug)")
    ap.add_argument("--chunk-tokens", type=int, default=1 << 20)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-sample-hyps", type=int, default=400, help="hypotheses in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--operand-dtype", default=None, choices=["bf16", "fp16"],
                    help="GEMM operand type; default bf16, fp16 for the 24-layer c4 config (needed for the 0.05-nat bound)")
    return ap.parse_args()


def model_cfg(name):
    from asr_rescoring_b200 import synth
    return {"bert-base-chinese": synth.BERT_BASE_CHINESE, "bert-large-shaped": synth.BERT_LARGE_SHAPED}[name]


def gemm_flops(lengths, cfg):
    """Algorithmic FLOPs of the dense GEMMs: per hypothesis of L tokens,
    T = L+2: L*[T*NL*(8H^2+4HI) + 2H^2 + 2HV]."""
    H, I, NL, V = cfg["hidden"], cfg["intermediate"], cfg["num_layers"], cfg["vocab"]
    L = np.asarray(lengths, np.float64)
    return float((L * ((L + 2) * NL * (8 * H * H + 4 * H * I) + 2 * H * H + 2 * H * V)).sum())


def total_flops(lengths, cfg):
    H, NL = cfg["hidden"], cfg["num_layers"]
    L = np.asarray(lengths, np.float64)
    return gemm_flops(lengths, cfg) + float((L * NL * 4 * (L + 2) ** 2 * H).sum())


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,power.limit")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})

        def _num(col):
            vals = []
            for smp in self.samples:
                try:
                    vals.append(float(smp[col]))
                except (IndexError, ValueError):
                    pass
            return float(np.median(vals)) if vals else None

        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples), "power_w": _num(6), "power_limit_w": _num(7)}


def cpu_port_hyps_per_s(cfg, sd, tok, off, n_hyps, threads=None):
    """The oracle port of the reference CPU path on the first n_hyps hypotheses."""
    import torch
    from oracle import pll_oracle
    if threads is None:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    torch.set_num_threads(threads)
    hyps = {"u": {f"hyp_{i + 1}": [int(t) for t in tok[off[i]:off[i + 1]]] for i in range(n_hyps)}}
    t0 = time.perf_counter()
    pll_oracle.score_hyps(sd, cfg, hyps, batch_size=32)      # score.yaml:16 batch_size
    dt = time.perf_counter() - t0
    copies = int(np.diff(off[:n_hyps + 1]).sum())
    return n_hyps / dt, dt, copies, torch.get_num_threads()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops_sustained"]), float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"


def run_reference(args):
    """--impl reference: the reference's torch CPU path (oracle port) on bounded samples."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from asr_rescoring_b200 import synth
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core it can
    torch.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
    n_utts, n_best, model, lo, hi = WORKLOADS[args.workload]
    cfg = model_cfg(model)
    nb = synth.make_nbest(max(args.cpu_sample_hyps // n_best + 1, 4), n_best, seed=0, min_len=lo, max_len=hi)
    tok, off = nb.packed_tokens()
    sd = synth.random_init_state_dict(cfg, 10)
    n = min(args.cpu_sample_hyps, len(off) - 1)
    for _ in range(args.warmup):
        cpu_port_hyps_per_s(cfg, sd, tok, off, min(n, 8))
    times = []
    threads = torch.get_num_threads()
    for _ in range(args.steps):
        _, dt, copies, threads = cpu_port_hyps_per_s(cfg, sd, tok, off, n)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = n / (ms / 1e3)
    sample = f"first {n} hypotheses ({copies} masked copies) of workload {args.workload}, batch 32, fp32, per step"
    _emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "n_best": n_best, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


_JSON_FD = None


def _claim_stdout():
    """Keep fd 1 for the JSON line only: everything else that writes to stdout during the run
    (NCCL's version banner, library chatter) is sent to stderr."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict):
    sys.stdout.flush()
    sys.stderr.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    args = parse_args()
    _claim_stdout()
    if args.operand_dtype is None:
        args.operand_dtype = "fp16" if args.workload == "c4" else "bf16"
    if args.workload == "c4" and not args.utts:
        # the full C4 list (358 800 hypotheses of 8..64 tokens on a 24-layer encoder) is ~6 minutes
        # PER PASS on one B200; the bench line for it is quoted on a 150-utterance sample
        args.utts = 150
        print("bench: workload c4 runs on the first 150 utterances unless --utts is given", file=sys.stderr)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from asr_rescoring_b200 import _lib, engine, shard, synth

    rank, world, local = shard.dist_env()
    if world > 1:
        shard.init_process_group("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    _lib.require_device()

    n_utts, n_best, model, lo, hi = WORKLOADS[args.workload]
    if args.utts:
        n_utts = args.utts
    cfg = model_cfg(model)
    # weak scaling: every rank scores its own full-size workload (different seed per rank);
    # strong: one workload, utterances LPT-partitioned over ranks (BASELINE.json configs[2]).
    if args.scaling == "weak" or world == 1:
        nb = synth.make_nbest(n_utts, n_best, seed=rank, min_len=lo, max_len=hi)
        utt_sel = np.arange(n_utts)
    else:
        nb = synth.make_nbest(n_utts, n_best, seed=0, min_len=lo, max_len=hi)
        lens_all = [[len(h) for h in hs] for hs in nb.hyps]
        utt_sel = shard.lpt_partition(shard.utterance_costs(lens_all), world)[rank]
    hyps = [nb.hyps[u] for u in utt_sel]
    refs = [nb.refs[u] for u in utt_sel]
    am = np.ascontiguousarray(nb.am[utt_sel])
    flat = [h for hs in hyps for h in hs]
    off = np.zeros(len(flat) + 1, np.int64)
    np.cumsum([len(h) for h in flat], out=off[1:])
    tok = np.fromiter((synth.synthetic_token_id(c) for h in flat for c in h), np.int32, int(off[-1]))
    lens = np.diff(off)
    n_hyp, N = len(flat), len(hyps)
    hyp_len = lens.reshape(N, n_best).astype(np.int64)
    rc, ro = engine.pack_strings(refs)
    hc, ho = engine.pack_strings(flat)
    pair_ref = np.repeat(np.arange(N, dtype=np.int32), n_best)
    weights = np.arange(0.0, 1.01, 0.01)
    W = len(weights)

    sd = synth.random_init_state_dict(cfg, 10)
    scorer = engine.PllScorer(sd, cfg, device=local, max_chunk_tokens=args.chunk_tokens, operand_dtype=args.operand_dtype)

    # device-resident inputs for `value`
    t = lambda a: torch.from_numpy(a).to(dev)
    d_tok, d_am, d_len = t(tok), t(am), t(hyp_len)
    d_rc, d_ro, d_hc, d_ho, d_pr, d_w = t(rc), t(ro), t(hc), t(ho), t(pair_ref), t(weights)
    d_pll = torch.zeros(n_hyp, dtype=torch.float64, device=dev)
    d_dist = torch.zeros(n_hyp, dtype=torch.int32, device=dev)
    d_arg = torch.zeros(W * N, dtype=torch.int32, device=dev)
    d_es = torch.zeros(W, dtype=torch.int64, device=dev)
    max_len = int(lens.max())
    P = lambda x: ctypes.c_void_p(x.data_ptr())
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)

    def step_device():
        scorer.score_device(d_tok, off, out=d_pll)
        _lib.check(lib.pllb_levenshtein(P(d_rc), P(d_ro), P(d_hc), P(d_ho), P(d_pr), n_hyp, max_len, P(d_dist), sp))
        _lib.check(lib.pllb_rescore_sweep(P(d_am), P(d_pll), P(d_len), P(d_dist), N, n_best, P(d_w), W, 0,
                                          P(d_arg), P(d_es), sp))
        return 3  # stage-4 launches (levenshtein, memset excluded, sweep) + ... counted below

    # pinned host buffers for `e2e`
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    h_tok, h_am, h_len = pin(tok), pin(am), pin(hyp_len)
    h_pll = torch.zeros(n_hyp, dtype=torch.float64).pin_memory().numpy()

    e2e_parts = [0.0, 0.0, 0.0]

    def step_host():
        t0 = time.perf_counter()
        pll = scorer.score_packed(h_tok, off)
        t1 = time.perf_counter()
        h_pll[:] = pll
        dist_h = engine.levenshtein_packed(rc, ro, hc, ho, pair_ref).reshape(N, n_best)
        t2 = time.perf_counter()
        arg, es = engine.rescore_sweep(h_am, h_pll.reshape(N, n_best), h_len, dist_h, weights, "B")
        t3 = time.perf_counter()
        e2e_parts[0] += t1 - t0; e2e_parts[1] += t2 - t1; e2e_parts[2] += t3 - t2
        return pll, es

    h2d = tok.nbytes + (rc.nbytes + ro.nbytes + hc.nbytes + ho.nbytes + pair_ref.nbytes) + \
        (am.nbytes + n_hyp * 8 + hyp_len.nbytes + n_hyp * 4 + weights.nbytes) + 3 * 4 * (n_hyp + 64)
    d2h = n_hyp * 8 + n_hyp * 4 + W * N * 4 + W * 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------------------------------------------------------- value (device resident)
    for _ in range(args.warmup):
        step_device()
    barrier()
    scorer.reset_stats()
    scorer.set_timing(True)
    sampler = ClockSampler(local)
    sampler.start()
    gemm_ms, gemm_launches = 0.0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_value = e0.elapsed_time(e1) / args.steps
    st = scorer.stats()          # timing of the LAST step's GEMM launches (events reused per call)
    gemm_ms = st["last_gemm_ms"]
    gemm_launches = st["last_gemm_launches"]
    launches = int(st["kernel_launches"]) + 2 * args.steps
    clocks = sampler.stop()
    scorer.set_timing(False)

    # ---------------------------------------------------------------- e2e (host buffers)
    for _ in range(min(args.warmup, 1)):
        step_host()
    barrier()
    e2e_parts[:] = [0.0, 0.0, 0.0]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pll_host, es_host = step_host()
    torch.cuda.synchronize(dev)
    ms_e2e = 1e3 * (time.perf_counter() - t0) / args.steps

    # max over ranks (device-timed), totals over ranks
    stats = torch.tensor([ms_value, ms_e2e, float(n_hyp), gemm_ms], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_value, ms_e2e, total_hyps = float(mx[0]), float(mx[1]), float(sm[2])
        # the tiny data-path exchange: gather per-hypothesis scores to every rank (NCCL)
        full = shard.gather_scores(pll_host, np.arange(n_hyp, dtype=np.int64) + rank * n_hyp, world * n_hyp)
        assert np.isfinite(full).all()
    else:
        total_hyps = float(n_hyp)

    if rank == 0:
        peak_tf, peak_hbm, peak_src = peaks()
        gf = gemm_flops(lens, cfg)                      # SURVEY.md §8(d) formula (all T rows in every layer)
        gf_exec = st["gemm_flops"] / args.steps         # 2*M*N*K summed over the launches of one step
        achieved = gf_exec / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("gemm_dram_bytes_per_launch")
        out = {
            "metric": METRIC, "value": total_hyps / (ms_value / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_value, "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": args.operand_dtype,
            "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "utterances_per_gpu": N, "n_best": n_best,
                       "hyps_per_gpu": n_hyp, "masked_copies_per_gpu": int(lens.sum()),
                       "packed_tokens_per_gpu": int((lens * (lens + 2)).sum()), "weights_grid": W,
                       "chunk_tokens": args.chunk_tokens, "parallelism": f"utterance-sharded x{world}",
                       "l2": "activations per chunk (>= 15 GB) exceed L2; no flush needed",
                       "algorithmic_tflop_per_step_per_gpu": total_flops(lens, cfg) / 1e12},
            "e2e": {"value": total_hyps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e,
                    "ms_parts": {"pllb_score_host": 1e3 * e2e_parts[0] / args.steps,
                                 "pllb_levenshtein_host": 1e3 * e2e_parts[1] / args.steps,
                                 "pllb_rescore_sweep_host": 1e3 * e2e_parts[2] / args.steps}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": achieved, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic,
                         "peak_source": peak_src, "launches_per_step": int(gemm_launches),
                         "gemm_ms_per_step": gemm_ms, "gemm_share_of_step": gemm_ms / ms_value,
                         "executed_gemm_tflop_per_step": gf_exec / 1e12,
                         "survey_formula_gemm_tflop_per_step": gf / 1e12,
                         "ms_by_kind": st["gemm_ms_by_kind"]},
            "whole_step_tflops": total_flops(lens, cfg) / (ms_value / 1e3) / 1e12,
            "pll_checksum": float(np.sum(pll_host)), "best_weight": float(weights[int(np.argmin(es_host))]),
        }
        if world == 1 and not args.no_cpu_baseline:
            v, dt, copies, threads = cpu_port_hyps_per_s(cfg, sd, tok, off, min(args.cpu_sample_hyps, n_hyp))
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                   "sample": f"first {min(args.cpu_sample_hyps, n_hyp)} hypotheses ({copies} masked copies) "
                                             f"of the same workload, oracle port of MLM_PLL/main.py run_one_epoch, "
                                             f"batch 32, fp32, {dt:.1f} s"}
    else:
        out = None
    scorer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        _emit(out)                             # the JSON line is the only thing on stdout


if __name__ == "__main__":
    main()
