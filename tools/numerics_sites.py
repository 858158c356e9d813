"""Per-rounding-site ablation of the 16-bit operand error (CPU emulation, development tool).

Every GEMM operand of the kernels is rounded to the operand dtype at these sites:
  w      weights of all GEMMs             x_qkv  LayerNorm output feeding QKV
  qkv    Q/K/V as stored for attention    ctx    attention output feeding the output projection
  x_ff1  LayerNorm output feeding FFN1    ffn    GELU output feeding FFN2
  head   transform input + decoder input (their weights are under `w`, or `w_head` when given)
With ONE site rounded to bf16 and the rest exact, the spread of (PLL - fp32 PLL) over hypotheses
gives that site's share of the error variance.
    python tools/numerics_sites.py [n_utts] [base|large]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asr_rescoring_b200 import synth  # noqa: E402
from oracle import pll_oracle  # noqa: E402

SITES = ("w", "x_qkv", "qkv", "ctx", "x_ff1", "ffn", "head")


def rnd(x, dt):
    return x.to(dt).float() if dt is not None else x


def logits(sd, cfg, ids, mpos, dts):
    B, T = ids.shape
    H, NH = cfg["hidden"], cfg["num_heads"]
    dh = H // NH
    eps = 1e-12
    ln = lambda x, p: F.layer_norm(x, (H,), sd[p + ".weight"], sd[p + ".bias"], eps)
    wdt = lambda site: dts.get("w_head", dts.get("w")) if site == "head" else dts.get("w")
    lin = lambda x, p, site: F.linear(rnd(x, dts.get(site)), rnd(sd[p + ".weight"], wdt(site)), sd[p + ".bias"])
    x = (sd["bert.embeddings.word_embeddings.weight"][ids] + sd["bert.embeddings.token_type_embeddings.weight"][0]
         + sd["bert.embeddings.position_embeddings.weight"][torch.arange(T)])
    x = ln(x, "bert.embeddings.LayerNorm")
    for i in range(cfg["num_layers"]):
        p = f"bert.encoder.layer.{i}."
        q, k, v = (rnd(lin(x, p + "attention.self." + n, "x_qkv"), dts.get("qkv")).view(B, T, NH, dh).transpose(1, 2)
                   for n in ("query", "key", "value"))
        s = (q @ k.transpose(-1, -2)) * dh ** -0.5
        ctx = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, T, H)
        x = ln(lin(ctx, p + "attention.output.dense", "ctx") + x, p + "attention.output.LayerNorm")
        f = F.gelu(lin(x, p + "intermediate.dense", "x_ff1"))
        x = ln(lin(f, p + "output.dense", "ffn") + x, p + "output.LayerNorm")
    hm = x[torch.arange(B), mpos]
    t = F.gelu(lin(hm, "cls.predictions.transform.dense", "head"))
    t = F.layer_norm(t, (H,), sd["cls.predictions.transform.LayerNorm.weight"],
                     sd["cls.predictions.transform.LayerNorm.bias"], eps)
    return F.linear(rnd(t, dts.get("head")), rnd(sd["cls.predictions.decoder.weight"], wdt("head")), sd["cls.predictions.bias"])


@torch.no_grad()
def pll(sd, cfg, toks, dts):
    rows = pll_oracle.expand_rows(toks, "u", "h")
    ids = torch.tensor([r["input_ids"] for r in rows])
    mpos = torch.tensor([r["mask_pos"] for r in rows])
    lp = logits(sd, cfg, ids, mpos, dts).log_softmax(-1)[torch.arange(len(toks)), torch.tensor(toks)]
    return float(lp.double().sum())


if __name__ == "__main__":
    n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    large = len(sys.argv) > 2 and sys.argv[2] == "large"
    cfg = synth.BERT_LARGE_SHAPED if large else synth.BERT_BASE_CHINESE
    sd = synth.random_init_state_dict(cfg, 10, False)
    nb = synth.make_nbest(n_utts, 4, seed=123, min_len=8 if large else None, max_len=64 if large else None)
    tok, off = nb.packed_tokens()
    bf, hf = torch.bfloat16, torch.float16
    modes = {"all bf16": {s: bf for s in SITES}, "all fp16": {s: hf for s in SITES}}
    for s in SITES:
        modes[f"only {s} bf16"] = {s: bf}
    modes["fp16 act, bf16 w"] = {**{s: hf for s in SITES}, "w": bf}
    modes["bf16 act, fp16 head"] = {**{s: bf for s in SITES}, "head": hf}
    # not runnable on B200 (kind::f16 rejects A = bf16 with B = fp16: illegal instruction), kept for the record
    modes["(bf16 act x fp16 w, fp16 head)"] = {**{s: bf for s in SITES}, "head": hf, "w": hf}
    modes["bf16 encoder + fp16 head"] = {**{s: bf for s in SITES}, "head": hf, "w_head": hf}     # operand mode 2
    if os.environ.get("SITES_FEW"):
        modes = {k: v for k, v in modes.items() if k in ("all bf16", "all fp16", "bf16 encoder + fp16 head")}
    t0 = time.time()
    Ls, errs = [], {m: [] for m in modes}
    for h in range(len(off) - 1):
        toks = [int(t) for t in tok[off[h]:off[h + 1]]]
        ref = pll(sd, cfg, toks, {})
        Ls.append(len(toks))
        for m, dts in modes.items():
            errs[m].append(pll(sd, cfg, toks, dts) - ref)
    Ls = np.array(Ls)
    print(f"{len(Ls)} hyps ({'large' if large else 'base'}), mean L {Ls.mean():.1f}, {time.time() - t0:.0f} s")
    for m in modes:
        e = np.array(errs[m])
        print(f"{m:30s} rms/sqrt(L) {np.sqrt(np.mean(e * e / Ls)):.5f}  max|e| {np.abs(e).max():.4f}  bias/L {np.mean(e / Ls):+.5f}")
