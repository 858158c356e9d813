"""CPU emulation of the kernel's rounding points (operand dtype bf16 vs fp16) to size the
PLL error budget against the fp32 oracle.  Development tool, not part of the product."""
import sys, os, time
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asr_rescoring_b200 import synth
from oracle import pll_oracle

def rnd(x, dt):
    return x.to(dt).float() if dt is not None else x

def emu_logits_masked(sd, cfg, ids, mpos, dt, qkv_dt=None, w_dt=None):
    w_dt = w_dt or dt
    qkv_dt = qkv_dt if qkv_dt is not None else dt
    B, T = ids.shape; H, NH = cfg["hidden"], cfg["num_heads"]; dh = H // NH; eps = 1e-12
    ln = lambda x, p: F.layer_norm(x, (H,), sd[p + ".weight"], sd[p + ".bias"], eps)
    lin = lambda x, p: F.linear(rnd(x, dt), rnd(sd[p + ".weight"], w_dt), sd[p + ".bias"])
    x = sd["bert.embeddings.word_embeddings.weight"][ids] + sd["bert.embeddings.token_type_embeddings.weight"][0] + sd["bert.embeddings.position_embeddings.weight"][torch.arange(T)]
    x = ln(x, "bert.embeddings.LayerNorm")
    for i in range(cfg["num_layers"]):
        p = f"bert.encoder.layer.{i}."
        q = rnd(lin(x, p + "attention.self.query"), qkv_dt).view(B, T, NH, dh).transpose(1, 2)
        k = rnd(lin(x, p + "attention.self.key"), qkv_dt).view(B, T, NH, dh).transpose(1, 2)
        v = rnd(lin(x, p + "attention.self.value"), qkv_dt).view(B, T, NH, dh).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) * dh ** -0.5
        ctx = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, T, H)
        x = ln(lin(ctx, p + "attention.output.dense") + x, p + "attention.output.LayerNorm")
        f = F.gelu(lin(x, p + "intermediate.dense"))
        x = ln(lin(f, p + "output.dense") + x, p + "output.LayerNorm")
    hm = x[torch.arange(B), mpos]
    t = F.gelu(lin(hm, "cls.predictions.transform.dense"))
    t = F.layer_norm(t, (H,), sd["cls.predictions.transform.LayerNorm.weight"], sd["cls.predictions.transform.LayerNorm.bias"], eps)
    return F.linear(rnd(t, dt), rnd(sd["cls.predictions.decoder.weight"], w_dt), sd["cls.predictions.bias"])

@torch.no_grad()
def pll(sd, cfg, toks, dt, **kw):
    L = len(toks)
    rows = pll_oracle.expand_rows(toks, "u", "h")
    ids = torch.tensor([r["input_ids"] for r in rows]); mpos = torch.tensor([r["mask_pos"] for r in rows])
    lg = emu_logits_masked(sd, cfg, ids, mpos, dt, **kw)
    lp = lg.log_softmax(-1)[torch.arange(L), torch.tensor(toks)]
    return float(lp.double().sum())

if __name__ == "__main__":
    n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    cfg = synth.BERT_BASE_CHINESE
    sd = synth.random_init_state_dict(cfg, 10, False)
    nb = synth.make_nbest(n_utts, 4, seed=123)
    tok, off = nb.packed_tokens()
    modes = {"bf16": dict(dt=torch.bfloat16), "fp16": dict(dt=torch.float16),
             "bf16_qkvf32": dict(dt=torch.bfloat16, qkv_dt=torch.float32)}
    t0 = time.time()
    Ls, errs = [], {m: [] for m in modes}
    for h in range(len(off) - 1):
        toks = [int(t) for t in tok[off[h]:off[h + 1]]]
        ref = pll(sd, cfg, toks, None)
        Ls.append(len(toks))
        for m, kw in modes.items():
            errs[m].append(pll(sd, cfg, toks, **kw) - ref)
    Ls = np.array(Ls)
    print(f"{len(Ls)} hyps, {time.time()-t0:.0f}s")
    for m in modes:
        e = np.array(errs[m])
        print(f"{m:12s} mean|e| {np.abs(e).mean():.4f} max|e| {np.abs(e).max():.4f} std {e.std():.4f} bias {e.mean():+.4f}  per-sqrt(L) std {(e/np.sqrt(Ls)).std():.4f}  frac>0.05 {np.mean(np.abs(e)>0.05):.3f}")
