"""Per-tensor gradient report of the MLM fine-tuning path against the autograd oracle (GPU box).
    python tools/train_probe.py [tiny|base]        prints loss and, per parameter tensor, rel L2 error + cosine
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asr_rescoring_b200 import engine, synth  # noqa: E402
from oracle import pll_oracle, train_oracle  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    cfg = synth.BERT_TINY if which == "tiny" else dict(synth.BERT_BASE_CHINESE, num_layers=int(os.environ.get("NL", "2")))
    sd = synth.random_init_state_dict(cfg, 10, perturb=True)
    nb = synth.make_nbest(6, 1, seed=3)
    tok, off = nb.packed_tokens(cfg["vocab"])
    rows = train_oracle.training_rows([[int(t) for t in tok[off[i]:off[i + 1]]] for i in range(len(off) - 1)])[3:35]
    ids, am, lab, *_ = pll_oracle.collate(rows)
    ids, am, lab = (x.numpy().astype(np.int32) for x in (ids, am, lab))
    t0 = time.time()
    o_loss, o_grads = train_oracle.loss_and_grads(sd, cfg, rows)
    print(f"oracle: loss {o_loss:.6f} ({time.time() - t0:.1f} s), batch {ids.shape}, pads {(am == 0).sum()}")
    with engine.MlmTrainer(sd, cfg, lr=1e-3, hidden_dropout=0.0, attention_dropout=0.0, max_rows=ids.size, max_seq=ids.shape[1]) as tr:
        l0 = tr.step(ids, am, lab, mode=0)
        l2 = tr.step(ids, am, lab, mode=2)
        g = tr.grads()
        t0 = time.time()
        for _ in range(5):
            tr.step(ids, am, lab, mode=1)
        dt = (time.time() - t0) / 5
        print(f"device: eval loss {l0:.6f}, train loss {l2:.6f}; {dt * 1e3:.2f} ms per full step; {tr.kernel_launches()} launches total")
    for k, og in o_grads.items():
        d, og = g[k].double(), og.double()
        rel = float((d - og).norm() / (og.norm() + 1e-30))
        cos = float((d * og).sum() / (d.norm() * og.norm() + 1e-30))
        flag = "" if (rel <= 0.06 and cos >= 0.998) else "   <<<<"
        print(f"{k:64s} |g| {float(og.norm()):.3e}  dev {float(d.norm()):.3e}  rel {rel:.4f}  cos {cos:.5f}{flag}")


if __name__ == "__main__":
    main()
