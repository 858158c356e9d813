"""Experiment: does running two half-size chunks concurrently (two handles, two streams, two
host threads) beat one full-size chunk stream?  Fills the SMs the 45x3 LayerNorm clusters
leave idle with the other stream's kernels."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from asr_rescoring_b200 import engine, synth

cfg = synth.BERT_BASE_CHINESE
sd = synth.random_init_state_dict(cfg, 10)
nb = synth.make_nbest(int(os.environ.get("UTTS", "3000")), 10, seed=0)
hyps = [[synth.synthetic_token_id(c) for c in h] for hs in nb.hyps for h in hs]
def pack(hs):
    off = np.zeros(len(hs) + 1, np.int64); np.cumsum([len(h) for h in hs], out=off[1:])
    return torch.tensor([t for h in hs for t in h], dtype=torch.int32, device="cuda"), off
dev = torch.device("cuda", 0)

def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    return min(ts)

full_tok, full_off = pack(hyps)
for chunk in (1 << 20, 1 << 19):
    sc = engine.PllScorer(sd, cfg, device=0, max_chunk_tokens=chunk)
    ref = sc.score_device(full_tok, full_off); torch.cuda.synchronize()
    t1 = timed(lambda: sc.score_device(full_tok, full_off))
    print(f"one stream, chunk {chunk}: {t1*1e3:.1f} ms  ({len(hyps)/t1:.0f} hyps/s)", flush=True)
    sc.close()

halves = [hyps[0::2], hyps[1::2]]
packs = [pack(h) for h in halves]
for chunk in (1 << 19, 1 << 20):
    scs = [engine.PllScorer(sd, cfg, device=0, max_chunk_tokens=chunk) for _ in range(2)]
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    outs = [None, None]
    def work(i):
        torch.cuda.set_device(0)
        with torch.cuda.stream(streams[i]):
            outs[i] = scs[i].score_device(packs[i][0], packs[i][1])
        streams[i].synchronize()
    def both():
        th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
        [t.start() for t in th]; [t.join() for t in th]
    t2 = timed(both)
    got = torch.empty_like(ref); got[0::2] = outs[0]; got[1::2] = outs[1]
    print(f"two streams, chunk {chunk} each: {t2*1e3:.1f} ms  ({len(hyps)/t2:.0f} hyps/s)  max|d| vs one stream "
          f"{(got-ref).abs().max().item():.2e}", flush=True)
    [s.close() for s in scs]
