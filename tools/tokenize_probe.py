"""Throughput of the text front end on the C2 hypothesis set: device table lookup vs the host
per-sentence tokenizer (the reference's way, MLM_PLL/preprocess.py:10)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from asr_rescoring_b200 import engine, synth
from asr_rescoring_b200.tokenizer import SyntheticCharTokenizer, encode_batch

nb = synth.make_nbest(7176, 10, seed=0)
strings = [h for hs in nb.hyps for h in hs]
tk = SyntheticCharTokenizer()
encode_batch(tk, strings[:100])
t = time.perf_counter(); cp, off = engine.pack_strings(strings); t_pack = time.perf_counter() - t
for _ in range(2):
    t = time.perf_counter(); ids, ooff, flag = engine.tokenize_packed(tk.char_table(), cp, off); t_dev = time.perf_counter() - t
t = time.perf_counter(); ref = [tk.encode(s) for s in strings]; t_host = time.perf_counter() - t
assert ids.tolist() == [x for r in ref for x in r]
print(f"{len(strings)} hyps, {len(cp)} code points: pack_strings {t_pack*1e3:.1f} ms, pllb_tokenize_host {t_dev*1e3:.2f} ms "
      f"({len(strings)/t_dev/1e6:.2f} M hyps/s), host per-sentence encode {t_host*1e3:.1f} ms")
