"""Derive the AISHELL-1-test shape table used by synth.make_nbest.

    python tools/make_synth_shape.py        (needs /root/reference; build container only)

Reads the reference's fixtures and keeps only integer shape statistics, in utterance order:
  * ref_len[u]    = len(ref_text[u])                          (espnet_data/alfred/test/ref_text.json)
  * edits[u][k]   = round(hyps_cer[u][hyp_k+1] * ref_len[u])  (espnet_data/alfred/test/hyps_cer.json)
No text and no scores are copied.  Output: asr-rescoring_b200/data/aishell1_test_shape.json
"""
import json
import os

REF = "/root/reference/espnet_data/alfred/test"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "asr-rescoring_b200", "data",
                   "aishell1_test_shape.json")

ref = json.load(open(os.path.join(REF, "ref_text.json"), encoding="utf-8"))
cer = json.load(open(os.path.join(REF, "hyps_cer.json"), encoding="utf-8"))
assert list(ref) == list(cer)
ref_len, edits = [], []
for u, text in ref.items():
    L = len(text)
    ref_len.append(L)
    row = [round(cer[u][f"hyp_{k + 1}"] * L) for k in range(10)]
    for k, d in enumerate(row):
        assert abs(d / L - cer[u][f"hyp_{k + 1}"]) < 1e-9
    edits.append("".join(chr(ord("0") + d) if d < 10 else chr(ord("a") + d - 10) for d in row))
assert sum(ref_len) == 104765 and len(ref_len) == 7176
json.dump({"source": "espnet_data/alfred/test/{ref_text,hyps_cer}.json (lengths and integer edit counts only)",
           "ref_len": ref_len, "edits_base36_per_utt": edits}, open(OUT, "w"), separators=(",", ":"))
print(OUT, os.path.getsize(OUT), "bytes;", sum(ref_len), "chars;",
      sum(int(c, 36) for e in edits for c in e), "edits")
