"""Stand-in for `jiwer` (unpinned by the reference, not installed here): only cer() is used
(rescore.py:8,40,118).  Corpus CER = sum of character Levenshtein distances / sum of reference
lengths over the stripped strings, an empty reference raises — as pinned by the reference's
Nbest_Align/cer.json (tests/test_oracle.py).  Distances come from the oracle's C routine: jiwer's
own backend (rapidfuzz) is compiled code too, so this is the fair stand-in for a CPU baseline."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import oracle  # noqa: E402


def cer(reference, hypothesis) -> float:
    if isinstance(reference, str):
        reference = [reference]
    if isinstance(hypothesis, str):
        hypothesis = [hypothesis]
    if len(reference) != len(hypothesis):
        raise ValueError("reference and hypothesis lists differ in length")
    refs = [r.strip() for r in reference]
    hyps = [h.strip() for h in hypothesis]
    if any(len(r) == 0 for r in refs):
        raise ValueError("one or more references are empty strings")
    return float(int(oracle.levenshtein_strings(refs, hyps).sum())) / float(sum(len(r) for r in refs))
