"""asr-rescoring_b200 — B200-native MLM-PLL N-best scoring path.

Drop-in for the ``MLM_PLL/main.py`` (task ``scoring``) + ``rescore.py`` path of
ishine/ASR-Rescoring.  Python is the host; all arithmetic runs in
``libpllb200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/pllb.h``).  There is no CPU fallback: compute calls raise if the
library or a B200 is missing.
"""
__version__ = "0.1.0"
