"""B200-native ECoG preprocessing, epoching and channel selection.

Host code is thin Python; the arithmetic runs in ``libecog_sm100.so`` (hand-written
sm_100a CUDA kernels behind a C ABI, see ``include/ecog_sm100.h``).  There is no CPU
fallback: importing an op without the built library raises.
"""
__version__ = "0.1.0"
