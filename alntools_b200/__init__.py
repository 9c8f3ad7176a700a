"""alntools_b200 — B200-native equivalence-class construction behind alntools bam2ec / bam2emase.

Host side mirrors the reference's converter API (bam_utils.convert, bam_utils_multisample.convert,
methods.*); the grouping / counting / ordering / matrix stages run in hand-written sm_100a kernels
behind the C ABI in include/ecb200.h (libecb200.so).  There is no CPU fallback.
"""
__version__ = "0.1.0"
