"""tiatoolbox_b200 — B200-native per-patch hot path behind tiatoolbox's Python seams.

Only the hot path of BASELINE.json's north_star lives here: stain normalisation
(``tools.stainnorm``) and the patch-level CNN engines (``models``), backed by
``libtiab200.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/tiab200.h``).
"""
__version__ = "0.1.0"
