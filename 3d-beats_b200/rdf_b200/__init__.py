"""rdf_b200: B200-native (sm_100a) per-pixel randomized decision forest hot path of carsonswope/3d-beats.

Public surface = the reference's own (see decision_tree.py, mean_shift.py); kernels live in librdf_b200.so
(C ABI: include/rdf_b200.h).  Importing this package does not load the library; constructing an evaluator,
trainer, forest handle or MeanShift does, and fails loudly if it is missing.
"""
from . import synth  # noqa: F401  (pure NumPy, no GPU needed)

__all__ = ['synth', 'decision_tree', 'mean_shift', 'buffers', 'dist']
