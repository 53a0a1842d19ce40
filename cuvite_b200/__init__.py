"""cuvite_b200 — B200-native Louvain local-move engine behind cuVite's GpuGraph interface.

The compute path lives in ``csrc/`` (hand-written sm_100a CUDA behind the C ABI
declared in ``include/cuvite_b200.h``); this package is the Python host-side
mirror of the reference's operator interface plus graph generators and Vite I/O.
"""
__version__ = "0.1.0"
