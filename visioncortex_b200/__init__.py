"""visioncortex_b200 — B200-native (sm_100a CUDA) implementation of visioncortex's clustering hot path.

Host-side mirror of the reference API (`ColorImage`, `BinaryImage`, `color_clusters.Runner`, `RunnerConfig`,
`Clusters`, `ClustersView`) over the C ABI declared in include/vcb200.h.  No CPU fallback exists.
"""
from ._ffi import HIERARCHICAL_MAX, KEEP, DISCARD, VcbError  # noqa: F401

__all__ = ["HIERARCHICAL_MAX", "KEEP", "DISCARD", "VcbError"]
