"""visioncortex_b200 — B200-native (sm_100a CUDA) implementation of visioncortex's clustering hot path.

`visioncortex_b200.api` mirrors the reference's Rust API for this path (`ColorImage`, `BinaryImage`,
`color_clusters::{Runner, RunnerConfig, Builder, IncrementalBuilder, Clusters, ClustersView, Cluster}`) over the C ABI declared
in include/vcb200.h; `visioncortex_b200.runtime` is the thin ctypes layer underneath (what the parity tests and the bench
call).  No CPU fallback exists.
"""
from ._ffi import HIERARCHICAL_MAX, KEEP, DISCARD, VcbError  # noqa: F401

__all__ = ["HIERARCHICAL_MAX", "KEEP", "DISCARD", "VcbError"]
