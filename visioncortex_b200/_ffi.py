"""ctypes declarations for the C ABI in include/vcb200.h.

The product library is ``visioncortex_b200/csrc/libvcb200.so`` (built by ``__graft_entry__.build()``).  There is no
CPU fallback: if the library is missing or no CUDA device can be opened, loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HIERARCHICAL_MAX = 0xFFFFFFFF
KEEP, DISCARD = 0, 1

VCB_OK = 0
VCB_ERR_INVALID_ARG = 1
VCB_ERR_LABEL_OVERFLOW = 2
VCB_ERR_UNSUPPORTED = 3
VCB_ERR_CUDA = 4
VCB_ERR_NCCL = 5
VCB_ERR_NOMEM = 6


class RunnerConfigC(C.Structure):
    _fields_ = [
        ("diagonal", C.c_int32),
        ("hierarchical", C.c_uint32),
        ("batch_size", C.c_int32),
        ("good_min_area", C.c_uint64),
        ("good_max_area", C.c_uint64),
        ("is_same_color_a", C.c_int32),
        ("is_same_color_b", C.c_int32),
        ("deepen_diff", C.c_int32),
        ("hollow_neighbours", C.c_uint64),
        ("key_color", C.c_uint8 * 4),
        ("keying_action", C.c_int32),
    ]


class ClustersInfoC(C.Structure):
    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("num_slots", C.c_uint32),
        ("num_outputs", C.c_uint32),
        ("indices_total", C.c_uint64),
        ("holes_total", C.c_uint64),
    ]


class HostResultC(C.Structure):
    _fields_ = [
        ("cluster_indices", C.c_void_p),
        ("records", C.c_void_p),
        ("outputs", C.c_void_p),
        ("records_cap", C.c_uint32),
        ("outputs_cap", C.c_uint32),
        ("info", ClustersInfoC),
        ("status", C.c_int),
        ("live_slots", C.c_void_p),
        ("num_live", C.c_uint32),
        ("compact", C.c_uint32),
    ]


class BinClustersInfoC(C.Structure):
    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("num_clusters", C.c_uint32),
        ("_pad", C.c_uint32),
        ("points_total", C.c_uint64),
        ("rect", C.c_int32 * 4),
    ]


# numpy mirrors of vcb_cluster_rec / vcb_bin_cluster_rec (C layout, natural alignment)
CLUSTER_REC_DTYPE = np.dtype(
    [
        ("sum", "<u4", (5,)),
        ("residue_sum", "<u4", (5,)),
        ("rect", "<i4", (4,)),
        ("num_holes", "<u4"),
        ("depth", "<u4"),
        ("merged_into", "<u4"),
        ("indices_len", "<u4"),
        ("indices_off", "<u8"),
        ("holes_off", "<u8"),
        ("holes_len", "<u4"),
        ("_pad", "<u4"),
    ],
    align=True,
)
BIN_CLUSTER_REC_DTYPE = np.dtype(
    [("rect", "<i4", (4,)), ("points_off", "<u8"), ("points_len", "<u4"), ("_pad", "<u4")], align=True
)
assert CLUSTER_REC_DTYPE.itemsize == 96, CLUSTER_REC_DTYPE.itemsize
assert BIN_CLUSTER_REC_DTYPE.itemsize == 32, BIN_CLUSTER_REC_DTYPE.itemsize

DEFAULT_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libvcb200.so")

# every symbol include/vcb200.h declares (tests assert the library exports all of them)
EXPORTED_SYMBOLS = [
    "vcb_strerror", "vcb_last_error", "vcb_abi_version", "vcb_runner_config_default",
    "vcb_ctx_create", "vcb_ctx_destroy", "vcb_ctx_device", "vcb_ctx_set_option", "vcb_clusters_take_image",
    "vcb_runner_run", "vcb_runner_run_batch", "vcb_runner_run_device", "vcb_runner_run_banded", "vcb_runner_run_batch_host",
    "vcb_clusters_info_get", "vcb_clusters_cluster_indices", "vcb_clusters_outputs", "vcb_clusters_records",
    "vcb_clusters_indices_pool", "vcb_clusters_holes_pool", "vcb_clusters_to_color_image",
    "vcb_clusters_device_labels", "vcb_clusters_free",
    "vcb_clusters_perimeters", "vcb_clusters_neighbours", "vcb_clusters_to_image",
    "vcb_clusters_components", "vcb_components_info", "vcb_components_get", "vcb_components_free",
    "vcb_binary_to_clusters", "vcb_binary_to_clusters_device", "vcb_bin_clusters_info_get",
    "vcb_bin_clusters_records", "vcb_bin_clusters_points", "vcb_bin_clusters_free",
    "vcb_ctx_synchronize", "vcb_ctx_kernel_launches", "vcb_ctx_profile_reset", "vcb_ctx_profile_read",
    "vcb_ctx_event_record", "vcb_ctx_event_elapsed_ms", "vcb_ctx_last_counters",
]


class VcbError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"vcb status {status}: {msg}")
        self.status = status


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Library:
    """A loaded C-ABI library.  ``prefix`` is ``vcb_`` for the CUDA product library."""

    def __init__(self, path: str = DEFAULT_LIB, prefix: str = "vcb_", has_ctx: bool = True):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)"
            )
        self.path = path
        self.prefix = prefix
        self.has_ctx = has_ctx
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        self._declare()

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _declare(self):
        p = C.c_void_p
        pp = C.POINTER(C.c_void_p)
        ctx = [p] if self.has_ctx else []
        f = self.fn
        f("runner_run").argtypes = ctx + [C.POINTER(RunnerConfigC), p, C.c_uint32, C.c_uint32, pp]
        f("clusters_info_get").argtypes = [p, C.POINTER(ClustersInfoC)]
        for n in ("cluster_indices", "outputs", "records", "indices_pool", "holes_pool", "to_color_image"):
            f("clusters_" + n).argtypes = [p, p]
            f("clusters_" + n).restype = C.c_int
        f("clusters_free").argtypes = [p]
        f("clusters_free").restype = None
        if self.has_ctx:                # product library only (the oracle exposes the host-side result shapes)
            f("clusters_perimeters").argtypes = [p, p]
            f("clusters_neighbours").argtypes = [p, C.c_uint32, p, C.c_size_t, C.POINTER(C.c_size_t)]
            f("clusters_to_image").argtypes = [p, C.c_uint32, C.c_int, p, C.c_size_t, C.POINTER(C.c_int32 * 4)]
            f("clusters_components").argtypes = [p, C.c_int, pp]
            f("components_info").argtypes = [p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
            f("components_get").argtypes = [p, p, p, p]
            f("components_free").argtypes = [p]
            f("components_free").restype = None
        f("binary_to_clusters").argtypes = ctx + [p, C.c_uint32, C.c_uint32, C.c_int, pp]
        f("bin_clusters_info_get").argtypes = [p, C.POINTER(BinClustersInfoC)]
        f("bin_clusters_records").argtypes = [p, p]
        f("bin_clusters_points").argtypes = [p, p]
        f("bin_clusters_free").argtypes = [p]
        f("bin_clusters_free").restype = None
        if self.has_ctx:
            f("strerror").argtypes = [C.c_int]
            f("strerror").restype = C.c_char_p
            f("last_error").argtypes = [p]
            f("last_error").restype = C.c_char_p
            f("abi_version").restype = C.c_uint32
            f("runner_config_default").argtypes = [C.POINTER(RunnerConfigC)]
            f("runner_config_default").restype = None
            f("ctx_create").argtypes = [C.c_int, pp]
            f("ctx_destroy").argtypes = [p]
            f("ctx_destroy").restype = None
            f("ctx_device").argtypes = [p]
            f("ctx_set_option").argtypes = [p, C.c_int, C.c_int64]
            f("clusters_take_image").argtypes = [p, p]
            f("runner_run_batch").argtypes = [p, C.POINTER(RunnerConfigC), p, p, p, C.c_size_t, p]
            f("runner_run_batch_host").argtypes = [p, C.POINTER(RunnerConfigC), p, C.c_uint32, C.c_uint32, C.c_size_t, p]
            f("runner_run_banded").argtypes = [p, C.c_int, C.POINTER(RunnerConfigC), p, C.c_uint32, C.c_uint32, pp]
            f("runner_run_device").argtypes = [p, C.POINTER(RunnerConfigC), p, C.c_uint32, C.c_uint32, C.c_size_t, p]
            f("clusters_device_labels").argtypes = [p]
            f("clusters_device_labels").restype = p
            f("binary_to_clusters_device").argtypes = [p, p, C.c_uint32, C.c_uint32, C.c_int, pp]
            f("ctx_synchronize").argtypes = [p]
            f("ctx_kernel_launches").argtypes = [p]
            f("ctx_kernel_launches").restype = C.c_uint64
            f("ctx_profile_reset").argtypes = [p, C.c_int]
            f("ctx_profile_read").argtypes = [p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
            f("ctx_last_counters").argtypes = [p, C.POINTER(C.c_uint32 * 8)]
            f("ctx_event_record").argtypes = [p, C.c_int]
            f("ctx_event_elapsed_ms").argtypes = [p, C.c_int, C.c_int, C.POINTER(C.c_double)]

    # ------------------------------------------------------------------ helpers
    def strerror(self, status: int) -> str:
        if self.has_ctx:
            return self.fn("strerror")(status).decode()
        return {1: "invalid argument", 2: "label overflow", 3: "unsupported", 6: "out of memory"}.get(status, "?")

    def check(self, status: int, ctx=None):
        if status != VCB_OK:
            msg = self.strerror(status)
            if ctx is not None and self.has_ctx:
                detail = self.fn("last_error")(ctx)
                if detail:
                    msg += ": " + detail.decode()
            raise VcbError(status, msg)

    # ------------------------------------------------------------------ result extraction (shared by both libs)
    def read_clusters(self, handle, *, pools: bool = True, free: bool = True, ctx=None) -> dict:
        f = self.fn
        info = ClustersInfoC()
        self.check(f("clusters_info_get")(handle, C.byref(info)), ctx)
        n = info.width * info.height
        res = {
            "width": info.width,
            "height": info.height,
            "num_slots": info.num_slots,
            "cluster_indices": np.empty(n, dtype=np.uint32),
            "clusters_output": np.empty(info.num_outputs, dtype=np.uint32),
            "records": np.zeros(info.num_slots, dtype=CLUSTER_REC_DTYPE),
        }
        self.check(f("clusters_cluster_indices")(handle, _ptr(res["cluster_indices"])), ctx)
        self.check(f("clusters_outputs")(handle, _ptr(res["clusters_output"])), ctx)
        self.check(f("clusters_records")(handle, _ptr(res["records"])), ctx)
        if pools:
            res["indices_pool"] = np.empty(info.indices_total, dtype=np.uint32)
            res["holes_pool"] = np.empty(info.holes_total, dtype=np.uint32)
            self.check(f("clusters_indices_pool")(handle, _ptr(res["indices_pool"])), ctx)
            self.check(f("clusters_holes_pool")(handle, _ptr(res["holes_pool"])), ctx)
        if free:
            f("clusters_free")(handle)
        return res

    def read_bin_clusters(self, handle, *, free: bool = True) -> dict:
        f = self.fn
        info = BinClustersInfoC()
        self.check(f("bin_clusters_info_get")(handle, C.byref(info)))
        res = {
            "width": info.width,
            "height": info.height,
            "rect": np.array(list(info.rect), dtype=np.int32),
            "records": np.zeros(info.num_clusters, dtype=BIN_CLUSTER_REC_DTYPE),
            "points": np.empty((info.points_total, 2), dtype=np.int32),
        }
        self.check(f("bin_clusters_records")(handle, _ptr(res["records"])))
        self.check(f("bin_clusters_points")(handle, _ptr(res["points"])))
        if free:
            f("bin_clusters_free")(handle)
        return res


def pack_bits(mask: np.ndarray) -> np.ndarray:
    """bool image (h, w) -> bit_vec::BitVec storage: u32 blocks, pixel i at block i/32, bit i%32 (LSB first)."""
    flat = np.ascontiguousarray(mask, dtype=bool).reshape(-1)
    nblocks = (flat.size + 31) // 32
    padded = np.zeros(nblocks * 32, dtype=np.uint8)
    padded[: flat.size] = flat
    return np.packbits(padded.reshape(-1, 8), axis=1, bitorder="little").reshape(-1).view("<u4").copy()
