"""Loader + thin helpers for the CPU oracle (TEST INFRASTRUCTURE).  Never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from visioncortex_b200._ffi import Library, RunnerConfigC, VcbError, pack_bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


def _cpu_tag() -> str:
    """-march=native output is only valid on the CPU model it was built on: key the library by the host's flags."""
    import hashlib

    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags"))
    except Exception:
        flags = "unknown"
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


ORACLE_SO = os.path.join(ORACLE_DIR, f"liboracle-{_cpu_tag()}.so")


def build_oracle():
    src = os.path.join(ORACLE_DIR, "vc_oracle.cpp")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", f"OUT={os.path.basename(ORACLE_SO)}"])
    return ORACLE_SO


class Oracle:
    def __init__(self):
        self.lib = Library(build_oracle(), prefix="vco_", has_ctx=False)
        self.lib.fn("clusters_perimeter").argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]

    def runner_run(self, rgba: np.ndarray, cfg: RunnerConfigC, pools=True, color_image=False) -> dict:
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w = rgba.shape[:2]
        handle = C.c_void_p()
        st = self.lib.fn("runner_run")(C.byref(cfg), rgba.ctypes.data_as(C.c_void_p), w, h, C.byref(handle))
        self.lib.check(st)
        res = self.lib.read_clusters(handle, pools=pools, free=False)
        if color_image:
            img = np.empty((h, w, 4), dtype=np.uint8)
            self.lib.check(self.lib.fn("clusters_to_color_image")(handle, img.ctypes.data_as(C.c_void_p)))
            res["color_image"] = img
        self.lib.fn("clusters_free")(handle)
        return res

    def to_clusters(self, mask: np.ndarray, diagonal: bool) -> dict:
        mask = np.ascontiguousarray(mask, dtype=bool)
        h, w = mask.shape
        bits = pack_bits(mask)
        handle = C.c_void_p()
        st = self.lib.fn("binary_to_clusters")(bits.ctypes.data_as(C.c_void_p), w, h, int(diagonal), C.byref(handle))
        self.lib.check(st)
        return self.lib.read_bin_clusters(handle)


_ORACLE = None


def load_oracle() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE
