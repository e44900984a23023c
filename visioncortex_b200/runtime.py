"""Context / result wrappers over the C ABI (include/vcb200.h).

`Context(device)` opens the CUDA product library (``csrc/libvcb200.so``) on one GPU.  There is no CPU fallback: if the
library is missing or no CUDA device can be opened this raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import BinClustersInfoC, ClustersInfoC, Library, RunnerConfigC, VcbError, pack_bits


def make_config(**kw) -> RunnerConfigC:
    """RunnerConfig::default() (color_clusters/runner.rs:23-39) with overrides."""
    cfg = RunnerConfigC()
    cfg.diagonal = 0
    cfg.hierarchical = _ffi.HIERARCHICAL_MAX
    cfg.batch_size = 25600
    cfg.good_min_area = 16
    cfg.good_max_area = 256 * 256
    cfg.is_same_color_a = 4
    cfg.is_same_color_b = 1
    cfg.deepen_diff = 64
    cfg.hollow_neighbours = 1
    cfg.keying_action = _ffi.KEEP
    for k, v in kw.items():
        if k == "key_color":
            for i in range(4):
                cfg.key_color[i] = int(v[i])
        else:
            if not hasattr(cfg, k):
                raise TypeError(f"unknown RunnerConfig field {k}")
            setattr(cfg, k, v)
    return cfg


class Context:
    def __init__(self, device: int = 0, lib: Library | None = None):
        self.lib = lib or Library()
        self.handle = C.c_void_p()
        st = self.lib.fn("ctx_create")(int(device), C.byref(self.handle))
        if st != 0:
            raise VcbError(st, self.lib.strerror(st) + f": cannot open CUDA device {device} (there is no CPU fallback)")

    def close(self):
        if self.handle:
            self.lib.fn("ctx_destroy")(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- color_clusters::Runner::run
    def runner_run_handle(self, rgba: np.ndarray, cfg: RunnerConfigC):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w = rgba.shape[:2]
        out = C.c_void_p()
        st = self.lib.fn("runner_run")(self.handle, C.byref(cfg), rgba.ctypes.data_as(C.c_void_p), w, h, C.byref(out))
        self.lib.check(st, self.handle)
        return out

    def runner_run(self, rgba: np.ndarray, cfg: RunnerConfigC, pools: bool = True, color_image: bool = False) -> dict:
        handle = self.runner_run_handle(rgba, cfg)
        try:
            res = self.lib.read_clusters(handle, pools=pools, free=False, ctx=self.handle)
            if color_image:
                h, w = rgba.shape[:2]
                img = np.empty((h, w, 4), dtype=np.uint8)
                self.lib.check(self.lib.fn("clusters_to_color_image")(handle, img.ctypes.data_as(C.c_void_p)), self.handle)
                res["color_image"] = img
        finally:
            self.lib.fn("clusters_free")(handle)
        return res

    def runner_run_batch(self, images, cfg: RunnerConfigC, pools: bool = False) -> list:
        imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        n = len(imgs)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        ws = (C.c_uint32 * n)(*[im.shape[1] for im in imgs])
        hs = (C.c_uint32 * n)(*[im.shape[0] for im in imgs])
        outs = (C.c_void_p * n)()
        st = self.lib.fn("runner_run_batch")(self.handle, C.byref(cfg), ptrs, ws, hs, n, outs)
        self.lib.check(st, self.handle)
        res = []
        for k in range(n):
            res.append(self.lib.read_clusters(C.c_void_p(outs[k]), pools=pools, free=True, ctx=self.handle))
        return res

    # ---- BinaryImage::to_clusters
    def to_clusters(self, mask: np.ndarray, diagonal: bool) -> dict:
        mask = np.ascontiguousarray(mask, dtype=bool)
        h, w = mask.shape
        bits = pack_bits(mask)
        out = C.c_void_p()
        st = self.lib.fn("binary_to_clusters")(self.handle, bits.ctypes.data_as(C.c_void_p), w, h, int(diagonal), C.byref(out))
        self.lib.check(st, self.handle)
        return self.lib.read_bin_clusters(out)

    def synchronize(self):
        self.lib.check(self.lib.fn("ctx_synchronize")(self.handle), self.handle)

    def kernel_launches(self) -> int:
        return int(self.lib.fn("ctx_kernel_launches")(self.handle))

    # ---- device-resident batch (bench / GPU-resident callers)
    def runner_run_device(self, d_ptr: int, width: int, height: int, n: int, cfg: RunnerConfigC):
        """d_ptr: device pointer to n RGBA8 images stored back to back.  Returns a list of raw result handles."""
        outs = (C.c_void_p * n)()
        st = self.lib.fn("runner_run_device")(self.handle, C.byref(cfg), C.c_void_p(d_ptr), width, height, n, outs)
        self.lib.check(st, self.handle)
        return [C.c_void_p(outs[k]) for k in range(n)]

    def free_handles(self, handles):
        for h in handles:
            self.lib.fn("clusters_free")(h)

    def profile_reset(self, enable: bool):
        self.lib.check(self.lib.fn("ctx_profile_reset")(self.handle, int(enable)), self.handle)

    def profile_read(self, name: str = ""):
        ms, n = C.c_double(), C.c_uint64()
        self.lib.check(self.lib.fn("ctx_profile_read")(self.handle, name.encode(), C.byref(ms), C.byref(n)), self.handle)
        return ms.value, n.value

    def event_record(self, slot: int):
        self.lib.check(self.lib.fn("ctx_event_record")(self.handle, slot), self.handle)

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_double()
        self.lib.check(self.lib.fn("ctx_event_elapsed_ms")(self.handle, a, b, C.byref(ms)), self.handle)
        return ms.value

    def last_counters(self) -> dict:
        arr = (C.c_uint32 * 8)()
        self.lib.check(self.lib.fn("ctx_last_counters")(self.handle, C.byref(arr)), self.handle)
        return {"candidate_events": arr[0], "new_events": arr[1], "leaves": arr[2], "msf_rounds": arr[3], "err": arr[4],
                "resurrected_leaves": arr[5], "stage1_passes": arr[6], "sequential": arr[7]}


def runner_run_banded(contexts, rgba: np.ndarray, cfg: RunnerConfigC, pools: bool = False, color_image: bool = False) -> dict:
    """One image split into row bands over several GPUs (contexts on distinct devices); result identical to runner_run."""
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = rgba.shape[:2]
    lib = contexts[0].lib
    arr = (C.c_void_p * len(contexts))(*[c.handle.value for c in contexts])
    out = C.c_void_p()
    st = lib.fn("runner_run_banded")(arr, len(contexts), C.byref(cfg), rgba.ctypes.data_as(C.c_void_p), w, h, C.byref(out))
    lib.check(st, contexts[0].handle)
    try:
        res = lib.read_clusters(out, pools=pools, free=False, ctx=contexts[0].handle)
        if color_image:
            img = np.empty((h, w, 4), dtype=np.uint8)
            lib.check(lib.fn("clusters_to_color_image")(out, img.ctypes.data_as(C.c_void_p)), contexts[0].handle)
            res["color_image"] = img
    finally:
        lib.fn("clusters_free")(out)
    return res


def runner_run_batch_host(ctx: Context, images, cfg: RunnerConfigC, records_cap: int | None = None, buffers=None, compact: bool = False) -> list:
    """vcb_runner_run_batch_host: equally sized host images in, host result buffers out, with the copies of neighbouring
    sub-batches overlapping the compute.  `buffers` may hold preallocated (ideally pinned) arrays: a list of
    (labels u32[w*h], records CLUSTER_REC_DTYPE[cap], outputs u32[cap]) per image."""
    from ._ffi import CLUSTER_REC_DTYPE, HostResultC

    imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
    n = len(imgs)
    h, w = imgs[0].shape[:2]
    cap = records_cap if records_cap is not None else w * h + 1
    if buffers is None:
        buffers = [(np.empty(w * h, dtype=np.uint32), np.zeros(cap, dtype=CLUSTER_REC_DTYPE), np.empty(2 * cap, dtype=np.uint32))
                   for _ in range(n)]
    ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
    res = (HostResultC * n)()
    live = [np.empty(len(buffers[k][1]), dtype=np.uint32) for k in range(n)] if compact else None
    for k in range(n):
        lab, rec, out = buffers[k]
        res[k].cluster_indices = lab.ctypes.data
        res[k].records = rec.ctypes.data
        res[k].outputs = out.ctypes.data
        res[k].records_cap = len(rec)
        res[k].outputs_cap = len(out)
        if compact:
            res[k].live_slots = live[k].ctypes.data
    st = ctx.lib.fn("runner_run_batch_host")(ctx.handle, C.byref(cfg), ptrs, w, h, n, res)
    ctx.lib.check(st, ctx.handle)
    out_list = []
    for k in range(n):
        ctx.lib.check(res[k].status, ctx.handle)
        lab, rec, out = buffers[k]
        info = res[k].info
        if res[k].compact:
            # compact records (vcb_host_result.live_slots): every slot that is not listed is Cluster::default()
            full = np.zeros(info.num_slots, dtype=CLUSTER_REC_DTYPE)
            full[live[k][: res[k].num_live]] = rec[: res[k].num_live]
            # indices_off / holes_off are this library's own layout fields (exclusive scans of the list lengths over the slots of
            # the image), not fields of the reference's Cluster: rebuilt here for the default slots as well
            for off, ln in (("indices_off", "indices_len"), ("holes_off", "holes_len")):
                c = np.cumsum(full[ln].astype(np.uint64))
                full[off][1:] = c[:-1]
                full[off][0] = 0
            records = full
        else:
            records = rec[: info.num_slots]
        out_list.append({"width": info.width, "height": info.height, "num_slots": info.num_slots, "cluster_indices": lab,
                         "records": records, "clusters_output": out[: info.num_outputs], "compact": int(res[k].compact),
                         "num_live": int(res[k].num_live)})
    return out_list
