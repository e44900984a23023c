#!/usr/bin/env python
"""bench.py — Mpix/s of color_clusters::Runner on B200 (BASELINE.json metric), one JSON line on stdout.

Headline workload = BASELINE config 4: a batch of 512 synthetic 1920x1080 RGBA images (G2, seed 100 + k),
`RunnerConfig::default()` (hierarchical MAX), sharded over the N GPUs by contiguous image blocks with no data-path
collective.  The total work is fixed (512 images), so `"scaling": "strong"`.  A "step" is one pass of the whole path
(stage 1, records, hierarchical stage) over a rank's block.
  value  device-resident call (`vcb_runner_run_device`, inputs already in HBM), CUDA events, max over ranks
  e2e    the reference-facing host-buffer call (`vcb_runner_run_batch_host`): H2D of the images and D2H of labels +
         records + clusters_output inside the timed region
The other BASELINE configs (C1 to_clusters, C2 hierarchical off, C3 4K two-pass cutout, C5 16384^2 on one GPU and, with
N > 1, row-banded over the N GPUs) are timed after the headline by rank 0 and reported under "configs", each with a
parity bit: the digest of the result equals the digest of the CPU oracle's result committed in
tests/golden/config_hashes.json (nothing under oracle/ runs for that).
`--impl reference` times the CPU restatement of the reference (oracle/) on all host threads, same config dict.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

W, H = 1920, 1080
NPX = W * H
TOTAL_IMAGES = 512
ALGO_BYTES_PER_PX = 8          # RGBA8 read once + u32 cluster_indices written once
METRIC = "Mpix/s color_clusters::Runner (config 4: 512 x 1920x1080, RunnerConfig::default())"
KERNELS = ["k_meta_init", "k_edges_tile", "k_flatten_events", "k_boruvka", "k_chain_insert", "scan_new", "k_attribute",
           "k_tournament", "scan_label", "k_image_meta", "k_leaf_final", "k_labels", "k_rec_init", "k_slot_img", "k_rec_accum",
           "k_rec_key", "scan_off", "k_img_base", "k_rec_finalize", "k_off_rebase", "k_off_rebase0", "rs_hist", "rs_scan",
           "rs_scatter", "k_s2_init", "k_s2_init_img", "k_rag_count", "scan_deg", "k_rag_fill", "k_s2_collect", "k_s2_epoch",
           "k_s2_labels", "scan_holes", "k_holes_first", "k_holes_rebase", "k_list_init", "k_list_jump", "k_list_keys"]


def workload_config(total_images):
    """identical in both arms (the driver compares the dicts)"""
    return {"workload": "BASELINE config 4: batch of %d 1920x1080 G2 synthetic RGBA images (seed 100 + k), RunnerConfig::default() "
                        "(hierarchical MAX), images sharded over the GPUs, no collective" % total_images,
            "images_total": total_images, "width": W, "height": H}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    def __init__(self, dev):
        self.dev, self.proc, self.lines = dev, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.dev),
                 "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0=None, t1=None):
        """Clocks / throttle reasons of the samples taken inside [t0, t1] (the timed region)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()

        def parse(lines):
            sm, mx, reasons = [], [], set()
            for ln in lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 6:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons

        window = "timed steps"
        inside = [ln for ts, ln in self.lines if t0 is None or (t0 <= ts <= t1 + 0.03)]
        sm, mx, reasons = parse(inside)
        if not sm:
            window = "whole loaded period (timed region shorter than one sampling period)"
            sm, mx, reasons = parse([ln for _, ln in self.lines])
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_oracle_rate(images, cfg, seconds, threads):
    """Mpix/s of the CPU restatement of the reference on `threads` host threads over ~`seconds` of wall time."""
    from oracle_lib import load_oracle

    oracle = load_oracle()
    done = [0] * threads
    stop = time.time() + seconds

    def work(t):
        k = t
        while time.time() < stop or done[t] == 0:
            oracle.runner_run(images[k % len(images)], cfg, pools=False)
            done[t] += 1
            k += threads

    t0 = time.time()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.time() - t0
    return sum(done) * NPX / dt / 1e6, sum(done), dt


def run_reference(args, rank, world):
    """The reference arm: the CPU restatement of the reference (oracle/; no Rust toolchain in this image, so no
    oracle/_ref) on every host thread, C4 images and config.  Rank 0 only."""
    if rank != 0:
        return
    from visioncortex_b200 import workloads

    cfg = workloads.c4_config()
    threads = os.cpu_count() or 1
    images = [workloads.c4_input(k) for k in range(min(8, max(2, threads // 4)))]
    rates = []
    for _ in range(args.warmup):
        cpu_oracle_rate(images, cfg, 1.0, threads)
    t0 = time.time()
    for _ in range(args.steps):
        r, n, dt = cpu_oracle_rate(images, cfg, max(2.0, min(10.0, 60.0 / max(1, args.steps))), threads)
        rates.append(r)
    v = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": (time.time() - t0) * 1e3 / max(1, args.steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic",
            "config": workload_config(args.images),
            "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": threads, "kind": "port",
                             "sample": "each step: all host threads run oracle/ (C++ restatement of the reference, not the Rust binary: no "
                                       "Rust toolchain in this image) on C4 images for a fixed wall-time slice"},
            "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ secondary configs
def timed_ms(ctx, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    ctx.synchronize()
    ctx.event_record(2)
    for _ in range(steps):
        fn()
    ctx.event_record(3)
    return ctx.event_elapsed_ms(2, 3) / steps


def secondary_configs(ctx, local, world, peak, steps=3):
    """C1, C2, C3 and C5 on rank 0's GPU: ms, Mpix/s, whole-step fraction of the HBM peak, parity vs the committed digests."""
    import torch

    from confighash import bin_clusters_digest, clusters_digest, load_hashes
    from visioncortex_b200 import synth, workloads
    from visioncortex_b200._ffi import pack_bits

    hashes = load_hashes()
    dev = torch.device("cuda", local)
    out = {}

    def frac(npx, ms):
        return (ALGO_BYTES_PER_PX * npx / (ms * 1e-3) / 1e9) / peak

    def colour(name, imgs_dev, cfg, hash_key, color_image=False):
        n, h, w = imgs_dev.shape[0], imgs_dev.shape[1], imgs_dev.shape[2]

        def step():
            ctx.free_handles(ctx.runner_run_device(imgs_dev.data_ptr(), w, h, n, cfg))

        ms = timed_ms(ctx, step, steps, 1)
        hs = ctx.runner_run_device(imgs_dev.data_ptr(), w, h, n, cfg)
        res = ctx.lib.read_clusters(hs[0], pools=False, free=False, ctx=ctx.handle)
        if color_image:
            img = np.empty((h, w, 4), dtype=np.uint8)
            ctx.lib.check(ctx.lib.fn("clusters_to_color_image")(hs[0], img.ctypes.data_as(C.c_void_p)), ctx.handle)
            res["color_image"] = img
        ctx.free_handles(hs)
        out[name] = {"images": n, "width": w, "height": h, "ms": ms, "Mpix_s": n * w * h / ms / 1e3,
                     "whole_step_frac": frac(n * w * h, ms), "parity": clusters_digest(res) == hashes.get(hash_key)}
        return res

    # C2: 64 x 1080p, vtracer-default parameters, hierarchical off (round 1's headline); image 0 is G2(seed 2)
    c2 = torch.stack([synth.g2_scene_torch(W, H, 2 + k, dev) for k in range(64)])
    colour("c2", c2, workloads.c2_config(), "c2")
    del c2
    # C3: 4K posterised image with key colour, cutout two-pass
    c3 = synth.g3_posterised_torch(3840, 2160, 3, dev, key=workloads.C3_KEY)[None]
    r1 = colour("c3_pass1", c3, workloads.c3_config_pass1(), "c3_pass1", color_image=True)
    c3b = torch.from_numpy(r1["color_image"]).to(dev)[None]
    colour("c3_pass2", c3b, workloads.c3_config_pass2(), "c3_pass2", color_image=True)
    out["c3"] = {"ms": out["c3_pass1"]["ms"] + out["c3_pass2"]["ms"], "Mpix_s": 3840 * 2160 / (out["c3_pass1"]["ms"] + out["c3_pass2"]["ms"]) / 1e3,
                 "parity": out["c3_pass1"]["parity"] and out["c3_pass2"]["parity"]}
    del c3, c3b
    # C1: BinaryImage::to_clusters on 1024^2 noise
    mask = synth.g1_binary(1024, 1024, 0.70, 1)
    bits = torch.from_numpy(pack_bits(mask).view(np.int32)).to(dev)
    nset = int(mask.sum())
    for diag in (0, 1):
        def step():
            o = C.c_void_p()
            ctx.lib.check(ctx.lib.fn("binary_to_clusters_device")(ctx.handle, C.c_void_p(bits.data_ptr()), 1024, 1024, diag, C.byref(o)), ctx.handle)
            ctx.lib.fn("bin_clusters_free")(o)

        ms = timed_ms(ctx, step, steps, 1)
        res = ctx.to_clusters(mask, bool(diag))
        algo = 1024 * 1024 / 8 + 8 * nset
        out["c1_diag%d" % diag] = {"ms": ms, "Mpix_s": 1024 * 1024 / ms / 1e3, "whole_step_frac": (algo / (ms * 1e-3) / 1e9) / peak,
                                   "parity": bin_clusters_digest(res) == hashes.get("c1_diag%d" % diag),
                                   "note": "includes D2H of the points list (the result object is host-side)"}
    # "next" rows of SURVEY section 8(f) on C4 image 0 (default config): f4 = Cluster::perimeter of every cluster in one device
    # pass, f1 = to_image_with_hole(..).to_clusters(false) of every output cluster as one batched device call (wall time incl. the
    # host-side assembly of the result lists); parity against the committed digests of the oracle's results
    try:
        from confighash import components_digest, perimeters_digest
        from visioncortex_b200 import api

        img0 = synth.g2_scene_torch(W, H, 100, dev).cpu().numpy()
        cl = api.Runner(api.RunnerConfig(), api.ColorImage.from_array(img0), ctx=ctx).run()
        outs = [int(v) for v in cl.view().clusters_output]
        for name, fn, dig in (("f4_perimeters_c4_image", lambda: cl.device_perimeters(), lambda r: perimeters_digest(r, outs)),
                              ("f1_components_c4_image", lambda: cl.device_components(True), components_digest)):
            fn()
            ctx.synchronize()
            t0 = time.time()
            for _ in range(steps):
                r = fn()
            ms = (time.time() - t0) * 1e3 / steps
            out[name] = {"ms_wall": ms, "output_clusters": len(outs), "parity": dig(r) == hashes.get("c4_0_" + name.split("_")[1])}
        del cl
    except Exception as e:      # never lose the headline line over a secondary row
        out["f1_f4_error"] = repr(e)
    # C5: one 16384^2 posterised image, default config, on this GPU
    torch.cuda.empty_cache()
    c5 = synth.g3_posterised_torch(16384, 16384, 5, dev, cell=256)[None]
    colour("c5_single_gpu", c5, workloads.c5_config(), "c5")
    return out, c5


def c5_banded(ctx0, world, c5_dev, peak):
    """C5 over the N GPUs of the box: rank 0 opens a context on every device and calls vcb_runner_run_banded."""
    import torch

    from confighash import clusters_digest, load_hashes
    from visioncortex_b200 import workloads
    from visioncortex_b200.runtime import Context

    host = torch.empty(c5_dev.shape[1:], dtype=torch.uint8).pin_memory()
    host.copy_(c5_dev[0])
    img = host.numpy()
    side = img.shape[0]
    cfg = workloads.c5_config()
    ctxs = [ctx0] + [Context(k) for k in range(1, world)]
    lib = ctx0.lib
    arr = (C.c_void_p * world)(*[c.handle.value for c in ctxs])

    def run():
        o = C.c_void_p()
        lib.check(lib.fn("runner_run_banded")(arr, world, C.byref(cfg), C.c_void_p(img.ctypes.data), side, side, C.byref(o)), ctx0.handle)
        return o

    lib.fn("clusters_free")(run())
    t0 = time.time()
    o = run()
    for c in ctxs:
        c.synchronize()
    ms = (time.time() - t0) * 1e3
    res = lib.read_clusters(o, pools=False, free=True, ctx=ctx0.handle)
    for c in ctxs[1:]:
        c.close()
    return {"gpus": world, "ms": ms, "Mpix_s": side * side / ms / 1e3, "whole_step_frac": (ALGO_BYTES_PER_PX * side * side / (ms * 1e-3) / 1e9) / peak,
            "parity": clusters_digest(res) == load_hashes().get("c5"),
            "note": "wall time of vcb_runner_run_banded incl. the 1.07 GB H2D from pinned memory"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--images", type=int, default=TOTAL_IMAGES, help="1080p images per step over all GPUs (BASELINE config 4: 512)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary configs (C1, C2, C3, C5)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--compact-records", action="store_true",
                    help="end-to-end leg: transfer only the records of the slots that are not Cluster::default() (vcb_host_result.live_slots)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from confighash import clusters_digest, load_hashes
    from visioncortex_b200 import synth, workloads
    from visioncortex_b200._ffi import CLUSTER_REC_DTYPE, HostResultC
    from visioncortex_b200.runtime import Context
    from visioncortex_b200.sharding import shard_range

    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = Context(local)      # raises without the CUDA library / a GPU: there is no CPU fallback
    cfg = workloads.c4_config()
    blk = shard_range(args.images, world, rank)                 # this rank's contiguous block of the batch
    i0, i1 = blk.start, blk.stop
    B = i1 - i0
    # every image of the batch is distinct: G2(1920, 1080, seed = 100 + k), generated on the device
    images = torch.empty((B, H, W, 4), dtype=torch.uint8, device=dev)
    for k in range(B):
        images[k] = synth.g2_scene_torch(W, H, 100 + i0 + k, dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        ctx.free_handles(ctx.runner_run_device(images.data_ptr(), W, H, B, cfg))

    lib = ctx.lib
    # ---- device-resident timing (value)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    l0 = ctx.kernel_launches()
    ctx.event_record(0)
    t0 = time.time()
    for _ in range(args.steps):
        step_device()
    ctx.event_record(1)
    ms_dev = ctx.event_elapsed_ms(0, 1)
    barrier()
    wall = time.time() - t0
    launches = ctx.kernel_launches() - l0
    t_timed = (t0, t0 + wall)
    ms_step = ms_dev / args.steps
    tt = torch.tensor([ms_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item())
    value = args.images * NPX / (ms_step * 1e-3) / 1e6

    # ---- parity bit of the headline: images of this block against the committed digests of the oracle's results
    hashes = load_hashes()
    parity = None
    check = [k for k in (0, B // 2, B - 1) if ("c4_%d" % (i0 + k)) in hashes]
    if check:
        hs = ctx.runner_run_device(images.data_ptr(), W, H, B, cfg)
        parity = all(clusters_digest(lib.read_clusters(hs[k], pools=False, free=False, ctx=ctx.handle)) == hashes["c4_%d" % (i0 + k)] for k in check)
        ctx.free_handles(hs)

    # ---- per-kernel profile of one more step (CUDA events around every launch, on the stream it is launched on)
    ctx.profile_reset(True)
    step_device()
    prof = {}
    for k in KERNELS:
        ms, n = ctx.profile_read(k)
        if n:
            prof[k] = {"ms": ms, "launches": int(n)}
    ctx.profile_reset(False)
    top = max(prof, key=lambda k: prof[k]["ms"]) if prof else None
    peak, peak_src = peaks()
    roof = None
    if top:
        per_launch_ms = prof[top]["ms"] / prof[top]["launches"]
        px_per_launch = B * NPX / prof[top]["launches"]          # every kernel of the pipeline runs once per device batch
        achieved = ALGO_BYTES_PER_PX * px_per_launch / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")     # dram read+write bytes per launch from the committed ncu capture
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("images_per_launch") == int(B / prof[top]["launches"]):
                traffic = tj.get("dram_bytes_per_launch", {}).get(top)
        stage1 = sum(v["ms"] for k, v in prof.items() if not (k.startswith("k_s2") or k.startswith("k_rag") or k in ("scan_deg", "scan_holes", "k_holes_first", "k_holes_rebase", "k_off_rebase", "k_off_rebase0")))
        roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX * px_per_launch, "peak_source": peak_src,
                "kernel_ms_per_launch": per_launch_ms, "launches_per_step": prof[top]["launches"],
                "share_of_step": prof[top]["ms"] / sum(v["ms"] for v in prof.values()),
                "whole_step_frac": (ALGO_BYTES_PER_PX * B * NPX / (ms_step * 1e-3) / 1e9) / peak,
                "stage1_kernel_ms": stage1, "stage2_kernel_ms": sum(v["ms"] for v in prof.values()) - stage1,
                "note": "the hierarchical stage runs on side streams and overlaps the next device batch's stage 1: kernel times add up to more than the step"}

    # ---- end-to-end through the host-buffer C-ABI call: pinned host images in, labels + records + clusters_output out
    e2e = None
    d2h = 0
    if not args.no_e2e:
        REC_CAP = 1 << 16
        host = torch.empty((B, H, W, 4), dtype=torch.uint8).pin_memory()
        host.copy_(images)
        pin_lab = torch.empty((B, NPX), dtype=torch.int32).pin_memory()
        pin_rec = torch.empty((B, REC_CAP * CLUSTER_REC_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
        pin_out = torch.empty((B, REC_CAP), dtype=torch.int32).pin_memory()
        # compact records (vcb_host_result.live_slots): after the hierarchical stage nine of ten slots are Cluster::default(); only
        # the others travel, with their slot indices -- the same information, 1.1 GB less per step (opt-in: measured 7.1-7.6 Gpix/s
        # against 7.2 with every record: the leg is bound by the uploads and the pipeline's lead-in as much as by the downloads)
        compact = args.compact_records
        pin_live = torch.empty((B, REC_CAP), dtype=torch.int32).pin_memory() if compact else None
        e2e_ptrs = (C.c_void_p * B)(*[host[k].data_ptr() for k in range(B)])
        e2e_res = (HostResultC * B)()
        for k in range(B):
            e2e_res[k].cluster_indices = pin_lab[k].data_ptr()
            e2e_res[k].records = pin_rec[k].data_ptr()
            e2e_res[k].outputs = pin_out[k].data_ptr()
            e2e_res[k].records_cap = REC_CAP
            e2e_res[k].outputs_cap = REC_CAP
            if compact:
                e2e_res[k].live_slots = pin_live[k].data_ptr()

        def step_e2e():
            lib.check(lib.fn("runner_run_batch_host")(ctx.handle, C.byref(cfg), e2e_ptrs, W, H, B, e2e_res), ctx.handle)
            n = 0
            for k in range(B):
                lib.check(e2e_res[k].status, ctx.handle)
                nrec = e2e_res[k].num_live if e2e_res[k].compact else e2e_res[k].info.num_slots
                n += NPX * 4 + nrec * (CLUSTER_REC_DTYPE.itemsize + (4 if e2e_res[k].compact else 0)) + e2e_res[k].info.num_outputs * 4
            return n

        e2e_steps = max(2, min(args.steps, 3))
        step_e2e()
        barrier()
        t0e = time.time()
        for _ in range(e2e_steps):
            d2h = step_e2e()
        barrier()
        dt = (time.time() - t0e) / e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": args.images * NPX / float(tt.item()) / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": B * NPX * 4,
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "bytes": "per rank",
               "records": ("compact: the slots that are not Cluster::default(), with their indices (vcb_host_result.live_slots)"
                           if compact else "every slot")}
        # the labels of image 0 of the block and its record count came through the host call: check them against the device-resident run
        hs = ctx.runner_run_device(images.data_ptr(), W, H, min(B, 2), cfg)
        r0 = lib.read_clusters(hs[0], pools=False, free=False, ctx=ctx.handle)
        ctx.free_handles(hs)
        e2e["matches_device_run"] = bool(np.array_equal(pin_lab[0].numpy().view(np.uint32), r0["cluster_indices"]) and
                                         e2e_res[0].info.num_slots == r0["num_slots"] and e2e_res[0].info.num_outputs == len(r0["clusters_output"]))
        del host, pin_lab, pin_rec, pin_out, pin_live
    clocks = sampler.stop(*t_timed)

    cpu = None
    if rank == 0 and args.cpu_seconds > 0:
        sample = [images[k].cpu().numpy() for k in range(min(B, 8))]
        r, n, dtc = cpu_oracle_rate(sample, cfg, args.cpu_seconds, 1)
        cpu = {"value": r, "unit": "Mpix/s", "cores": 1, "kind": "port",
               "sample": f"{n} of the batch's 1080p images through oracle/ (C++ restatement of the reference) in {dtc:.1f} s, 1 thread"}

    # ---- the other BASELINE configs (rank 0; the other ranks release their GPUs first so that C5 can be banded over them)
    configs = None
    counters = ctx.last_counters()
    del images
    torch.cuda.empty_cache()
    if world > 1 and rank != 0:
        ctx.close()
    barrier()
    if rank == 0 and not args.no_configs:
        try:
            configs, c5_dev = secondary_configs(ctx, local, world, peak)
            if world > 1:
                configs["c5_banded"] = c5_banded(ctx, world, c5_dev, peak)
        except Exception as e:      # a failing secondary config must not lose the headline line
            configs = dict(configs or {}, error=repr(e))
    barrier()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/u32 integer",
                "data": "synthetic", "config": workload_config(args.images),
                "run": {"images_per_step_per_gpu": B, "distinct_images": B, "parity_vs_oracle_digests": parity,
                        "l2": "inputs larger than L2 (%.0f MB per rank per step)" % (B * NPX * 4 / 1e6),
                        "sharding": "contiguous image blocks per GPU, no collective"},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "graph_counters_last_batch": counters, "kernel_profile_ms": {k: round(v["ms"], 4) for k, v in prof.items()},
                "configs": configs, "wall_s_timed": wall}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
