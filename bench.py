#!/usr/bin/env python
"""bench.py — Mpix/s of color_clusters::Runner on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" is one pass of the hot path over one batch of synthetic 1920x1080 RGBA images (BASELINE config 2: vtracer-default
parameters, hierarchical off).  `value` times the device-resident call (inputs already in HBM), `e2e` times the
reference-facing C-ABI call with host buffers (H2D of the images and D2H of labels + records inside the timed region).
`--impl reference` times the CPU restatement of the reference (oracle/) on the host cores instead.
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

W, H = 1920, 1080
NPX = W * H
ALGO_BYTES_PER_PX = 8          # RGBA8 read once + u32 cluster_indices written once
KERNELS = ["k_meta_init", "k_edges_tile", "k_flatten_events", "k_boruvka", "k_chain_insert", "scan_new", "k_attribute",
           "k_tournament", "scan_label", "k_image_meta", "k_leaf_final", "k_labels", "k_rec_init",
           "k_slot_img", "k_rec_accum", "k_rec_key", "k_rec_final", "scan_off", "k_img_base", "k_rec_finalize", "k_off_rebase", "k_off_rebase0",
           "k_out_keys", "rs_hist", "rs_scan", "rs_scatter"]


def c2_config():
    from visioncortex_b200.runtime import make_config

    return make_config(diagonal=0, hierarchical=0, batch_size=25600, good_min_area=16, good_max_area=NPX,
                       is_same_color_a=2, is_same_color_b=1, deepen_diff=16, hollow_neighbours=1)


def make_images(n_distinct, rank):
    from visioncortex_b200 import synth

    return [synth.g2_scene(W, H, seed=2 + 1000 * rank + k) for k in range(n_distinct)]


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
        """Clocks / throttle reasons of the samples taken inside [t0, t1] (the timed region).  nvidia-smi needs a few hundred
        ms to start, so the sampler runs from before the warm-up; if the timed region was too short to catch a sample, the
        samples of the whole loaded period (warm-up, timed steps, end-to-end leg) are reported and `window` says so."""
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
            window = "warm-up + timed steps + end-to-end leg (timed region shorter than one sampling period)"
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
    if rank != 0:
        return
    cfg = c2_config()
    images = make_images(4, 0)
    threads = os.cpu_count() or 1
    rates = []
    for _ in range(args.warmup):
        cpu_oracle_rate(images, cfg, 1.0, threads)
    t0 = time.time()
    for _ in range(args.steps):
        r, n, dt = cpu_oracle_rate(images, cfg, max(2.0, min(10.0, 60.0 / max(1, args.steps))), threads)
        rates.append(r)
    v = float(np.mean(rates))
    line = {"impl": "reference", "metric": "Mpix/s color_clusters::Runner (config 2, 1920x1080, hierarchical off)", "value": v,
            "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": (time.time() - t0) * 1e3 / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic",
            "config": {"workload": "BASELINE config 2: Runner vtracer-default, hierarchical off, 1920x1080 G2 synthetic RGBA",
                       "note": "C++ restatement of the reference algorithm (oracle/), not the Rust binary: no Rust toolchain in this image"},
            "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": threads, "kind": "port",
                             "sample": "each step: all host threads run the oracle on 1080p images for a fixed wall-time slice"},
            "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=64, help="1080p images per step per GPU")
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic images generated per GPU (cycled to fill the batch)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from visioncortex_b200.runtime import Context

    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    ctx = Context(local)      # raises without the CUDA library / a GPU: there is no CPU fallback
    cfg = c2_config()
    B = args.batch
    distinct = make_images(min(args.distinct, B), rank)
    host = torch.empty((B, H, W, 4), dtype=torch.uint8).pin_memory()
    for k in range(B):
        host[k] = torch.from_numpy(distinct[k % len(distinct)])
    dev = host.to(f"cuda:{local}", non_blocking=False)
    torch.cuda.synchronize()
    host_np = [host[k].numpy() for k in range(B)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        hs = ctx.runner_run_device(dev.data_ptr(), W, H, B, cfg)
        ctx.free_handles(hs)

    lib = ctx.lib
    from visioncortex_b200._ffi import CLUSTER_REC_DTYPE, HostResultC

    # end-to-end leg: host images in pinned memory -> vcb_runner_run_batch_host -> labels, records and clusters_output of
    # every image in pinned host memory.  The call pipelines H2D / compute / D2H over sub-batches.
    REC_CAP = NPX // 8
    pin_lab = torch.empty((B, NPX), dtype=torch.int32).pin_memory()
    pin_rec = torch.empty((B, REC_CAP * CLUSTER_REC_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
    pin_out = torch.empty((B, REC_CAP), dtype=torch.int32).pin_memory()
    e2e_ptrs = (C.c_void_p * B)(*[im.ctypes.data for im in host_np])
    e2e_res = (HostResultC * B)()
    for k in range(B):
        e2e_res[k].cluster_indices = pin_lab[k].data_ptr()
        e2e_res[k].records = pin_rec[k].data_ptr()
        e2e_res[k].outputs = pin_out[k].data_ptr()
        e2e_res[k].records_cap = REC_CAP
        e2e_res[k].outputs_cap = REC_CAP

    def step_e2e():
        lib.check(lib.fn("runner_run_batch_host")(ctx.handle, C.byref(cfg), e2e_ptrs, W, H, B, e2e_res), ctx.handle)
        d2h = 0
        for k in range(B):
            lib.check(e2e_res[k].status, ctx.handle)
            d2h += NPX * 4 + e2e_res[k].info.num_slots * CLUSTER_REC_DTYPE.itemsize + e2e_res[k].info.num_outputs * 4
        return d2h

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
    tt = torch.tensor([ms_step], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item())
    value = world * B * NPX / (ms_step * 1e-3) / 1e6

    # ---- per-kernel profile of one more step (CUDA events around every launch on the launching stream)
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
        achieved = ALGO_BYTES_PER_PX * B * NPX / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")     # dram read+write bytes per launch from the committed ncu capture
        if os.path.exists(tpath) and B == 64:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch", {}).get(top)
        roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX * B * NPX, "peak_source": peak_src, "kernel_ms_per_launch": per_launch_ms,
                "share_of_step": prof[top]["ms"] / sum(v["ms"] for v in prof.values()),
                "whole_step_frac": (ALGO_BYTES_PER_PX * B * NPX / (ms_step * 1e-3) / 1e9) / peak}

    # ---- end-to-end through the host-buffer C-ABI call
    e2e_steps = max(2, min(args.steps, 5))
    step_e2e()
    barrier()
    t0 = time.time()
    d2h = 0
    for _ in range(e2e_steps):
        d2h = step_e2e()
    barrier()
    dt = (time.time() - t0) / e2e_steps
    tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e = world * B * NPX / float(tt.item()) / 1e6
    clocks = sampler.stop(*t_timed)

    cpu = None
    if rank == 0 and world == 1 and args.cpu_seconds > 0:
        r, n, dtc = cpu_oracle_rate(distinct, cfg, args.cpu_seconds, 1)
        cpu = {"value": r, "unit": "Mpix/s", "cores": 1, "kind": "port",
               "sample": f"{n} of the batch's 1080p images through oracle/ (C++ restatement of the reference) in {dtc:.1f} s, 1 thread"}

    if rank == 0:
        line = {"metric": "Mpix/s color_clusters::Runner (config 2, 1920x1080, hierarchical off)", "value": value, "unit": "Mpix/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic",
                "config": {"workload": "BASELINE config 2: Runner vtracer-default, hierarchical off, 1920x1080 G2 synthetic RGBA",
                           "images_per_step_per_gpu": B, "distinct_images": len(distinct),
                           "l2": "inputs larger than L2 (%.0f MB per step)" % (B * NPX * 4 / 1e6),
                           "sharding": "independent images per GPU, no collective"},
                "roofline": roof, "cpu_baseline": cpu,
                "e2e": {"value": e2e, "unit": "Mpix/s", "h2d_bytes_per_step": B * NPX * 4, "d2h_bytes_per_step": int(d2h),
                        "steps": e2e_steps},
                "gpu_launches": int(launches), "clocks": clocks, "graph_counters_per_step": ctx.last_counters(), "kernel_profile_ms": {k: round(v["ms"], 4) for k, v in prof.items()},
                "wall_s_timed": wall}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
