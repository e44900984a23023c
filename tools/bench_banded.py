#!/usr/bin/env python
"""BASELINE config 5: one 16384x16384 posterised image, RunnerConfig::default(), row bands over N GPUs of one box."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ctypes as C
from visioncortex_b200 import synth
from visioncortex_b200.runtime import Context, make_config

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=8)
ap.add_argument("--side", type=int, default=16384)
ap.add_argument("--steps", type=int, default=2)
args = ap.parse_args()
import torch
host = torch.from_numpy(synth.g3_posterised(args.side, args.side, seed=5, cell=256)).pin_memory()
img = host.numpy()
cfg = make_config()
for n in sorted({1, args.gpus}):
    ctxs = [Context(k) for k in range(n)]
    lib = ctxs[0].lib
    arr = (C.c_void_p * n)(*[c.handle.value for c in ctxs])
    def step():
        out = C.c_void_p()
        lib.check(lib.fn("runner_run_banded")(arr, n, C.byref(cfg), C.c_void_p(img.ctypes.data), args.side, args.side, C.byref(out)), ctxs[0].handle)
        lib.fn("clusters_free")(out)
    step()
    ctxs[0].profile_reset(True)
    t0 = time.time()
    for _ in range(args.steps):
        step()
    for c in ctxs:
        c.synchronize()
    dt = (time.time() - t0) / args.steps
    prof = {}
    for k in ["k_edges_tile", "k_flatten_events", "k_attribute", "k_labels", "k_boruvka", "k_tournament", "k_s2_epoch"]:
        ms, cnt = ctxs[0].profile_read(k)
        if cnt: prof[k] = round(ms / args.steps, 3)
    ctxs[0].profile_reset(False)
    print(json.dumps({"config": f"C5 {args.side}^2 default config, row bands over {n} GPU(s), host image in pinned memory (H2D included)",
                      "gpus": n, "ms_per_step_wall": dt * 1e3, "Mpix_s": args.side * args.side / dt / 1e6, "device0_kernel_ms": prof}), flush=True)
    for c in ctxs:
        c.close()
