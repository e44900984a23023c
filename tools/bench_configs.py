#!/usr/bin/env python
"""Times every BASELINE.json configuration that fits one GPU (C1..C4) through the C ABI and prints one JSON line per
config: device-resident Mpix/s, the per-kernel profile and (optionally) the CPU oracle beside it."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from visioncortex_b200 import synth  # noqa: E402
from visioncortex_b200.runtime import Context, make_config  # noqa: E402

HM = 0xFFFFFFFF


def timed(ctx, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    ctx.synchronize()
    ctx.event_record(0)
    for _ in range(steps):
        fn()
    ctx.event_record(1)
    return ctx.event_elapsed_ms(0, 1) / steps


def profile(ctx, fn):
    ctx.profile_reset(True)
    fn()
    ms_all, _ = ctx.profile_read("")
    out = {}
    for name in ["k_edges_tile", "k_flatten_events", "k_boruvka", "k_chain_insert", "k_attribute", "k_tournament", "k_labels",
                 "k_rec_accum", "k_rag_count", "k_rag_fill", "k_s2_collect", "k_s2_epoch", "k_s2_labels", "rs_hist", "rs_scan",
                 "rs_scatter", "k_list_jump", "k_list_keys"]:
        ms, n = ctx.profile_read(name)
        if n:
            out[name] = round(ms, 3)
    ctx.profile_reset(False)
    return ms_all, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch

    ctx = Context(0)
    oracle = None
    if args.cpu:
        from oracle_lib import load_oracle

        oracle = load_oracle()

    def run_colour(name, imgs, cfg, note=""):
        w, h = imgs[0].shape[1], imgs[0].shape[0]
        n = len(imgs)
        dev = torch.from_numpy(np.stack(imgs)).cuda()
        torch.cuda.synchronize()

        def step():
            hs = ctx.runner_run_device(dev.data_ptr(), w, h, n, cfg)
            ctx.free_handles(hs)

        ms = timed(ctx, step, args.steps, args.warmup)
        ms_k, prof = profile(ctx, step)
        line = {"config": name, "images": n, "w": w, "h": h, "ms_per_step": ms, "Mpix_s": n * w * h / ms / 1e3,
                "algo_GBs": 8 * n * w * h / ms / 1e6, "kernel_ms_sum": ms_k, "kernels": prof, "note": note}
        if oracle is not None:
            t0 = time.time()
            oracle.runner_run(imgs[0], cfg, pools=False)
            line["cpu_oracle_Mpix_s_1thread"] = w * h / (time.time() - t0) / 1e6
        print(json.dumps(line), flush=True)

    only = set(args.only.split(",")) if args.only else None
    B = args.batch
    if not only or "C2" in only:
        imgs = [synth.g2_scene(1920, 1080, seed=2 + k) for k in range(min(B, 4))]
        imgs = [imgs[k % len(imgs)] for k in range(B)]
        run_colour("C2 Runner vtracer-default hierarchical=0 1080p", imgs,
                   make_config(hierarchical=0, is_same_color_a=2, deepen_diff=16, good_max_area=1920 * 1080))
    if not only or "C4" in only:
        imgs = [synth.g2_scene(1920, 1080, seed=100 + k) for k in range(min(B, 4))]
        imgs = [imgs[k % len(imgs)] for k in range(B)]
        run_colour("C4 Runner default (hierarchical MAX) 1080p batch", imgs, make_config())
    if not only or "C3" in only:
        key = (40, 200, 120, 255)
        img = synth.g3_posterised(3840, 2160, seed=3, key=key)
        cfg1 = make_config(hierarchical=HM, good_min_area=16, good_max_area=3840 * 2160, is_same_color_a=2, deepen_diff=16,
                           hollow_neighbours=1, key_color=key, keying_action=0)
        run_colour("C3 pass 1 (hierarchical MAX + deepen + key, Keep) 4K", [img], cfg1)
        res = ctx.runner_run(img, cfg1, pools=False, color_image=True)
        cfg2 = make_config(hierarchical=64, good_min_area=0, good_max_area=3840 * 2160, is_same_color_a=0, deepen_diff=0,
                           hollow_neighbours=0, key_color=key, keying_action=1)
        run_colour("C3 pass 2 (hierarchical 64, Discard) 4K", [res["color_image"]], cfg2)
        run_colour("C3 stage 1 only (hierarchical 0) 4K", [img], make_config(hierarchical=0, is_same_color_a=2, good_max_area=3840 * 2160,
                                                                            key_color=key))
    if only and "C5" in only:
        # BASELINE config 5 on ONE GPU (the 8-GPU row-band mode is not built in round 1): 16384x16384 posterised, default config
        img = synth.g3_posterised(16384, 16384, seed=5, cell=256)
        run_colour("C5 single 16384x16384 image, Runner default (hierarchical MAX), one GPU", [img], make_config(),
                   note="single GPU; release-profile wrapping u32 colour sums apply (cluster > 16.8 Mpx)")
    if not only or "C1" in only:
        from visioncortex_b200._ffi import pack_bits

        mask = synth.g1_binary(1024, 1024, 0.70, 1)
        bits = torch.from_numpy(pack_bits(mask).view(np.int32)).cuda()
        lib = ctx.lib
        for diag in (0, 1):
            def step():
                out = C.c_void_p()
                lib.check(lib.fn("binary_to_clusters_device")(ctx.handle, C.c_void_p(bits.data_ptr()), 1024, 1024, diag, C.byref(out)), ctx.handle)
                lib.fn("bin_clusters_free")(out)

            ms = timed(ctx, step, args.steps, args.warmup)
            ms_k, prof = profile(ctx, step)
            nset = int(mask.sum())
            line = {"config": f"C1 to_clusters 1024x1024 p=0.70 diagonal={diag}", "ms_per_step": ms, "Mpix_s": 1024 * 1024 / ms / 1e3,
                    "algo_GBs": (1024 * 1024 / 8 + 8 * nset) / ms / 1e6, "kernel_ms_sum": ms_k, "kernels": prof,
                    "note": "includes D2H of the points list (the result object is host-side)"}
            if oracle is not None:
                t0 = time.time()
                oracle.to_clusters(mask, bool(diag))
                line["cpu_oracle_Mpix_s_1thread"] = 1024 * 1024 / (time.time() - t0) / 1e6
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
