#!/usr/bin/env python
"""Pinned host<->device copy rates of this box, alone and both directions at once, on 1..N GPUs concurrently (one process, one
pair of streams per GPU): the ceiling of the end-to-end leg of bench.py, which moves 4.25 GB up and 5.5 GB down per 512-image
step.  usage: pcie_probe.py [max_gpus]"""
import json
import sys
import time

import torch


def rate(fn, nbytes, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(reps):
        fn()
    for d in range(torch.cuda.device_count()):
        torch.cuda.synchronize(d)
    return nbytes * reps / (time.time() - t0) / 1e9


def main():
    maxg = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    size = 1 << 30
    out = {}
    for ng in [g for g in (1, 2, 4, 8) if g <= min(maxg, torch.cuda.device_count())]:
        hs_up = [torch.empty(size, dtype=torch.uint8).pin_memory() for _ in range(ng)]
        hs_dn = [torch.empty(size, dtype=torch.uint8).pin_memory() for _ in range(ng)]
        ds_up = [torch.empty(size, dtype=torch.uint8, device=f"cuda:{g}") for g in range(ng)]
        ds_dn = [torch.empty(size, dtype=torch.uint8, device=f"cuda:{g}") for g in range(ng)]
        s_up = [torch.cuda.Stream(device=g) for g in range(ng)]
        s_dn = [torch.cuda.Stream(device=g) for g in range(ng)]

        def up():
            for g in range(ng):
                with torch.cuda.stream(s_up[g]):
                    ds_up[g].copy_(hs_up[g], non_blocking=True)

        def dn():
            for g in range(ng):
                with torch.cuda.stream(s_dn[g]):
                    hs_dn[g].copy_(ds_dn[g], non_blocking=True)

        def both():
            up()
            dn()

        r = {"h2d_GBs": rate(up, size * ng), "d2h_GBs": rate(dn, size * ng), "both_GBs_each_way": rate(both, size * ng)}
        # the bench step: 4.25 GB up + 5.5 GB down per 512 images, both directions busy
        r["floor_ms_per_512_images_at_both_rate"] = 5.515e9 / (r["both_GBs_each_way"] / ng * 1e9) * 1e3 / 1.0 if ng == 1 else None
        out[f"gpus_{ng}"] = r
        del hs_up, hs_dn, ds_up, ds_dn
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
