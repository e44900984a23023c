#!/usr/bin/env python
"""Writes tests/golden/config_hashes.json: digests (confighash.py) of the CPU oracle's results on the BASELINE
configurations at full size (C1, C2, C3 two-pass, C4 images, C5 16384^2).  Run in the build
container (`python tests/golden/make_config_hashes.py [--only c5]`); the oracle is the C++ restatement of the reference
(oracle/vc_oracle.cpp), itself cross-checked against the independent Python restatement (tests/test_two_oracles.py)."""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from confighash import HASH_FILE, bin_clusters_digest, clusters_digest, components_digest, perimeters_digest  # noqa: E402
from oracle_lib import load_oracle  # noqa: E402
from visioncortex_b200 import synth  # noqa: E402
from visioncortex_b200 import workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--c4-images", type=int, default=64)
    args = ap.parse_args()
    only = set(args.only.split(",")) if args.only else None
    out = json.load(open(HASH_FILE)) if os.path.exists(HASH_FILE) else {}
    o = load_oracle()

    def want(k):
        return only is None or k in only

    if want("c1"):
        mask = W.c1_input()
        for diag in (False, True):
            out[f"c1_diag{int(diag)}"] = bin_clusters_digest(o.to_clusters(mask, diag))
    if want("c2"):
        out["c2"] = clusters_digest(o.runner_run(W.c2_input(), W.c2_config(), pools=False))
    if want("c3"):
        img = W.c3_input()
        r1 = o.runner_run(img, W.c3_config_pass1(), pools=False, color_image=True)
        out["c3_pass1"] = clusters_digest(r1)
        r2 = o.runner_run(r1["color_image"], W.c3_config_pass2(), pools=False, color_image=True)
        out["c3_pass2"] = clusters_digest(r2)
    if want("c4"):
        for k in range(args.c4_images):
            out[f"c4_{k}"] = clusters_digest(o.runner_run(W.c4_input(k), W.c4_config(), pools=False))
    if want("c4_queries"):
        # rows f1 / f4 of SURVEY section 8(f) on C4 image 0: per output cluster the perimeter (cluster.rs:46-48) and the binary
        # clusters of to_image_with_hole(width, true).to_clusters(false) (cluster.rs:108-128), from the oracle's pools
        import numpy as np
        r = o.runner_run(W.c4_input(0), W.c4_config(), pools=True)
        w = r["width"]
        comps, perims = [], np.zeros(r["num_slots"], dtype=np.int64)
        for idx in r["clusters_output"]:
            rec = r["records"][int(idx)]
            left, top, right, bottom = [int(v) for v in rec["rect"]]
            m = np.zeros((bottom - top, right - left), dtype=bool)
            for pool, off, ln, val in ((r["indices_pool"], int(rec["indices_off"]), int(rec["indices_len"]), True),
                                       (r["holes_pool"], int(rec["holes_off"]), int(rec["holes_len"]), False)):
                ii = pool[off:off + ln].astype(np.int64)
                m[ii // w - top, ii % w - left] = val
            bc = o.to_clusters(m, False)
            comps.append([(tuple(int(v) for v in q["rect"]), bc["points"][int(q["points_off"]):int(q["points_off"]) + int(q["points_len"])])
                          for q in bc["records"]])
            pm = np.pad(m, 1)
            c = pm[1:-1, 1:-1]
            perims[int(idx)] = int((c & ~(pm[1:-1, :-2] & pm[1:-1, 2:] & pm[:-2, 1:-1] & pm[2:, 1:-1])).sum())
        out["c4_0_components"] = components_digest(comps)
        out["c4_0_perimeters"] = perimeters_digest(perims, r["clusters_output"])
    if want("c5"):
        t0 = time.time()
        img = W.c5_input()
        print("c5 generated in %.1f s" % (time.time() - t0), flush=True)
        t0 = time.time()
        r = o.runner_run(img, W.c5_config(), pools=False)
        print("c5 oracle in %.1f s, %d slots, %d outputs" % (time.time() - t0, r["num_slots"], len(r["clusters_output"])), flush=True)
        out["c5"] = clusters_digest(r)
    if want("c5_8192"):
        r = o.runner_run(W.c5_input(8192), W.c5_config(), pools=False)
        out["c5_8192"] = clusters_digest(r)
    with open(HASH_FILE, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", HASH_FILE, len(out), "digests")


if __name__ == "__main__":
    main()
