#!/usr/bin/env python
"""Time-boxed differential fuzz of the CUDA path against the CPU oracle (test infrastructure; run on a GPU box):
random small / medium images over every configuration family -- hierarchical levels, keys clean and unclean, both keying
actions, diagonal, deepen / hollow / good-area parameters, single runs and batches -- comparing labels, every record field,
clusters_output, the ordered indices / holes pools and to_color_image.  usage: fuzz_gpu.py [seconds] [seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from oracle_lib import load_oracle  # noqa: E402
from parity import assert_same_bin_clusters, assert_same_clusters, random_case  # noqa: E402
from quirk_cases import unclean_key_case  # noqa: E402
from visioncortex_b200._ffi import VcbError  # noqa: E402
from visioncortex_b200.runtime import Context  # noqa: E402

HM = 0xFFFFFFFF


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2024
    rng = np.random.default_rng(seed)
    gpu, oracle = Context(0), load_oracle()
    t0 = time.time()
    n = {"colour": 0, "unclean": 0, "binary": 0, "batch": 0, "rejected": 0, "sequential": 0, "resurrected": 0}
    k = 0
    while time.time() - t0 < budget:
        k += 1
        fam = k % 8
        if fam in (0, 1, 2, 3, 4):
            side = [24, 60, 150, 300, 40][fam]
            hier = [0, HM, 64, 3, HM][int(rng.integers(0, 5))]
            img, cfg = random_case(rng, max_side=side, keyed=bool(rng.integers(0, 2)), hierarchical=hier, diagonal=bool(rng.integers(0, 2)))
            cfg.good_min_area = int(rng.choice([0, 2, 16])); cfg.good_max_area = int(rng.choice([50, 65536]))
            cfg.deepen_diff = int(rng.choice([0, 16, 64])); cfg.hollow_neighbours = int(rng.choice([0, 1, 2]))
            kind = "colour"
        elif fam in (5, 6):
            img, cfg = unclean_key_case(rng, max_side=[20, 70][fam - 5], hierarchical=[0, HM, 48][int(rng.integers(0, 3))])
            kind = "unclean"
        else:
            mask = rng.random((int(rng.integers(1, 200)), int(rng.integers(1, 200)))) < float(rng.choice([0.3, 0.55, 0.62, 0.8]))
            diag = bool(rng.integers(0, 2))
            assert_same_bin_clusters(gpu.to_clusters(mask, diag), oracle.to_clusters(mask, diag), ctx=f"binary {k}")
            n["binary"] += 1
            continue
        try:
            ref = oracle.runner_run(img, cfg, pools=True, color_image=True)
        except VcbError as e:
            # the reference panics on this input (division by zero on the empty cluster an orphan pixel's label names,
            # SURVEY Appendix E item 5): the device must reject it too
            assert e.status == 3 and cfg.hierarchical != 0 and cfg.diagonal, (k, str(e))
            try:
                gpu.runner_run(img, cfg, pools=False)
                raise AssertionError(f"case {k}: the reference panics, the device returned a result")
            except VcbError as e2:
                assert e2.status == 3 and "overwritten" in str(e2), (k, str(e2))
            n["panic"] = n.get("panic", 0) + 1
            continue
        try:
            got = gpu.runner_run(img, cfg, pools=True, color_image=True)
        except VcbError as e:
            assert e.status == 3 and cfg.hierarchical != 0 and cfg.diagonal and "overwritten" in str(e), (k, str(e))
            n["rejected"] += 1
            continue
        lc = gpu.last_counters()
        n["sequential"] += lc["sequential"]
        n["resurrected"] += 1 if lc["resurrected_leaves"] else 0
        assert_same_clusters(got, ref, pools=True, ctx=f"{kind} case {k} {img.shape}")
        n[kind] += 1
        if k % 16 == 0:                                        # the same image three times in one batch
            res = gpu.runner_run_batch([img, img, img], cfg, pools=False)
            for r in res:
                assert_same_clusters(r, ref, pools=False, ctx=f"batch case {k}")
            n["batch"] += 1
    print("fuzz ok: %s in %.0f s (seed %d)" % (n, time.time() - t0, seed), flush=True)


if __name__ == "__main__":
    main()
