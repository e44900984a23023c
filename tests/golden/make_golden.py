#!/usr/bin/env python
"""Generates tests/golden/golden_small.npz: small seeded inputs (stored verbatim) with the complete expected results of
`color_clusters::Runner` and `BinaryImage::to_clusters`.

Provenance, stated plainly: the expected values are produced by `oracle/` (the C++ restatement of the reference algorithm),
NOT by the Rust reference itself -- there is no Rust toolchain in this image.  The oracle is pinned to the reference by the
reference's own in-source known-answer tests (tests/test_oracle_kat.py, citing clusters.rs / builder.rs).  These vectors
therefore (a) freeze the oracle against accidental change and (b) let the GPU tests check the CUDA path against committed
data without building or calling anything under oracle/.

    python tests/golden/make_golden.py        # rewrites golden_small.npz next to this file
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle_lib import load_oracle  # noqa: E402
from parity import REC_FIELDS, random_case  # noqa: E402

CFG_FIELDS = ("diagonal", "hierarchical", "batch_size", "good_min_area", "good_max_area", "is_same_color_a",
              "is_same_color_b", "deepen_diff", "hollow_neighbours", "keying_action")


def cfg_to_array(cfg):
    return np.array([int(getattr(cfg, f)) for f in CFG_FIELDS] + [int(x) for x in cfg.key_color], dtype=np.int64)


def main():
    oracle = load_oracle()
    rng = np.random.default_rng(20261020)
    out = {}
    n_colour = 0
    for k in range(36):
        hier = [0, 0xFFFFFFFF, 64][k % 3]
        img, cfg = random_case(rng, max_side=30, hierarchical=hier, keyed=(k % 4 == 3), diagonal=(k % 5 == 4))
        try:
            res = oracle.runner_run(img, cfg, pools=True, color_image=True)
        except Exception:
            continue                      # inputs the reference treats as order-dependent (fenced) are not golden material
        i = n_colour
        out[f"c{i}_img"] = img
        out[f"c{i}_cfg"] = cfg_to_array(cfg)
        out[f"c{i}_labels"] = res["cluster_indices"]
        out[f"c{i}_num_slots"] = np.array([res["num_slots"]], dtype=np.int64)
        for f in REC_FIELDS:
            out[f"c{i}_rec_{f}"] = np.ascontiguousarray(res["records"][f])
        out[f"c{i}_outputs"] = res["clusters_output"]
        out[f"c{i}_indices"] = res["indices_pool"]
        out[f"c{i}_holes"] = res["holes_pool"]
        out[f"c{i}_color_image"] = res["color_image"]
        n_colour += 1
    n_bin = 0
    for k in range(16):
        w, h = int(rng.integers(1, 34)), int(rng.integers(1, 34))
        mask = rng.random((h, w)) < rng.choice([0.3, 0.5, 0.7, 0.9])
        diag = bool(k % 2)
        res = oracle.to_clusters(mask, diag)
        i = n_bin
        out[f"b{i}_mask"] = mask
        out[f"b{i}_diag"] = np.array([int(diag)], dtype=np.int64)
        out[f"b{i}_rect"] = np.asarray(res["rect"])
        for f in ("rect", "points_off", "points_len"):
            out[f"b{i}_rec_{f}"] = np.ascontiguousarray(res["records"][f])
        out[f"b{i}_points"] = res["points"]
        n_bin += 1
    out["n_colour"] = np.array([n_colour], dtype=np.int64)
    out["n_bin"] = np.array([n_bin], dtype=np.int64)
    path = os.path.join(HERE, "golden_small.npz")
    np.savez_compressed(path, **out)
    print(f"{n_colour} colour cases, {n_bin} binary cases -> {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
