"""Committed golden vectors (tests/golden/golden_small.npz, written by tests/golden/make_golden.py): stored inputs with the
complete expected results.  CPU: the oracle must reproduce them (freezes the checker).  GPU: the CUDA path must reproduce
them through the C ABI with nothing under oracle/ involved.  Provenance of the vectors: see make_golden.py."""
import os

import numpy as np
import pytest

from parity import REC_FIELDS, assert_same_bin_clusters, assert_same_clusters
from visioncortex_b200._ffi import RunnerConfigC

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.npz")
CFG_FIELDS = ("diagonal", "hierarchical", "batch_size", "good_min_area", "good_max_area", "is_same_color_a",
              "is_same_color_b", "deepen_diff", "hollow_neighbours", "keying_action")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def cfg_from_array(a) -> RunnerConfigC:
    cfg = RunnerConfigC()
    for f, v in zip(CFG_FIELDS, a[:len(CFG_FIELDS)]):
        setattr(cfg, f, int(v))
    for k in range(4):
        cfg.key_color[k] = int(a[len(CFG_FIELDS) + k])
    return cfg


def colour_expected(g, i) -> dict:
    img = g[f"c{i}_img"]
    n = int(g[f"c{i}_num_slots"][0])
    rec = {f: g[f"c{i}_rec_{f}"] for f in REC_FIELDS}
    return {"width": img.shape[1], "height": img.shape[0], "cluster_indices": g[f"c{i}_labels"], "num_slots": n,
            "records": _Rec(rec, n), "clusters_output": g[f"c{i}_outputs"], "indices_pool": g[f"c{i}_indices"],
            "holes_pool": g[f"c{i}_holes"], "color_image": g[f"c{i}_color_image"]}


class _Rec(dict):
    """field -> array mapping that also answers len() like the structured record array the C ABI readers return"""

    def __init__(self, fields, n):
        super().__init__(fields)
        self._n = n

    def __len__(self):
        return self._n


def bin_expected(g, i) -> dict:
    rec = {f: g[f"b{i}_rec_{f}"] for f in ("rect", "points_off", "points_len")}
    return {"rect": g[f"b{i}_rect"], "records": _Rec(rec, len(rec["points_len"])), "points": g[f"b{i}_points"]}


def check_all(g, run_colour, run_binary):
    for i in range(int(g["n_colour"][0])):
        got = run_colour(g[f"c{i}_img"], cfg_from_array(g[f"c{i}_cfg"]))
        assert_same_clusters(got, colour_expected(g, i), pools=True, ctx=f"golden colour case {i}")
    for i in range(int(g["n_bin"][0])):
        got = run_binary(g[f"b{i}_mask"], bool(g[f"b{i}_diag"][0]))
        assert_same_bin_clusters(got, bin_expected(g, i), ctx=f"golden binary case {i}")


def test_golden_file_is_complete(golden):
    assert int(golden["n_colour"][0]) >= 30 and int(golden["n_bin"][0]) >= 12
    hier = {int(golden[f"c{i}_cfg"][1]) for i in range(int(golden["n_colour"][0]))}
    assert {0, 64, 0xFFFFFFFF} <= hier
    assert any(golden[f"c{i}_cfg"][len(CFG_FIELDS):].any() for i in range(int(golden["n_colour"][0])))   # keyed cases


def test_oracle_reproduces_golden(golden, oracle):
    check_all(golden, lambda im, cfg: oracle.runner_run(im, cfg, pools=True, color_image=True), oracle.to_clusters)


@pytest.mark.gpu
def test_gpu_reproduces_golden(golden, gpu):
    check_all(golden, lambda im, cfg: gpu.runner_run(im, cfg, pools=True, color_image=True), gpu.to_clusters)
