"""Fuzzes the parallel id model (tests/model_stage1.py — the executable spec of what the CUDA kernels compute)
against the sequential oracle on CPU.  Sized to run in well under a minute; `python tests/test_model_vs_oracle.py N`
runs a longer campaign."""
import sys

import numpy as np

from model_stage1 import Unsupported, run_model, run_model_fixpoint, static_edges_binary, static_edges_colour
from test_oracle_kat import _cfg


def random_colour_image(rng, w, h):
    fam = rng.integers(0, 4)
    if fam == 0:      # few flat colours
        pal = rng.integers(0, 256, size=(rng.integers(2, 5), 3))
        img = pal[rng.integers(0, len(pal), size=(h, w))]
    elif fam == 1:    # near-threshold noise around one colour
        img = 100 + rng.integers(-9, 10, size=(h, w, 3))
    elif fam == 2:    # blocky
        b = int(rng.integers(2, 4))
        pal = rng.integers(0, 8, size=(3, 3)) * 32
        idx = rng.integers(0, 3, size=((h + b - 1) // b, (w + b - 1) // b))
        img = pal[idx].repeat(b, 0).repeat(b, 1)[:h, :w] + rng.integers(-3, 4, size=(h, w, 3))
    else:             # gradient + noise
        ys, xs = np.mgrid[0:h, 0:w]
        img = np.stack([xs * 9 + ys * 3, ys * 11, (xs + ys) * 5], -1) + rng.integers(-4, 5, size=(h, w, 3))
    out = np.empty((h, w, 4), dtype=np.uint8)
    out[..., :3] = np.clip(img, 0, 255).astype(np.uint8)
    out[..., 3] = 255
    return out


def check_colour_case(oracle, rng, diagonal=False, keyed=False):
    w, h = int(rng.integers(1, 14)), int(rng.integers(1, 14))
    img = random_colour_image(rng, w, h)
    shift = int(rng.choice([0, 2, 4, 5]))
    thres = int(rng.choice([0, 1, 2]))
    key = (0, 0, 0, 0)
    action = 0
    if keyed:
        key = (7, 250, 3, 255)
        m = rng.random((h, w)) < 0.25
        img[m] = np.array(key, dtype=np.uint8)
        action = int(rng.integers(0, 2))
    cfg = _cfg(hierarchical=0, diagonal=int(diagonal), is_same_color_a=shift, is_same_color_b=thres,
               key_color=key, keying_action=action)
    try:
        J, M, kmask = static_edges_colour(img, diagonal, shift, thres, key, keyed)
        mod = run_model_fixpoint(J, M, w, h, base=1) if diagonal else run_model(J, M, w, h, base=1, check_stale_upleft=False)
    except Unsupported:
        return False
    ref = oracle.runner_run(img, cfg)
    labels = np.where(mod["labels"] < 0, 0, mod["labels"]).astype(np.uint32)
    assert (labels == ref["cluster_indices"]).all(), (img.tolist(), shift, thres, diagonal)
    assert mod["num_slots"] == ref["num_slots"], (mod["num_slots"], ref["num_slots"])
    for s in range(ref["num_slots"]):
        r = ref["records"][s]
        got = ref["indices_pool"][int(r["indices_off"]): int(r["indices_off"]) + int(r["indices_len"])].tolist()
        if s == 0:
            want = np.nonzero(kmask)[0].tolist() if (keyed and action == 0) else []
        else:
            want = mod["listing"].get(s, [])
        assert got == want, (s, got, want)
    return True


def check_binary_case(oracle, rng):
    w, h = int(rng.integers(1, 16)), int(rng.integers(1, 16))
    mask = rng.random((h, w)) < rng.choice([0.3, 0.5, 0.7, 0.9])
    diagonal = bool(rng.integers(0, 2))
    J, M = static_edges_binary(mask, diagonal)
    mod = run_model(J, M, w, h, base=0, check_stale_upleft=False)
    ref = oracle.to_clusters(mask, diagonal)
    live = sorted(mod["listing"].keys())
    assert len(live) == len(ref["records"])
    for k, s in enumerate(live):
        r = ref["records"][k]
        pts = ref["points"][int(r["points_off"]): int(r["points_off"]) + int(r["points_len"])]
        got = (pts[:, 1] * w + pts[:, 0]).tolist()
        assert got == mod["listing"][s], (k, got, mod["listing"][s])


def test_model_colour(oracle):
    rng = np.random.default_rng(1)
    for _ in range(300):
        check_colour_case(oracle, rng)


def test_model_colour_keyed(oracle):
    rng = np.random.default_rng(2)
    n = sum(check_colour_case(oracle, rng, keyed=True) for _ in range(300))
    assert n > 100


def test_model_colour_diagonal(oracle):
    rng = np.random.default_rng(3)
    n = sum(check_colour_case(oracle, rng, diagonal=True, keyed=bool(k % 2)) for k in range(300))
    assert n > 200


def test_model_binary(oracle):
    rng = np.random.default_rng(4)
    for _ in range(300):
        check_binary_case(oracle, rng)


if __name__ == "__main__":
    from oracle_lib import load_oracle

    o = load_oracle()
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    rng = np.random.default_rng(1234)
    ok = 0
    for k in range(N):
        ok += check_colour_case(o, rng, diagonal=bool(k % 3 == 0), keyed=bool(k % 2))
        check_binary_case(o, rng)
    print("colour cases checked:", ok, "of", N, "(rest fenced as unsupported); binary:", N)
