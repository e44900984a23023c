"""Shared differential-parity helpers: compare a result produced through the C ABI (CUDA product library on the GPU,
or the host-simulator build of the same kernels) with the CPU oracle, field by field, bit-exact."""
import numpy as np

from visioncortex_b200.runtime import make_config

REC_FIELDS = ("sum", "residue_sum", "rect", "num_holes", "depth", "merged_into", "indices_len", "indices_off",
              "holes_off", "holes_len")


def assert_same_clusters(got: dict, ref: dict, pools: bool = True, ctx=""):
    assert got["width"] == ref["width"] and got["height"] == ref["height"], ctx
    w = got["width"]
    gl, rl = got["cluster_indices"], ref["cluster_indices"]
    if not np.array_equal(gl, rl):
        bad = np.nonzero(gl != rl)[0]
        raise AssertionError(f"{ctx}: cluster_indices differ at {len(bad)} pixels, first {bad[0]} (x={bad[0] % w}, y={bad[0] // w}): got {gl[bad[0]]} want {rl[bad[0]]}")
    assert got["num_slots"] == ref["num_slots"], f"{ctx}: slots {got['num_slots']} != {ref['num_slots']}"
    for f in REC_FIELDS:
        if not np.array_equal(got["records"][f], ref["records"][f]):
            bad = np.nonzero((got["records"][f] != ref["records"][f]).reshape(len(got["records"]), -1).any(axis=1))[0]
            raise AssertionError(f"{ctx}: record field {f} differs at slot {bad[0]}: got {got['records'][f][bad[0]]} want {ref['records'][f][bad[0]]}")
    assert np.array_equal(got["clusters_output"], ref["clusters_output"]), \
        f"{ctx}: clusters_output differ: got {got['clusters_output'][:20]} want {ref['clusters_output'][:20]}"
    if pools:
        assert np.array_equal(got["indices_pool"], ref["indices_pool"]), f"{ctx}: indices pool differs"
        assert np.array_equal(got["holes_pool"], ref["holes_pool"]), f"{ctx}: holes pool differs"
    if "color_image" in ref and "color_image" in got:
        assert np.array_equal(got["color_image"], ref["color_image"]), f"{ctx}: to_color_image differs"


def assert_same_bin_clusters(got: dict, ref: dict, ctx=""):
    assert np.array_equal(got["rect"], ref["rect"]), f"{ctx}: overall rect {got['rect']} != {ref['rect']}"
    assert len(got["records"]) == len(ref["records"]), f"{ctx}: cluster count {len(got['records'])} != {len(ref['records'])}"
    for f in ("rect", "points_off", "points_len"):
        assert np.array_equal(got["records"][f], ref["records"][f]), f"{ctx}: bin record field {f} differs"
    assert np.array_equal(got["points"], ref["points"]), f"{ctx}: points differ"


def random_colour_image(rng, w, h):
    fam = int(rng.integers(0, 5))
    if fam == 0:      # few flat colours
        pal = rng.integers(0, 256, size=(int(rng.integers(2, 5)), 3))
        img = pal[rng.integers(0, len(pal), size=(h, w))]
    elif fam == 1:    # near-threshold noise around one colour
        img = 100 + rng.integers(-9, 10, size=(h, w, 3))
    elif fam == 2:    # blocky
        b = int(rng.integers(2, 5))
        pal = rng.integers(0, 8, size=(3, 3)) * 32
        idx = rng.integers(0, 3, size=((h + b - 1) // b, (w + b - 1) // b))
        img = pal[idx].repeat(b, 0).repeat(b, 1)[:h, :w] + rng.integers(-3, 4, size=(h, w, 3))
    elif fam == 3:    # gradient + noise
        ys, xs = np.mgrid[0:h, 0:w]
        img = np.stack([xs * 9 + ys * 3, ys * 11, (xs + ys) * 5], -1) + rng.integers(-4, 5, size=(h, w, 3))
    else:             # large blocks with light noise (big clusters absorbing speckles)
        b = int(rng.integers(5, 12))
        pal = rng.integers(0, 4, size=(4, 3)) * 64
        idx = rng.integers(0, 4, size=((h + b - 1) // b, (w + b - 1) // b))
        img = pal[idx].repeat(b, 0).repeat(b, 1)[:h, :w] + rng.integers(-5, 6, size=(h, w, 3))
    out = np.empty((h, w, 4), dtype=np.uint8)
    out[..., :3] = np.clip(img, 0, 255).astype(np.uint8)
    out[..., 3] = 255
    return out


def random_case(rng, max_side=40, hierarchical=0, keyed=False, diagonal=False):
    w, h = int(rng.integers(1, max_side + 1)), int(rng.integers(1, max_side + 1))
    img = random_colour_image(rng, w, h)
    kw = dict(hierarchical=hierarchical, diagonal=int(diagonal), is_same_color_a=int(rng.choice([0, 2, 4, 5])),
              is_same_color_b=int(rng.choice([0, 1, 2])))
    if keyed:
        key = (7, 250, 3, 255)
        m = rng.random((h, w)) < 0.25
        img[m] = np.array(key, dtype=np.uint8)
        kw["key_color"] = key
        kw["keying_action"] = int(rng.integers(0, 2))
    return img, make_config(**kw)
