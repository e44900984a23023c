"""The two CPU restatements of the reference — oracle/vc_oracle.cpp (C++) and oracle/vc_restate.py (pure Python, written
independently from the Rust sources) — must agree bit for bit: labels, slot count, every record field, clusters_output,
the ordered indices / holes lists and to_color_image.  Both are sequential, so they must also agree on the
order-dependent quirk families the parallel formulation has to treat specially: `diagonal = true` (stale
`cluster_upleft`, builder.rs:301-305 vs :341-343) and unclean key colours (builder.rs:330-334 with runner.rs:116-129).
This is the pin of the colour path (the reference has no colour tests): SURVEY.md section 8(c), DESIGN.md section 2."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import vc_restate as R  # noqa: E402
from parity import assert_same_bin_clusters, assert_same_clusters, random_colour_image  # noqa: E402
from visioncortex_b200._ffi import VcbError  # noqa: E402
from visioncortex_b200.runtime import make_config  # noqa: E402

HM = 0xFFFFFFFF


def both_configs(**kw):
    return make_config(**kw), R.Config(**{k: (bool(v) if k == "diagonal" else v) for k, v in kw.items()})


def fuzz_case(rng, k, max_side):
    w, h = int(rng.integers(1, max_side + 1)), int(rng.integers(1, max_side + 1))
    img = random_colour_image(rng, w, h)
    kw = dict(diagonal=int(k % 2), hierarchical=[0, HM, 64, 3, HM, 0][k % 6], is_same_color_a=int(rng.choice([0, 2, 4, 5])),
              is_same_color_b=int(rng.choice([0, 1, 2])), good_min_area=int(rng.choice([0, 2, 16])),
              good_max_area=int(rng.choice([50, 65536])), deepen_diff=int(rng.choice([0, 16, 64])),
              hollow_neighbours=int(rng.choice([0, 1, 2])))
    fam = k % 5
    if fam in (2, 3, 4):
        if fam == 2:                      # clean key: far from every image colour family in shifted space is not guaranteed; still sequential
            key = (7, 250, 3, 255)
        else:                             # unclean key: the key is one of the image's own colours / a near neighbour of them
            src = img[int(rng.integers(0, h)), int(rng.integers(0, w))]
            key = (int(src[0]), int(src[1]), int(src[2]), 255)
        m = rng.random((h, w)) < (0.25 if fam != 4 else 0.08)
        img[m] = np.array(key, dtype=np.uint8)
        kw["key_color"] = key
        kw["keying_action"] = int(rng.integers(0, 2))
    return img, kw


def run_pair(oracle, img, kw):
    ccfg, pcfg = both_configs(**kw)
    try:
        want = R.runner_run(img, pcfg, color_image=True)
    except R.ReferencePanic as e:
        # the Rust itself would panic here; the C++ oracle must say so too (status 3) when it is the division by zero in
        # ColorSum::average -- the panic the hierarchical stage can really reach (orphan pixel of an overwritten slot)
        if "divide by zero" in str(e):
            from visioncortex_b200._ffi import VcbError
            with pytest.raises(VcbError) as ce:
                oracle.runner_run(img, ccfg, pools=False)
            assert ce.value.status == 3
        return "panic", str(e)
    got = oracle.runner_run(img, ccfg, pools=True, color_image=True)
    return got, want


@pytest.mark.parametrize("chunk", range(12))
def test_colour_restatements_agree(oracle, chunk):
    rng = np.random.default_rng(9000 + chunk)
    compared = panics = 0
    for k in range(1000):
        img, kw = fuzz_case(rng, k, max_side=11 if k % 10 else 20)
        r = run_pair(oracle, img, kw)
        if r[0] == "panic":
            panics += 1          # inputs on which the Rust itself would panic (division by zero in average()): not comparable
            continue
        got, want = r
        assert_same_clusters(got, want, pools=True, ctx=f"chunk {chunk} case {k} {img.shape} {kw}")
        compared += 1
    assert compared >= 950, (compared, panics)


def test_colour_restatements_agree_medium(oracle):
    rng = np.random.default_rng(77)
    for k in range(40):
        img, kw = fuzz_case(rng, k, max_side=64)
        r = run_pair(oracle, img, kw)
        if r[0] == "panic":
            continue
        assert_same_clusters(r[0], r[1], pools=True, ctx=f"medium case {k} {img.shape} {kw}")


def test_binary_restatements_agree(oracle):
    rng = np.random.default_rng(5)
    for k in range(1500):
        w, h = int(rng.integers(1, 24)), int(rng.integers(1, 24))
        mask = rng.random((h, w)) < rng.choice([0.05, 0.3, 0.5, 0.7, 0.9, 1.0])
        diag = bool(k % 2)
        assert_same_bin_clusters(oracle.to_clusters(mask, diag), R.to_clusters(mask, diag), ctx=f"bin {k} {w}x{h} {diag}")


def test_python_restatement_reference_kats():
    """The reference's own to_clusters unit tests (clusters.rs:355-418) on the Python restatement."""
    img = np.zeros((3, 3), dtype=bool)
    img[0, 0] = img[1, 1] = img[2, 2] = True
    res = R.to_clusters(img, False)
    assert len(res["records"]) == 3 and res["points"].tolist() == [[0, 0], [1, 1], [2, 2]]
    assert tuple(res["records"][2]["rect"]) == (2, 2, 3, 3)
    res = R.to_clusters(img, True)
    assert len(res["records"]) == 1 and res["points"].tolist() == [[0, 0], [1, 1], [2, 2]]
    img = np.zeros((4, 4), dtype=bool)
    img[1:3, 1:3] = True
    res = R.to_clusters(img, False)
    assert len(res["records"]) == 1 and tuple(res["records"][0]["rect"]) == (1, 1, 3, 3)
