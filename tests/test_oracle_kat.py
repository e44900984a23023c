"""Pins the CPU oracle against the reference's own unit tests for `BinaryImage::to_clusters`
(reference src/clusters.rs:351-526) and supporting KATs (src/bound.rs:659-682).  `color_clusters` has no tests in
the reference, so the colour path has only structural checks here (Appendix A consequences in SURVEY.md).
"""
import numpy as np

from visioncortex_b200._ffi import RunnerConfigC, HIERARCHICAL_MAX


def from_string(s: str) -> np.ndarray:
    """BinaryImage::from_string (image/format.rs:153-169): '*' = set, rows separated by newline."""
    rows = s.strip("\n").split("\n")
    return np.array([[c == "*" for c in r] for r in rows], dtype=bool)


def cluster_points(res, k):
    r = res["records"][k]
    return res["points"][int(r["points_off"]): int(r["points_off"]) + int(r["points_len"])]


def cluster_to_string(res, k) -> str:
    """clusters::Cluster::to_binary_image (clusters.rs:38-49) + Display (format.rs:127-140)."""
    l, t, rgt, b = [int(v) for v in res["records"][k]["rect"]]
    img = np.zeros((b - t, rgt - l), dtype=bool)
    for x, y in cluster_points(res, k):
        img[y - t, x - l] = True
    return "".join("".join("*" if v else "-" for v in row) + "\n" for row in img)


# ---- clusters.rs:355-377
def test_clusters_3x3(oracle):
    img = np.zeros((3, 3), dtype=bool)
    img[0, 0] = img[1, 1] = img[2, 2] = True
    res = oracle.to_clusters(img, False)
    assert len(res["records"]) == 3
    for k, (x, y) in enumerate([(0, 0), (1, 1), (2, 2)]):
        assert res["records"][k]["points_len"] == 1
        assert tuple(cluster_points(res, k)[0]) == (x, y)
    assert tuple(res["records"][2]["rect"]) == (2, 2, 3, 3)       # BoundingRect::default().add_x_y(2,2)
    assert cluster_to_string(res, 0) == "*\n"


# ---- clusters.rs:379-392 (asserts point order)
def test_clusters_3x3_diagonal(oracle):
    img = np.zeros((3, 3), dtype=bool)
    img[0, 0] = img[1, 1] = img[2, 2] = True
    res = oracle.to_clusters(img, True)
    assert len(res["records"]) == 1
    assert res["records"][0]["points_len"] == 3
    assert [tuple(p) for p in cluster_points(res, 0)] == [(0, 0), (1, 1), (2, 2)]


# ---- clusters.rs:394-418
def test_clusters_4x4(oracle):
    img = np.zeros((4, 4), dtype=bool)
    img[1:3, 1:3] = True
    res = oracle.to_clusters(img, False)
    assert len(res["records"]) == 1
    assert res["records"][0]["points_len"] == 4
    assert tuple(res["records"][0]["rect"]) == (1, 1, 3, 3)
    assert cluster_to_string(res, 0) == "**\n**\n"


# ---- break_cluster (clusters.rs:181-235) restated on top of the oracle's to_clusters, to replay the six
#      order-sensitive indirect tests clusters.rs:420-525.
def _break_cluster_recursive(oracle, img: np.ndarray, left: int, top: int, out: list):
    image = img.copy()
    hgt, wid = image.shape
    broke = False
    if wid >= 2 and hgt >= 3:
        for y in range(hgt - 3 + 1):
            for x in range(wid - 2 + 1):
                if (image[y, x] != image[y, x + 1] and image[y + 1, x] and image[y + 1, x + 1]
                        and image[y + 2, x] != image[y + 2, x + 1] and image[y, x] == image[y + 2, x + 1]):
                    if x < wid - 2 and image[y + 1, x + 2]:
                        image[y + 1, x + 1] = False
                        broke = True
                    elif x > 0 and image[y + 1, x - 1]:
                        image[y + 1, x] = False
                        broke = True
                    if broke:
                        break
            if broke:
                break
    res = oracle.to_clusters(image, False)
    if broke and min(int(r["points_len"]) for r in res["records"]) < 5:
        broke = False
    if broke:
        for k in range(len(res["records"])):
            l, t, rgt, b = [int(v) for v in res["records"][k]["rect"]]
            sub = np.zeros((b - t, rgt - l), dtype=bool)
            for x, y in cluster_points(res, k):
                sub[y - t, x - l] = True
            _break_cluster_recursive(oracle, sub, left + l, top + t, out)
    else:
        out.append((left, top, img))


def break_cluster(oracle, s: str):
    img = from_string(s)
    res = oracle.to_clusters(img, False)
    assert len(res["records"]) >= 1
    l, t, rgt, b = [int(v) for v in res["records"][0]["rect"]]
    sub = np.zeros((b - t, rgt - l), dtype=bool)
    for x, y in cluster_points(res, 0):
        sub[y - t, x - l] = True
    out = []
    _break_cluster_recursive(oracle, sub, l, t, out)
    return out


def _str(img):
    return "".join("".join("*" if v else "-" for v in row) + "\n" for row in img)


def test_break_cluster_noop(oracle):
    s = "***\n***\n-*-\n"
    out = break_cluster(oracle, s)
    assert len(out) == 1 and _str(out[0][2]) == s


def test_break_cluster(oracle):
    out = break_cluster(oracle, "***---\n******\n---***\n")
    assert len(out) == 2
    assert _str(out[0][2]) == "***\n***\n" and out[0][:2] == (0, 0)
    assert _str(out[1][2]) == "-**\n***\n" and out[1][:2] == (3, 1)


def test_break_cluster_alt(oracle):
    out = break_cluster(oracle, "---***\n******\n***---\n")
    assert len(out) == 2
    assert _str(out[0][2]) == "***\n-**\n" and out[0][:2] == (3, 0)
    assert _str(out[1][2]) == "***\n***\n" and out[1][:2] == (0, 1)


def test_break_cluster_cant_break(oracle):
    s = "*--\n**-\n-**\n"
    out = break_cluster(oracle, s)
    assert len(out) == 1 and _str(out[0][2]) == s


def test_break_cluster_cant_break_2(oracle):
    s = "*---\n****\n--**\n"
    out = break_cluster(oracle, s)
    assert len(out) == 1 and _str(out[0][2]) == s


def test_break_cluster_big(oracle):
    out = break_cluster(oracle, "***---***\n*********\n---***---\n")
    assert len(out) == 3
    assert _str(out[0][2]) == "***\n***\n" and out[0][:2] == (0, 0)
    assert _str(out[1][2]) == "***\n-**\n" and out[1][:2] == (6, 0)
    assert _str(out[2][2]) == "-**\n***\n" and out[2][:2] == (3, 1)


# ---- bound.rs:659-682 (bounding_rect_1x1 / 2x2) through a to_clusters call
def test_bounding_rect_kats(oracle):
    one = np.zeros((3, 3), dtype=bool)
    one[1, 1] = True
    r = oracle.to_clusters(one, False)
    assert tuple(r["rect"]) == (1, 1, 2, 2)
    two = np.zeros((4, 4), dtype=bool)
    two[1, 1] = two[2, 2] = True
    r = oracle.to_clusters(two, True)
    assert tuple(r["rect"]) == (1, 1, 3, 3)


# ---- u16 label overflow (clusters.rs:322-324): isolated pixels on a checker grid, 65535th NEW label panics
def test_label_overflow(oracle):
    import pytest
    from visioncortex_b200._ffi import VcbError, VCB_ERR_LABEL_OVERFLOW

    img = np.zeros((512, 514), dtype=bool)
    img[::2, ::2] = True                      # 256*257 = 65792 isolated pixels > 65534
    with pytest.raises(VcbError) as e:
        oracle.to_clusters(img, False)
    assert e.value.status == VCB_ERR_LABEL_OVERFLOW
    img2 = np.zeros((512, 512), dtype=bool)
    img2[::2, ::2] = True                     # 65536 isolated pixels -> also overflows (label 65534 is fatal)
    with pytest.raises(VcbError):
        oracle.to_clusters(img2, False)
    img3 = np.zeros((510, 514), dtype=bool)
    img3[::2, ::2] = True                     # 255*257 = 65535 NEW labels -> counter reaches 65535 -> panic
    with pytest.raises(VcbError):
        oracle.to_clusters(img3, False)
    img4 = img3.copy()
    img4[0, 0] = False                        # 65534 labels: counter ends at 65534 -> fine
    assert len(oracle.to_clusters(img4, False)["records"]) == 65534


# ---- SURVEY Appendix A consequences for the colour path (structural, not reference-pinned)
def _cfg(**kw):
    c = RunnerConfigC()
    c.diagonal = 0
    c.hierarchical = HIERARCHICAL_MAX
    c.batch_size = 25600
    c.good_min_area = 16
    c.good_max_area = 256 * 256
    c.is_same_color_a = 4
    c.is_same_color_b = 1
    c.deepen_diff = 64
    c.hollow_neighbours = 1
    c.keying_action = 0
    for k, v in kw.items():
        if k == "key_color":
            for j in range(4):
                c.key_color[j] = v[j]
        else:
            setattr(c, k, v)
    return c


def test_flat_3x3_colour(oracle):
    img = np.full((3, 3, 4), 200, dtype=np.uint8)
    res = oracle.runner_run(img, _cfg(hierarchical=0))
    assert res["cluster_indices"].reshape(3, 3).tolist() == [[1, 2, 2], [2, 2, 2], [2, 2, 2]]
    assert [int(r["indices_len"]) for r in res["records"]] == [0, 1, 8, 0, 0]
    assert res["clusters_output"].tolist() == [1, 2]
    # batch_size only slices the work (builder.rs:281,357)
    res2 = oracle.runner_run(img, _cfg(hierarchical=0, batch_size=1))
    assert (res2["cluster_indices"] == res["cluster_indices"]).all()


def test_invalid_args(oracle):
    import pytest
    from visioncortex_b200._ffi import VcbError

    img = np.zeros((2, 2, 4), dtype=np.uint8)
    with pytest.raises(VcbError):
        oracle.runner_run(img, _cfg(is_same_color_a=8))
    with pytest.raises(VcbError):
        oracle.runner_run(img, _cfg(batch_size=0))
