"""CPU-side check of the kernel *logic*: the product kernel sources compiled against the host CUDA simulator
(tests/hostsim, TEST INFRASTRUCTURE) and compared with the oracle on small inputs.  This is not a product path: the
product library has no CPU build; the real parity tests are the `-m gpu` ones."""
import numpy as np
import pytest

from parity import assert_same_bin_clusters, assert_same_clusters, random_case
from test_oracle_kat import from_string
from visioncortex_b200._ffi import VcbError


@pytest.fixture(scope="module")
def sim():
    from sim_lib import load_sim

    return load_sim()


def test_sim_colour_stage1(sim, oracle):
    rng = np.random.default_rng(21)
    done = 0
    for k in range(12):
        img, cfg = random_case(rng, max_side=28, keyed=(k % 3 == 2))
        ref = oracle.runner_run(img, cfg, pools=True, color_image=True)
        got = sim.runner_run(img, cfg, pools=True, color_image=True)
        assert_same_clusters(got, ref, pools=True, ctx=f"case {k}")
        done += 1
    assert done == 12


def test_sim_binary(sim, oracle):
    rng = np.random.default_rng(22)
    for k in range(12):
        w, h = int(rng.integers(1, 30)), int(rng.integers(1, 30))
        mask = rng.random((h, w)) < rng.choice([0.3, 0.5, 0.7, 0.9])
        diag = bool(k % 2)
        assert_same_bin_clusters(sim.to_clusters(mask, diag), oracle.to_clusters(mask, diag), ctx=f"case {k}")


def test_sim_binary_reference_kats(sim):
    # clusters.rs:355-392
    img = np.zeros((3, 3), dtype=bool)
    img[0, 0] = img[1, 1] = img[2, 2] = True
    res = sim.to_clusters(img, False)
    assert len(res["records"]) == 3 and res["points"].tolist() == [[0, 0], [1, 1], [2, 2]]
    res = sim.to_clusters(img, True)
    assert len(res["records"]) == 1 and res["points"].tolist() == [[0, 0], [1, 1], [2, 2]]
    # clusters.rs:394-418
    img = np.zeros((4, 4), dtype=bool)
    img[1:3, 1:3] = True
    res = sim.to_clusters(img, False)
    assert len(res["records"]) == 1 and tuple(res["records"][0]["rect"]) == (1, 1, 3, 3)
    assert tuple(res["rect"]) == (1, 1, 3, 3)


def test_sim_empty_and_full(sim, oracle):
    for mask in (np.zeros((5, 7), dtype=bool), np.ones((6, 4), dtype=bool)):
        for diag in (False, True):
            assert_same_bin_clusters(sim.to_clusters(mask, diag), oracle.to_clusters(mask, diag))


def test_sim_diagonal_stale_upleft(sim, oracle):
    """diagonal = true inputs on which the reference's stale cluster_upleft read resurrects a dead slot (builder.rs:301-305 vs
    :341-343): the executable id model (model_stage1.run_model_fixpoint) picks images that do hit it, the kernels must agree
    with the oracle on them and must have taken the J_RES path."""
    from model_stage1 import Unsupported, run_model_fixpoint, static_edges_colour
    from visioncortex_b200.runtime import make_config

    rng = np.random.default_rng(91)
    done = 0
    for k in range(2000):
        w, h = int(rng.integers(6, 18)), int(rng.integers(6, 18))
        amp = int(rng.integers(3, 14))
        img = np.empty((h, w, 4), np.uint8)
        img[..., :3] = np.clip(rng.integers(40, 200, size=3) + rng.integers(-amp, amp + 1, size=(h, w, 3)), 0, 255)
        img[..., 3] = 255
        shift, thres = int(rng.choice([0, 1, 2, 3])), int(rng.choice([1, 2, 3, 6]))
        J, M, _ = static_edges_colour(img, True, shift, thres, (0, 0, 0, 0), False)
        try:
            mod = run_model_fixpoint(J, M, w, h, 1)
        except Unsupported:
            continue
        if not mod["stale_set"]:
            continue
        cfg = make_config(hierarchical=[0, 0xFFFFFFFF][done % 2], diagonal=1, is_same_color_a=shift, is_same_color_b=thres,
                          good_min_area=2, deepen_diff=8)
        ref = oracle.runner_run(img, cfg, pools=True, color_image=True)
        got = sim.runner_run(img, cfg, pools=True, color_image=True)
        lc = sim.last_counters()
        assert lc["resurrected_leaves"] == len(mod["stale_set"]) and lc["stage1_passes"] == mod["iterations"]
        assert_same_clusters(got, ref, pools=True, ctx=f"stale case {k}")
        done += 1
        if done == 4:
            break
    assert done == 4


def test_sim_unclean_key_sequential_resolver(sim, oracle):
    """a non-key pixel color_same to the key: label 0 joins, merges, is emptied and refilled.  The
    device replays such a batch with the sequential resolver, the hierarchical stage on top included (under
    KeyingAction::Discard the discarded pixels get the INERT stage-1 label: label 0 in the result, members of nothing)."""
    from quirk_cases import unclean_key_case

    rng = np.random.default_rng(606)
    took = 0
    for k in range(40):
        hier = [0, 0, 0xFFFFFFFF, 48][k % 4]
        img, cfg = unclean_key_case(rng, max_side=16, hierarchical=hier)
        ref = oracle.runner_run(img, cfg, pools=True, color_image=True)
        got = sim.runner_run(img, cfg, pools=True, color_image=True)
        took += sim.last_counters()["sequential"]
        assert_same_clusters(got, ref, pools=True, ctx=f"unclean case {k}")
    assert took >= 30
    # a batch of two images, one of them unclean
    img, cfg = unclean_key_case(rng, max_side=12, hierarchical=0, diagonal=0)
    clean = img.copy()
    clean[...] = (3, 3, 3, 255)
    got = sim.runner_run_batch([clean, img], cfg, pools=True)
    for g, im in zip(got, [clean, img]):
        assert_same_clusters(g, oracle.runner_run(im, cfg, pools=True), pools=True, ctx="unclean batch")
    # hierarchical stage + KeyingAction::Discard + unclean key: the discarded pixels are label 0 in the result but members of nothing
    img = np.full((4, 4, 4), 100, dtype=np.uint8)
    img[1, 1] = (101, 100, 100, 255)
    from visioncortex_b200.runtime import make_config
    cfg = make_config(hierarchical=0xFFFFFFFF, is_same_color_a=0, is_same_color_b=2, key_color=(101, 100, 100, 255), keying_action=1)
    assert_same_clusters(sim.runner_run(img, cfg, pools=True, color_image=True), oracle.runner_run(img, cfg, pools=True, color_image=True), pools=True, ctx="discard 4x4")


def test_sim_empty_colour_image(sim, oracle):
    from visioncortex_b200.runtime import make_config

    for shape in [(0, 0, 4), (0, 5, 4)]:
        img = np.zeros(shape, dtype=np.uint8)
        cfg = make_config(hierarchical=0)
        assert_same_clusters(sim.runner_run(img, cfg, pools=True), oracle.runner_run(img, cfg, pools=True), pools=True)


def test_sim_hierarchical_and_banded(sim, oracle):
    """stage 2 (exact pop order) and the row-band driver (device set emulated on one simulated device)."""
    from sim_lib import load_sim  # noqa: F401
    from visioncortex_b200.runtime import Context, make_config, runner_run_banded

    rng = np.random.default_rng(33)
    ctxs = [sim, Context(0, lib=sim.lib)]
    for k in range(4):
        img, cfg = random_case(rng, max_side=26, keyed=(k == 2), hierarchical=[0xFFFFFFFF, 64][k % 2])
        ref = oracle.runner_run(img, cfg, pools=True, color_image=True)
        got = sim.runner_run(img, cfg, pools=True, color_image=True)
        assert_same_clusters(got, ref, pools=True, ctx=f"hier {k}")
        got2 = runner_run_banded(ctxs, img, cfg, pools=False)
        assert_same_clusters(got2, ref, pools=False, ctx=f"banded {k}")


def test_sim_rag_row_sweep_shapes(sim, oracle):
    """k_rag_rows (widths that are a multiple of 4): one and several 128-column strips, partial strips, heights around the
    32-row strip height, a batch of two images -- the hierarchical result depends on every adjacency / boundary record."""
    from parity import random_colour_image
    from visioncortex_b200.runtime import make_config

    rng = np.random.default_rng(77)
    for (w, h) in [(4, 1), (8, 33), (128, 5), (132, 31), (256, 3), (260, 34), (12, 65)]:
        img = random_colour_image(rng, w, h)
        cfg = make_config(hierarchical=0xFFFFFFFF, is_same_color_a=int(rng.choice([2, 4])), is_same_color_b=1, good_min_area=2)
        ref = oracle.runner_run(img, cfg, pools=True, color_image=True)
        got = sim.runner_run(img, cfg, pools=True, color_image=True)
        assert_same_clusters(got, ref, pools=True, ctx=f"rag rows {w}x{h}")
    imgs = [random_colour_image(rng, 136, 9) for _ in range(2)]
    cfg = make_config(hierarchical=0xFFFFFFFF, is_same_color_a=4, is_same_color_b=1)
    got = sim.runner_run_batch(imgs, cfg, pools=True)
    for k, im in enumerate(imgs):
        assert_same_clusters(got[k], oracle.runner_run(im, cfg, pools=True), pools=True, ctx=f"rag rows batch {k}")


def test_sim_batch_matches_single(sim, oracle):
    """several images in one launch sequence: per-image slot ranges, list offsets and the fused NEW / leaf compaction
    across image boundaries (one image is keyed away completely and owns a single slot)."""
    from visioncortex_b200.runtime import make_config

    rng = np.random.default_rng(44)
    for hier in (0, 0xFFFFFFFF):
        imgs = [rng.integers(0, 4, size=(13, 18, 4), dtype=np.uint8) * 60 for _ in range(4)]
        imgs[2][:] = (0, 0, 0, 0)
        imgs[2][..., 3] = 0
        cfg = make_config(hierarchical=hier, is_same_color_a=0, is_same_color_b=1)
        got = sim.runner_run_batch(imgs, cfg, pools=True)
        for k, im in enumerate(imgs):
            ref = oracle.runner_run(im, cfg, pools=True)
            assert_same_clusters(got[k], ref, pools=True, ctx=f"hier {hier} image {k}")


def test_sim_colour_threshold_extremes(sim, oracle):
    """color_same (runner.rs:116-129) at the ends of its parameter range: the byte-wise SIMD restatement must agree for a
    negative threshold (nothing is the same colour), one above 255 (everything is) and every shift."""
    from visioncortex_b200.runtime import make_config

    rng = np.random.default_rng(45)
    img = rng.integers(0, 256, size=(11, 14, 4), dtype=np.uint8)
    for shift, thres in [(0, -1), (0, 300), (0, 255), (0, 254), (7, 0), (7, 1), (3, 5), (1, 127), (0, 0)]:
        cfg = make_config(hierarchical=0, is_same_color_a=shift, is_same_color_b=thres)
        ref = oracle.runner_run(img, cfg, pools=True)
        got = sim.runner_run(img, cfg, pools=True)
        assert_same_clusters(got, ref, pools=True, ctx=f"shift {shift} thres {thres}")


def test_sim_chunked_batches_overlap_stage2(sim, oracle, monkeypatch):
    """a hierarchical batch split into device batches whose stage 2 runs on rotating workspace sets (two sets, chunks of two
    images: the third chunk has to complete the first one before it reuses its set), and the fused kernel's global-memory key path"""
    from visioncortex_b200.runtime import make_config, runner_run_batch_host

    rng = np.random.default_rng(46)
    imgs = [rng.integers(0, 4, size=(12, 16, 4), dtype=np.uint8) * 60 for _ in range(7)]
    cfg = make_config(is_same_color_a=0, is_same_color_b=1, good_min_area=2, deepen_diff=16)
    refs = [oracle.runner_run(im, cfg, pools=True) for im in imgs]
    monkeypatch.setenv("VCB_CHUNK_IMAGES", "2")
    monkeypatch.setenv("VCB_S2_SETS", "2")
    for keycap in ("2048", "3"):
        monkeypatch.setenv("VCB_S2_KEYCAP", keycap)
        got = sim.runner_run_batch(imgs, cfg, pools=True)
        for k in range(len(imgs)):
            assert_same_clusters(got[k], refs[k], pools=True, ctx=f"keycap {keycap} image {k}")
    monkeypatch.setenv("VCB_HOST_CHUNKS", "4")
    got = runner_run_batch_host(sim, imgs, cfg)
    for k in range(len(imgs)):
        assert_same_clusters(got[k], refs[k], pools=False, ctx=f"host batch image {k}")
    # compact records (vcb_host_result.live_slots): only the slots that are not Cluster::default() travel
    got = runner_run_batch_host(sim, imgs, cfg, compact=True)
    for k in range(len(imgs)):
        assert got[k]["compact"] == 1 and 0 < got[k]["num_live"] <= got[k]["num_slots"]
        assert_same_clusters(got[k], refs[k], pools=False, ctx=f"compact host batch image {k}")


def test_sim_stage2_with_gpu_block_sizes(sim, oracle, monkeypatch):
    """the hierarchical stage with the block sizes the GPU launches (256 and 512 threads: four / two passes of the 32 candidates
    of a round; the simulator's default is 128) and with the single-pop path only"""
    rng = np.random.default_rng(57)
    cases = [random_case(rng, max_side=22, keyed=(k == 1), hierarchical=0xFFFFFFFF) for k in range(3)]
    for img, cfg in cases:
        cfg.good_min_area = 2
        cfg.deepen_diff = 8
    refs = [oracle.runner_run(img, cfg, pools=True, color_image=True) for img, cfg in cases]
    for threads, nobatch in (("256", None), ("512", None), ("256", "1")):
        monkeypatch.setenv("VCB_S2_THREADS", threads)
        if nobatch:
            monkeypatch.setenv("VCB_S2_NOBATCH", nobatch)
        for (img, cfg), ref in zip(cases, refs):
            got = sim.runner_run(img, cfg, pools=True, color_image=True)
            assert_same_clusters(got, ref, pools=True, ctx=f"threads {threads} nobatch {nobatch} {img.shape}")
