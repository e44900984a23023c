"""GPU parity tests: the CUDA product library, called through the C ABI, against the CPU oracle on the same inputs.
Bit-exact: cluster_indices, slot count, every record field, clusters_output order."""
import numpy as np
import pytest

from parity import assert_same_clusters, random_case
from visioncortex_b200 import synth
from visioncortex_b200.runtime import make_config

pytestmark = pytest.mark.gpu


def c2_config(w, h, **kw):
    """BASELINE config 2: vtracer-default parameters, hierarchical off."""
    base = dict(diagonal=0, hierarchical=0, batch_size=25600, good_min_area=16, good_max_area=w * h,
                is_same_color_a=2, is_same_color_b=1, deepen_diff=16, hollow_neighbours=1)
    base.update(kw)
    return make_config(**base)


def test_random_small_colour(gpu, oracle):
    rng = np.random.default_rng(11)
    n = 0
    for k in range(150):
        img, cfg = random_case(rng, max_side=48, keyed=(k % 3 == 2))
        ref = oracle.runner_run(img, cfg, pools=False)
        got = gpu.runner_run(img, cfg, pools=False)
        assert_same_clusters(got, ref, pools=False, ctx=f"case {k} {img.shape}")
        n += 1
    assert n == 150


def test_random_medium_colour(gpu, oracle):
    rng = np.random.default_rng(12)
    for k in range(20):
        img, cfg = random_case(rng, max_side=300, keyed=(k % 4 == 3))
        ref = oracle.runner_run(img, cfg, pools=False)
        got = gpu.runner_run(img, cfg, pools=False)
        assert_same_clusters(got, ref, pools=False, ctx=f"case {k} {img.shape}")


@pytest.mark.parametrize("seed", [2, 100])
def test_c2_full_size(gpu, oracle, seed):
    img = synth.g2_scene(1920, 1080, seed=seed)
    cfg = c2_config(1920, 1080)
    ref = oracle.runner_run(img, cfg, pools=False, color_image=True)
    got = gpu.runner_run(img, cfg, pools=False, color_image=True)
    assert_same_clusters(got, ref, pools=False, ctx=f"C2 seed {seed}")


def test_c3_pass1_stage1_only(gpu, oracle):
    """4K posterised image with a clean key colour; stage 1 + stage_1_output only (hierarchical = 0)."""
    key = (40, 200, 120, 255)
    img = synth.g3_posterised(3840, 2160, seed=3, key=key)
    for action in (0, 1):
        cfg = c2_config(3840, 2160, key_color=key, keying_action=action)
        ref = oracle.runner_run(img, cfg, pools=False)
        got = gpu.runner_run(img, cfg, pools=False)
        assert_same_clusters(got, ref, pools=False, ctx=f"C3 stage1 action {action}")


def test_batch_matches_single(gpu, oracle):
    imgs = [synth.g2_scene(640, 360, seed=100 + k) for k in range(6)]
    cfg = c2_config(640, 360)
    got = gpu.runner_run_batch(imgs, cfg)
    for k, im in enumerate(imgs):
        ref = oracle.runner_run(im, cfg, pools=False)
        assert_same_clusters(got[k], ref, pools=False, ctx=f"batch image {k}")


def test_deterministic(gpu):
    img = synth.g2_scene(800, 600, seed=5)
    cfg = c2_config(800, 600)
    a = gpu.runner_run(img, cfg, pools=False)
    b = gpu.runner_run(img, cfg, pools=False)
    assert_same_clusters(a, b, pools=False, ctx="run twice")


def test_edge_shapes(gpu, oracle):
    rng = np.random.default_rng(5)
    for (w, h) in [(1, 1), (1, 37), (41, 1), (2, 2), (64, 32), (65, 33), (63, 31), (129, 3), (3, 130)]:
        img = rng.integers(0, 4, size=(h, w, 4)).astype(np.uint8) * 60
        img[..., 3] = 255
        cfg = c2_config(w, h)
        ref = oracle.runner_run(img, cfg, pools=False)
        got = gpu.runner_run(img, cfg, pools=False)
        assert_same_clusters(got, ref, pools=False, ctx=f"shape {w}x{h}")
