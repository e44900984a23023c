"""The reference-API mirror (Runner::start / IncrementalBuilder, Clusters / ClustersView / Cluster queries, BinaryImage) driven
through the C ABI.  CPU: the host-simulator build of the kernels; GPU: the product library."""
import numpy as np
import pytest

from api_checks import check_batch_queries, check_binary_api, check_builder_contract, check_runner_api
from parity import random_colour_image

HM = 0xFFFFFFFF
CASES = [dict(hierarchical=0, is_same_color_a=2), dict(hierarchical=HM, is_same_color_a=0, good_min_area=2, deepen_diff=16),
         dict(hierarchical=HM, is_same_color_a=2, key_color=(7, 250, 3, 255), keying_action=0),
         dict(hierarchical=64, is_same_color_a=4, key_color=(7, 250, 3, 255), keying_action=1, good_min_area=0)]


def _images(rng, n, side):
    out = []
    for k in range(n):
        img = random_colour_image(rng, int(rng.integers(2, side)), int(rng.integers(2, side)))
        if k % 2:
            img[rng.random(img.shape[:2]) < 0.2] = np.array((7, 250, 3, 255), dtype=np.uint8)
        out.append(img)
    return out


def test_api_on_simulator():
    from sim_lib import load_sim

    sim = load_sim()
    rng = np.random.default_rng(8)
    for k, img in enumerate(_images(rng, 4, 18)):
        check_runner_api(sim, img, **CASES[k % len(CASES)])
    # a wider image: cluster rects beyond 32 pixels put the batched components (row f1) into several size classes
    wide = random_colour_image(np.random.default_rng(12), 70, 37)
    check_runner_api(sim, wide, hierarchical=HM, is_same_color_a=5, good_min_area=2, deepen_diff=8)
    check_runner_api(sim, np.zeros((0, 0, 4), dtype=np.uint8), hierarchical=0)
    check_builder_contract(sim)
    check_binary_api(sim)
    check_batch_queries(sim, side=14)


@pytest.mark.gpu
def test_api_on_gpu(gpu):
    rng = np.random.default_rng(9)
    for k, img in enumerate(_images(rng, 12, 60)):
        check_runner_api(gpu, img, **CASES[k % len(CASES)])
    check_builder_contract(gpu)
    check_binary_api(gpu)
    check_batch_queries(gpu, side=150)
