"""Inputs that hit the reference's order-dependent quirks, shared by the simulator and the
GPU parity tests: the device replays them with the sequential resolver (csrc/sequential.cuh)."""
import numpy as np

from visioncortex_b200.runtime import make_config


def unclean_key_case(rng, max_side=20, diagonal=None, hierarchical=0, keying_action=None):
    """near-threshold noise around one colour; the key is the colour of one of the pixels and is planted on a fifth of the
    image, so non-key pixels that are color_same to the key touch keyed pixels everywhere"""
    w, h = int(rng.integers(3, max_side + 1)), int(rng.integers(3, max_side + 1))
    amp = int(rng.integers(1, 6))
    img = np.empty((h, w, 4), np.uint8)
    img[..., :3] = np.clip(rng.integers(60, 180, size=3) + rng.integers(-amp, amp + 1, size=(h, w, 3)), 0, 255)
    img[..., 3] = 255
    key = tuple(int(v) for v in img[int(rng.integers(0, h)), int(rng.integers(0, w))])
    img[rng.random((h, w)) < 0.2] = np.array(key, dtype=np.uint8)
    cfg = make_config(hierarchical=hierarchical, diagonal=int(rng.integers(0, 2)) if diagonal is None else int(diagonal),
                      is_same_color_a=int(rng.choice([0, 1, 2])), is_same_color_b=int(rng.choice([1, 2, 3])),
                      key_color=key, keying_action=int(rng.integers(0, 2)) if keying_action is None else keying_action,
                      good_min_area=int(rng.choice([0, 2])), deepen_diff=int(rng.choice([0, 4, 64])))
    return img, cfg
