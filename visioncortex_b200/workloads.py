"""The BASELINE.json configurations as concrete inputs and RunnerConfigs.
Shared by bench.py, the tests and the golden-digest generator so that all of them run exactly the same workloads."""
from __future__ import annotations

from . import synth
from .runtime import make_config

HM = 0xFFFFFFFF
C3_KEY = (40, 200, 120, 255)


def c1_input():
    """C1: BinaryImage::to_clusters on G1(1024, 1024, p = 0.70, seed = 1)."""
    return synth.g1_binary(1024, 1024, 0.70, 1)


def c2_input(seed: int = 2):
    return synth.g2_scene(1920, 1080, seed=seed)


def c2_config(w: int = 1920, h: int = 1080):
    """C2: vtracer-default parameters, hierarchical off."""
    return make_config(diagonal=0, hierarchical=0, batch_size=25600, good_min_area=16, good_max_area=w * h,
                       is_same_color_a=2, is_same_color_b=1, deepen_diff=16, hollow_neighbours=1)


def c3_input():
    return synth.g3_posterised(3840, 2160, seed=3, key=C3_KEY)


def c3_config_pass1(w: int = 3840, h: int = 2160):
    """C3 pass 1: hierarchical MAX + deepen_diff + key colour (Keep)."""
    return make_config(diagonal=0, hierarchical=HM, good_min_area=16, good_max_area=w * h, is_same_color_a=2,
                       is_same_color_b=1, deepen_diff=16, hollow_neighbours=1, key_color=C3_KEY, keying_action=0)


def c3_config_pass2(w: int = 3840, h: int = 2160):
    """C3 pass 2 on the rendered pass-1 image: hierarchical 64, Discard keying."""
    return make_config(diagonal=0, hierarchical=64, good_min_area=0, good_max_area=w * h, is_same_color_a=0,
                       is_same_color_b=1, deepen_diff=0, hollow_neighbours=0, key_color=C3_KEY, keying_action=1)


def c4_input(k: int):
    """C4: image k of the 512-image batch, G2(1920, 1080, seed = 100 + k)."""
    return synth.g2_scene(1920, 1080, seed=100 + k)


def c4_config():
    """C4 / C5: RunnerConfig::default() (color_clusters/runner.rs:23-39)."""
    return make_config()


def c5_input(side: int = 16384):
    return synth.g3_posterised(side, side, seed=5, cell=256)


c5_config = c4_config
