"""The torch (device-side) synthetic generators used by bench.py are bit-identical to the numpy ones."""
import numpy as np
import torch

from visioncortex_b200 import synth


def test_torch_generators_match_numpy():
    dev = torch.device("cpu")
    for (w, h, seed) in [(97, 61, 2), (200, 130, 100), (48, 48, 7), (1, 1, 3)]:
        assert np.array_equal(synth.g2_scene_torch(w, h, seed, dev).numpy(), synth.g2_scene(w, h, seed))
        assert np.array_equal(synth.g3_posterised_torch(w, h, seed, dev, cell=32).numpy(), synth.g3_posterised(w, h, seed, cell=32))
        key = (40, 200, 120, 255)
        assert np.array_equal(synth.g3_posterised_torch(w, h, seed, dev, cell=16, key=key).numpy(), synth.g3_posterised(w, h, seed, cell=16, key=key))
        assert np.array_equal(synth.g1_binary_torch(w, h, 0.7, seed, dev).numpy(), synth.g1_binary(w, h, 0.7, seed))


def test_torch_generator_full_size_row_sample():
    """one 1080p scene: every 97th row equal (keeps the CPU test fast while covering real coordinates)"""
    a = synth.g2_scene_torch(1920, 1080, 100, torch.device("cpu")).numpy()
    b = synth.g2_scene(1920, 1080, 100)
    assert np.array_equal(a, b)
