"""Row-band mode on several GPUs of one box (BASELINE config 5 path): one image, workspace striped over the devices,
per-pixel kernels per band, border label equivalences resolved through peer memory.  Bit-exact vs the CPU oracle."""
import numpy as np
import pytest

from parity import assert_same_clusters
from visioncortex_b200 import synth
from visioncortex_b200.runtime import Context, make_config, runner_run_banded

pytestmark = pytest.mark.gpu


def _contexts(n):
    import torch

    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    return [Context(k) for k in range(n)]


@pytest.mark.parametrize("ngpu", [2, 4, 8])
def test_banded_matches_oracle(oracle, ngpu):
    ctxs = _contexts(ngpu)
    cases = [
        (synth.g2_scene(1920, 1080, seed=2), make_config(hierarchical=0, is_same_color_a=2, deepen_diff=16, good_max_area=1920 * 1080)),
        (synth.g3_posterised(4096, 4096, seed=5, cell=256), make_config()),
        (synth.g2_scene(2500, 1700, seed=11), make_config(is_same_color_a=3)),        # width not a multiple of 4: plain-load tiles
    ]
    for k, (img, cfg) in enumerate(cases):
        ref = oracle.runner_run(img, cfg, pools=False, color_image=True)
        got = runner_run_banded(ctxs, img, cfg, pools=False, color_image=True)
        assert_same_clusters(got, ref, pools=False, ctx=f"banded {ngpu} GPUs case {k}")
    for c in ctxs:
        c.close()
