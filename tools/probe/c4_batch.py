import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from visioncortex_b200 import synth
from visioncortex_b200.runtime import Context, make_config
ctx = Context(0)
imgs = [synth.g2_scene(1920, 1080, seed=100 + k) for k in range(4)]
for B in [int(x) for x in sys.argv[1:]]:
    dev = torch.from_numpy(np.stack([imgs[k % 4] for k in range(B)])).cuda()
    torch.cuda.synchronize()
    try:
        hs = ctx.runner_run_device(dev.data_ptr(), 1920, 1080, B, make_config())
        ctx.free_handles(hs); ctx.synchronize()
        print("B", B, "ok", flush=True)
    except Exception as e:
        print("B", B, "FAIL", e, flush=True); break
