import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from visioncortex_b200 import synth, api
from visioncortex_b200.runtime import Context
ctx = Context(0)
img = synth.g2_scene(1920, 1080, seed=100)
for cfg in (api.RunnerConfig(), api.RunnerConfig(hierarchical=0, is_same_color_a=2, good_max_area=1920 * 1080)):
    cl = api.Runner(cfg, api.ColorImage.from_array(img), ctx=ctx).run()
    lib = ctx.lib
    for rep in range(3):
        hnd = C.c_void_p()
        ctx.profile_reset(True)
        t0 = time.time()
        lib.check(lib.fn("clusters_components")(cl._handle, 1, C.byref(hnd)), ctx.handle)
        t1 = time.time()
        ms, n = ctx.profile_read("")
        ctx.profile_reset(False)
        lib.fn("components_free")(hnd)
        t2 = time.time()
        comps = cl.device_components(True)
        t3 = time.time()
        print(f"outputs {cl.output_len()}: C call {1e3*(t1-t0):.1f} ms (kernel time {ms:.1f} ms in {n} launches), python wrapper total {1e3*(t3-t2):.1f} ms", flush=True)
