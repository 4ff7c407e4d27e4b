#!/usr/bin/env python
"""One launch of each environment-baking kernel at the reference's sizes (main.cpp:1307-1315), for
ncu:  ncu --set full --clock-control none --import-source on -k regex:'k_cube_map|k_irradiance' \
          -f -o gpurun_out/prof_cubemap python tools/cubemap_profile.py
No torch import in that mode (start-up time counts on the GPU box).

    python tools/cubemap_profile.py --time   # wall time per call, faces left on the device, both math modes"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vk_cinematic_b200 import sp, workloads as W  # noqa: E402

assert sp.lib.sp_b200_Init(0) == 0
env = W.make_env_map(4096, 2048)

if "--time" in sys.argv:
    import ctypes as C
    import json
    import time
    import numpy as np
    import torch
    img = np.ascontiguousarray(env, dtype=np.float32)
    hdr = sp.HdrImage(img.ctypes.data_as(C.POINTER(C.c_float)), img.shape[1], img.shape[0])
    faces = torch.empty(6 * 1024 * 1024 * 4, dtype=torch.float32, device="cuda")
    out = {"what": "launch + kernel + synchronize per call (best of 10), 4096x2048 map resident, faces left on the device"}
    for mode, tag in ((0, "f64_rounded"), (1, "fast_f32")):
        sp.set_params(mathMode=mode)
        for name, call in (
                ("cube_map_6x1024x1024", lambda: sp.lib.sp_b200_CreateCubeMap(C.byref(hdr), 1024, 1024, None, faces.data_ptr())),
                ("irradiance_uniform_6x32x32", lambda: sp.lib.sp_b200_CreateIrradianceCubeMap(C.byref(hdr), 32, 32, 32, 0, 0.1, None, faces.data_ptr())),
                ("irradiance_random_6x32x32_spp32", lambda: sp.lib.sp_b200_CreateIrradianceCubeMap(C.byref(hdr), 32, 32, 32, 1, 0.1, None, faces.data_ptr()))):
            call()
            best = 1e9
            for _ in range(10):
                t0 = time.perf_counter()
                call()
                best = min(best, time.perf_counter() - t0)
            out["%s_%s_ms" % (name, tag)] = round(best * 1e3, 4)
    sp.set_params(mathMode=0)
    print(json.dumps(out))
    sys.exit(0)

cube = sp.create_cube_map(env, 1024, 1024)
uniform = sp.create_irradiance_cube_map(env, 32, 32)
rnd = sp.create_irradiance_cube_map(env, 32, 32, spp=32, sampling=sp.IRRADIANCE_RANDOM)
print("cube mean %.6f  irradiance uniform mean %.6f  random mean %.6f" %
      (cube[..., :3].mean(), uniform[..., :3].mean(), rnd[..., :3].mean()))
