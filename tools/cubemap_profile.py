#!/usr/bin/env python
"""One launch of each environment-baking kernel at the reference's sizes (main.cpp:1307-1315), for
ncu:  ncu --set full --clock-control none --import-source on -k regex:'k_cube_map|k_irradiance' \
          -f -o gpurun_out/prof_cubemap python tools/cubemap_profile.py
No torch import (start-up time counts on the GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vk_cinematic_b200 import sp, workloads as W  # noqa: E402

assert sp.lib.sp_b200_Init(0) == 0
env = W.make_env_map(4096, 2048)
cube = sp.create_cube_map(env, 1024, 1024)
uniform = sp.create_irradiance_cube_map(env, 32, 32)
rnd = sp.create_irradiance_cube_map(env, 32, 32, spp=32, sampling=sp.IRRADIANCE_RANDOM)
print("cube mean %.6f  irradiance uniform mean %.6f  random mean %.6f" %
      (cube[..., :3].mean(), uniform[..., :3].mean(), rnd[..., :3].mean()))
