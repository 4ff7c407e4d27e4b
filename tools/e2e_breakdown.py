#!/usr/bin/env python
"""Where the end-to-end step of bench.py spends its time at N=1 (development aid, GPU box)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vk_cinematic_b200 import sp, workloads as W

wl = W.config3(3840, 2160, spp=64, bounces=5)
env_pinned = torch.from_numpy(wl.textures[W.IMAGE_ENV]).pin_memory()
wl.textures[W.IMAGE_ENV] = env_pinned.numpy()
host_image = torch.zeros((wl.height, wl.width, 4), dtype=torch.float32).pin_memory()
assert sp.lib.sp_b200_Init(0) == 0
r = sp.Renderer(0).load_workload(wl, pixels=host_image.numpy())
sp.set_params(samplesPerPixel=64, bounceCount=5, cullByDistance=1, mathMode=0, envFilter=0, radianceClamp=10.0,
              tileWidth=64, tileHeight=64, renderMode=0, samplesPerPass=0)
def T(f, n=3):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    return min(ts)
frame = [0]
def render_host():
    r.render_frame(frame=frame[0]); frame[0] += 1
def render_dev():
    r.render_rows(0, wl.height, frame=frame[0], host=False); frame[0] += 1
render_host(); render_dev()
print("render_frame (D2H incl.)   %.2f ms" % T(render_host))
print("render_rows device only    %.2f ms" % T(render_dev))
st = sp.last_stats(); print("  kernelMs %.2f totalMs %.2f" % (st.kernelMs, st.totalMs))
print("flush textures             %.2f ms" % T(lambda: sp.lib.sp_b200_FlushTextureCache()))
def flush_render():
    sp.lib.sp_b200_FlushTextureCache(); render_dev()
print("flush + render device      %.2f ms" % T(flush_render))
print("build                      %.2f ms" % T(lambda: r.build()))
def full():
    sp.lib.sp_b200_FlushTextureCache(); r.build(); render_host()
print("flush + build + render     %.2f ms" % T(full))
a = torch.empty(wl.height, wl.width, 4, device="cuda")
print("torch D2H 133 MB pinned    %.2f ms" % T(lambda: host_image.copy_(a)))
print("torch H2D 134 MB pinned    %.2f ms" % T(lambda: torch.empty_like(env_pinned, device="cuda").copy_(env_pinned)))
# the bench's loop: consecutive steps, wall clock
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    for k in range(5):
        t1 = time.perf_counter(); sp.lib.sp_b200_FlushTextureCache(); t2 = time.perf_counter(); r.build(); t3 = time.perf_counter(); render_host(); t4 = time.perf_counter()
        print("   step %d: flush %.2f build %.2f render %.2f ms" % (k, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
    torch.cuda.synchronize()
    print("5 consecutive e2e steps: %.2f ms/step" % ((time.perf_counter() - t) * 1e3 / 5))
