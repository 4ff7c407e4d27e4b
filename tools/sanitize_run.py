"""Small renders that touch every kernel family, for compute-sanitizer (SURVEY.md §5 prescribes
memcheck / racecheck on the path):

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
    compute-sanitizer --tool initcheck python tools/sanitize_run.py

  1. smoke scene (bunny, 192x128, 2 spp, 3 bounces): wavefront kernels (k_cover, k_list_blocks,
     k_sky, k_sky_listed, k_candidates, k_trace primary + bounce, k_shade_miss, k_shade_hit_tiles,
     k_shade_hit, k_accumulate), per-pixel kernel, k_primary_hits, sp_PathTraceTile (k_tiles_serial);
  2. C5 thumbnail (182 instances, 9.98 M instanced triangles, 160x90, 1 spp, 5 bounces) with the
     meshes built by the device LBVH builder (k_lbvh_keys / k_lbvh_nodes / k_lbvh_fit).
No result is checked here (the parity tests do that); the point is the sanitizer's report.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vk_cinematic_b200 import sp, workloads as W  # noqa: E402


def main():
    assert sp.lib.sp_b200_Init(0) == 0
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "smoke"):
        wl = W.config1(192, 128, env_size=(512, 256))
        r = sp.Renderer(0).load_workload(wl)
        sp.set_params(samplesPerPixel=2, bounceCount=3, cullByDistance=1, mathMode=0, envFilter=0, radianceClamp=10.0)
        img, m = r.render_frame(frame=1)
        print("wavefront rays", int(m[2]), flush=True)
        sp.set_params(samplesPerPixel=2, bounceCount=3, renderMode=1)
        img, m = r.render_frame(frame=1)
        print("per-pixel rays", int(m[2]), flush=True)
        sp.set_params(samplesPerPixel=2, bounceCount=3, renderMode=0, envFilter=1)
        img, m = r.render_frame(frame=2)
        hits = r.primary_hits(sample=0, frame=1)
        print("primary hits", int((hits["tri"] >= 0).sum()), flush=True)
        sp.set_params(samplesPerPixel=1, bounceCount=3, envFilter=0)
        r.path_trace_tile((16, 16, 48, 40), 0xF51C0E49)
        r.close()
    if which in ("all", "c5"):
        wl = W.config5(160, 90, spp=1, bounces=5, env_size=(256, 128))
        sp.lib.sp_b200_SetMeshBuilder(sp.BUILDER_DEVICE_LBVH)
        r = sp.Renderer(0).load_workload(wl)
        sp.lib.sp_b200_SetMeshBuilder(sp.BUILDER_HOST_SAH)
        sp.set_params(samplesPerPixel=1, bounceCount=5, cullByDistance=1, mathMode=0, envFilter=0, radianceClamp=10.0)
        img, m = r.render_frame(frame=3)
        print("c5 rays", int(m[2]), "finite", bool(np.isfinite(img).all()), flush=True)
        r.close()
    sp.lib.sp_b200_Shutdown()
    print("sanitize_run done", flush=True)


if __name__ == "__main__":
    main()
