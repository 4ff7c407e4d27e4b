"""Diagnostic (GPU box): C5-shaped scene, GPU vs the port checker under every scheduler / cull
combination; prints how many pixels differ and where.  Not a test; a debugging aid."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ora  # noqa: E402
from vk_cinematic_b200 import sp, workloads as W  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def main():
    assert sp.lib.sp_b200_Init(0) == 0
    wl = W.config5(480, 270, spp=2, bounces=5, env_size=(512, 256))
    chk = ora.load_port_dm().scene().load_workload(wl)
    cimg, cm = chk.render_seeded(spp=2, bounces=5, frame=4)
    e = chk.primary_hits()
    r = sp.Renderer().load_workload(wl)
    for mode in (0, 1):
        for cull in (1, 0):
            sp.set_params(samplesPerPixel=2, bounceCount=5, radianceClamp=10.0, envFilter=0, mathMode=0,
                          cullByDistance=cull, tileWidth=64, tileHeight=64, renderMode=mode, samplesPerPass=0)
            img, m = r.render_frame(frame=4)
            diff = np.any(bits(img) != bits(cimg), axis=2)
            ys, xs = np.nonzero(diff)
            print(f"mode {mode} cull {cull}: differing pixels {diff.sum()} metrics gpu {m[1:5]} chk {cm[1:5]}")
            for y, x in list(zip(ys, xs))[:8]:
                print("   ", x, y, img[y, x], cimg[y, x])
            g = r.primary_hits()
            print("    primary: obj mismatches", int((g["obj"] != e["obj"]).sum()), "tri", int((g["tri"] != e["tri"]).sum()),
                  "t", int((bits(g["t"]) != bits(e["t"])).sum()))
    chk.close()
    r.close()


if __name__ == "__main__":
    main()
