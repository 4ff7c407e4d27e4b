#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libspref*.so).

Run HERE (where /root/reference is mounted and `make -C oracle ref` works); the outputs are
committed so that the GPU box -- which has no reference mount -- and any later checkout can pin
both the oracle port and the CUDA path against the reference's own outputs:

    python tools/make_golden.py

Every fixture records the workload parameters it was made with; tests rebuild the same inputs
from vk_cinematic_b200.workloads (seeded, deterministic) and compare.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ora  # noqa: E402
from vk_cinematic_b200 import workloads as W  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def golden_rays(n, seed, lo, hi):
    """Seeded rays through the scene's bounding region (segment p->q, like
    unit_tests/test_simd_path_tracer.cpp:496-527 builds its rays)."""
    rng = np.random.RandomState(seed)
    p = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    q = rng.uniform(lo * 0.3, hi * 0.3, (n, 3)).astype(np.float32)
    d = (q - p).astype(np.float32)
    ln = np.sqrt((d * d).sum(axis=1, dtype=np.float32)).astype(np.float32)
    d = (d / ln[:, None]).astype(np.float32)
    # a few axis-parallel rays: 1/0 = inf in Inverse(), the NaN lanes of the slab test
    d[0] = (0, 0, -1)
    p[0] = (0.01, 0.1, 3.0)
    d[1] = (1, 0, 0)
    p[1] = (-3.0, 0.1, 0.02)
    d[2] = (0, -1, 0)
    p[2] = (0.0, 3.0, 0.0)
    return p, d


def main():
    if not ora.have_ref():
        ora.build_oracles()
    ref, ref_dm = ora.load_ref(), ora.load_ref_dm()
    os.makedirs(OUT, exist_ok=True)

    # G1: bunny image, 96x64, 2 spp, frame 1, 512x256 env (glibc libm and deterministic math)
    wl = W.config1(96, 64, env_size=(512, 256))
    out = {}
    for name, L in (("ref", ref), ("ref_dm", ref_dm)):
        s = L.scene().load_workload(wl)
        img, m = s.render_seeded(spp=2, bounces=3, frame=1, threads=4)
        out["image_" + name] = img
        out["metrics_" + name] = m[1:5]
        if name == "ref":
            ph = s.primary_hits(sample=0, frame=1, threads=4)
            out["tri"], out["obj"], out["t"] = ph["tri"], ph["obj"], ph["t"]
        # one native tile: serial stream seeded 0xF51C0E49 (main.cpp:738-739)
        tile_img = np.zeros((64, 96, 4), np.float32)
        state, tm = s.path_trace_tile(tile_img, (16, 8, 48, 40), 2, 3, 0xF51C0E49)
        out["tile_image_" + name] = tile_img
        out["tile_state_" + name] = np.uint32(state)
        out["tile_metrics_" + name] = tm[1:5]
        s.close()
    np.savez_compressed(os.path.join(OUT, "g1_bunny_96x64.npz"), **out)

    # G2: monkey primary hits 160x90 (flat shading)
    wl = W.config2(160, 90, env_size=(64, 32))
    s = ref.scene().load_workload(wl)
    ph = s.primary_hits(sample=0, frame=0, threads=4)
    np.savez_compressed(os.path.join(OUT, "g2_monkey_160x90_primary.npz"), tri=ph["tri"],
                        obj=ph["obj"], t=ph["t"])
    s.close()

    # G3: multi-object scene: ray batch (sp_RayIntersectScene) + image
    wl = W.multi_object_workload(width=80, height=60, spp=2, env_size=(256, 128))
    p, d = golden_rays(3000, 1234, -4.0, 4.0)
    out = {"origins": p, "dirs": d}
    for name, L in (("ref", ref), ("ref_dm", ref_dm)):
        s = L.scene().load_workload(wl)
        if name == "ref":
            r = s.intersect_rays(p, d)
            for k in ("t", "material", "normal", "uv", "tri", "obj"):
                out["rays_" + k] = r[k]
        img, m = s.render_seeded(spp=2, bounces=3, frame=5, threads=4)
        out["image_" + name] = img
        out["metrics_" + name] = m[1:5]
        s.close()
    np.savez_compressed(os.path.join(OUT, "g3_multi_80x60.npz"), **out)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
