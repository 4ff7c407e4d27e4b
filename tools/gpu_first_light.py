#!/usr/bin/env python
"""GPU bring-up report: parity of libspb200 against the CPU checkers plus first timings.
Run on a GPU box:  python tools/gpu_first_light.py [--full]   (writes gpurun_out/first_light.json)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ora  # noqa: E402
from vk_cinematic_b200 import sp, workloads as W  # noqa: E402


def image_stats(a, b):
    d = np.abs(a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64))
    rel = d / np.maximum(np.abs(b[..., :3]), 1e-3)
    bits = (a.view(np.uint32) != b.view(np.uint32)).any(axis=2)
    return {"pixels": int(bits.size), "bit_differing_pixels": int(bits.sum()),
            "max_abs": float(d.max()), "max_rel": float(rel.max()),
            "rmse": float(np.sqrt((d ** 2).mean())),
            "frac_rel_gt_1e-4": float((rel.max(axis=2) > 1e-4).mean()),
            "frac_rel_gt_1e-2": float((rel.max(axis=2) > 1e-2).mean()),
            "nan_a": int(np.isnan(a).sum()), "nan_b": int(np.isnan(b).sum())}


def main():
    full = "--full" in sys.argv
    out = {}
    libs = {}
    if ora.have_ref():
        libs["ref"] = ora.load_ref()
        libs["ref_dm"] = ora.load_ref_dm()
    if ora.have_port():
        libs["port"] = ora.load_port()
    print("checkers:", list(libs))

    # ---- C1-like: bunny, image parity
    w, h = (1024, 768) if full else (512, 384)
    wl = W.config1(w, h)
    r = sp.Renderer().load_workload(wl)
    for cull in (1, 0):
        sp.set_params(samplesPerPixel=1, bounceCount=3, cullByDistance=cull)
        t0 = time.time()
        img, m = r.render_frame(frame=0)
        dt = time.time() - t0
        st = sp.last_stats()
        key = f"C1_{w}x{h}_cull{cull}"
        out[key] = {"metrics": m.tolist(), "kernel_ms": st.kernelMs, "total_ms": st.totalMs,
                    "wall_s": dt, "mrays_s_kernel": m[2] / st.kernelMs / 1e3}
        print(key, out[key])
        gpu_img = img.copy()
        for name, L in libs.items():
            s = L.scene().load_workload(wl)
            cimg, cm = s.render_seeded(spp=1, bounces=3, frame=0)
            out[key + "_vs_" + name] = image_stats(gpu_img, cimg)
            out[key + "_vs_" + name]["metrics_equal"] = bool((cm[1:5] == m[1:5]).all())
            print("  vs", name, out[key + "_vs_" + name])
            s.close()
    # multi-sample
    sp.set_params(samplesPerPixel=4, bounceCount=3, cullByDistance=1)
    img, m = r.render_frame(frame=3)
    gpu_img = img.copy()
    for name, L in libs.items():
        s = L.scene().load_workload(wl)
        cimg, cm = s.render_seeded(spp=4, bounces=3, frame=3)
        out["C1_4spp_vs_" + name] = image_stats(gpu_img, cimg)
        print("4spp vs", name, out["C1_4spp_vs_" + name])
        s.close()
    r.close()

    # ---- C2: monkey primary hits
    w, h = (1920, 1080) if full else (960, 540)
    wl = W.config2(w, h)
    r = sp.Renderer().load_workload(wl)
    sp.lib.sp_b200_EnableStats(1)
    for cull in (1, 0):
        sp.set_params(samplesPerPixel=1, bounceCount=1, cullByDistance=cull)
        g = r.primary_hits()
        st = sp.last_stats()
        key = f"C2_{w}x{h}_cull{cull}"
        out[key] = {"kernel_ms": st.kernelMs, "rays": st.rays, "node_visits_per_ray": st.nodeVisits / max(1, st.rays),
                    "tri_tests_per_ray": st.triangleTests / max(1, st.rays)}
        print(key, out[key])
        for name, L in libs.items():
            if name == "ref_dm":
                continue
            s = L.scene().load_workload(wl)
            c = s.primary_hits()
            mism = (c["tri"] != g["tri"])
            tb = (c["t"].view(np.uint32) != g["t"].view(np.uint32))
            out[key + "_vs_" + name] = {"tri_mismatch": int(mism.sum()), "rate": float(mism.mean()),
                                        "t_bit_mismatch": int(tb.sum()), "hits": int((c["tri"] >= 0).sum())}
            print("  vs", name, out[key + "_vs_" + name])
            s.close()
    sp.lib.sp_b200_EnableStats(0)
    sp.set_params(cullByDistance=1)
    for _ in range(3):
        g = r.primary_hits()
    st = sp.last_stats()
    out["C2_timing"] = {"kernel_ms": st.kernelMs, "mrays_s": w * h / st.kernelMs / 1e3}
    print("C2 timing", out["C2_timing"])
    r.close()

    # ---- multi-object scene
    wl = W.multi_object_workload()
    r = sp.Renderer().load_workload(wl)
    sp.set_params(samplesPerPixel=2, bounceCount=3, cullByDistance=1)
    img, m = r.render_frame(frame=1)
    gpu_img = img.copy()
    g = r.primary_hits()
    for name, L in libs.items():
        s = L.scene().load_workload(wl)
        cimg, cm = s.render_seeded(spp=2, bounces=3, frame=1)
        out["multi_vs_" + name] = image_stats(gpu_img, cimg)
        c = s.primary_hits()
        out["multi_vs_" + name]["tri_mismatch"] = int((c["tri"] != g["tri"]).sum())
        out["multi_vs_" + name]["obj_mismatch"] = int((c["obj"] != g["obj"]).sum())
        print("multi vs", name, out["multi_vs_" + name])
        s.close()
    r.close()

    # ---- timing at C3 shape, few samples
    w, h, spp = (3840, 2160, 8) if full else (1920, 1080, 4)
    wl = W.config3(w, h, spp=spp, bounces=5)
    r = sp.Renderer().load_workload(wl)
    for math in (0, 1):
        sp.set_params(samplesPerPixel=spp, bounceCount=5, cullByDistance=1, mathMode=math)
        for it in range(2):
            m, _ = r.render_rows(0, h, frame=it, host=False)
        st = sp.last_stats()
        key = f"C3_{w}x{h}_{spp}spp_math{math}"
        out[key] = {"kernel_ms": st.kernelMs, "rays": int(m[2]), "mrays_s": m[2] / st.kernelMs / 1e3}
        print(key, out[key])
    sp.set_params(mathMode=0)
    sp.lib.sp_b200_EnableStats(1)
    m, _ = r.render_rows(0, h, frame=0, host=False)
    st = sp.last_stats()
    out["C3_stats"] = {"node_visits_per_ray": st.nodeVisits / st.rays, "tri_tests_per_ray": st.triangleTests / st.rays,
                       "env_clamped": st.envClampedLookups, "kernel_ms_stats": st.kernelMs}
    print("C3 stats", out["C3_stats"])
    sp.lib.sp_b200_EnableStats(0)
    r.close()

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "first_light.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
