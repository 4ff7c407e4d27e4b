"""tools/warpsim.py -- drives tools/warpsim.cpp (lane-occupancy model of k_trace's warp loop; development tool).

    python tools/warpsim.py [--block-step 40] [--spp 64] [--bounces 5] [--policy name=v,...] ...
"""
import argparse
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
LIB = os.path.join(ROOT, "tools", "_build", "libwarpsim.so")
CSRC = os.path.join(ROOT, "vk_cinematic_b200", "csrc")


def build():
    src = os.path.join(ROOT, "tools", "warpsim.cpp")
    deps = [src, os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp"), os.path.join(CSRC, "spb_core.cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(d) < os.path.getmtime(LIB) for d in deps):
        return
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-pthread",
                           "-Wno-unused-function", "-o", LIB, src, os.path.join(CSRC, "spb_capture.cpp"),
                           os.path.join(CSRC, "spb_bvh.cpp"), "-lm"])


KEYS = ["refill", "speculate", "early_lanes", "early_min_steps", "second_threshold", "vote_bias", "later_threshold", "bin_res", "pixel_major", "double_node"]
DEFAULT = {"refill": 1, "speculate": 0, "early_lanes": 0, "early_min_steps": 0, "second_threshold": 1, "vote_bias": 0,
           "later_threshold": 12, "bin_res": 16, "pixel_major": 0, "double_node": 0}


def run(scene, lib, policy, block_step, spp, bounces, weights=None, sort_later=0):
    p = dict(DEFAULT)
    p.update(policy)
    pol = np.array([p[k] for k in KEYS], np.uint32)
    out = np.zeros((8, 24), np.float64)
    hist = np.zeros(33, np.float64)
    wptr = None
    if weights is not None:
        weights = np.asarray(weights, np.float64)
        wptr = weights.ctypes.data_as(C.POINTER(C.c_double))
    lib.warpsim_run(scene.h, block_step, spp, bounces, 0, pol.ctypes.data_as(C.POINTER(C.c_uint32)), wptr,
                    out.ctypes.data_as(C.POINTER(C.c_double)), hist.ctypes.data_as(C.POINTER(C.c_double)), sort_later)
    return out, hist


def describe(out, bounces):
    lines = []
    total = 0.0
    for b in range(bounces - 1):
        o = out[b]
        if o[0] == 0:
            break
        lanes = o[3] / o[2]
        lines.append("  bounce %d: rays %8d hit %.3f | warp-inst/ray %7.1f lanes %5.2f | node it %8d lanes %5.2f | leaf it %8d lanes %5.2f"
                     " | steps/ray node %5.2f leaf %5.2f | lines/ray node %5.2f leaf %5.2f | parked %.3f (%.1f entries) second %.1f%%" % (
                         b + 1, o[0], o[1] / o[0], o[2] / o[0], lanes, o[4], o[5] / max(o[4], 1), o[6], o[7] / max(o[6], 1),
                         o[12] / o[0], o[13] / o[0], o[14] / o[0], o[15] / o[0], o[8] / o[0], o[9] / max(o[8], 1), 100 * o[10] / o[2]))
        if o[16]:
            lines.append("            object entries: %d iterations at %.2f lanes (%.2f per ray); exits: %d iterations at %.2f lanes" % (
                o[16], o[17] / o[16], o[17] / o[0], o[18], o[19] / max(o[18], 1)))
        total += o[2]
    return lines, total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--block-step", type=int, default=40)
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--bounces", type=int, default=5)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--policy", action="append", default=[])
    ap.add_argument("--hist", action="store_true")
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--order", type=int, default=0, help="bit 1: random order inside a direction bin; bit 2: no binning")
    args = ap.parse_args()
    build()
    import ora
    from vk_cinematic_b200 import workloads
    o = ora.OracleLib(LIB)
    o.lib.warpsim_run.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32),
                                  C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint32]
    if args.workload == "c5":
        wl = workloads.config5(width=args.width, height=args.height, spp=args.spp, bounces=args.bounces, env_size=(64, 32))
    else:
        wl = workloads.config3(width=args.width, height=args.height, spp=args.spp, bounces=args.bounces, env_size=(64, 32))
    scene = o.scene().load_workload(wl)
    policies = args.policy or [""]
    base = None
    for spec in policies:
        pol = {}
        for kv in filter(None, spec.split(",")):
            k, v = kv.split("=")
            pol[k] = int(v)
        out, hist = run(scene, o.lib, pol, args.block_step, args.spp, args.bounces, sort_later=args.order)
        lines, total = describe(out, args.bounces)
        base = base or total
        print("policy {%s}: total warp instructions %.4g (%.3f of first)" % (spec, total, total / base))
        print("\n".join(lines))
        if args.hist:
            h = hist / max(hist.sum(), 1)
            print("  walking-lane histogram of bounce 1 vote iterations:", " ".join("%d:%.3f" % (i, h[i]) for i in range(1, 33)))


if __name__ == "__main__":
    main()
