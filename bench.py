#!/usr/bin/env python
"""bench.py -- throughput of the sp_ path-tracing hot path on N B200s of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the one the metric's target is quoted on; it fits one GPU):
bunny.obj + a 4096x2048 equirect environment ("kiara" stand-in: the reference's EXR files are
git-LFS pointers), 3840x2160, 64 spp, 5 bounces.  One STEP = one frame of it: ray
generation, BVH traversal + triangle tests, material/BSDF evaluation, environment lookups and
the framebuffer write, with per-(pixel, sample, frame) XorShift32 streams.  Frame index advances
every step, so no step repeats another's rays.

Printed JSON (one line, rank 0):
  value    Mrays/s, whole job: rays traced by all ranks in the K timed steps / device time
           (CUDA events on the launch stream, max over ranks).  Scene, env map and BVH are
           resident in HBM before the timed region (the path uploads them once).
  e2e      same metric through the reference-facing C ABI with HOST buffers: every step
           re-uploads the scene (sp_BuildSceneBroadphase), re-uploads the environment map and
           material table (texture cache flushed), renders, and copies the finished RGBA-f32
           frame to pinned host memory.
  roofline dominant kernel (the render kernel): algorithmic bytes = rays x (128 I + 48 L + 64)
           + 16 B/pixel + shading bytes (DESIGN.md "Roofline"), I and L counted by a stats
           launch of the same frame; time = that kernel's CUDA-event duration.
  cpu_baseline  the CPU checker timed on this box's host cores on a bounded sample (rank 0, N=1).

--impl reference times the reference's CPU implementation on the host cores (all threads): the
port at the workload's 5 bounces (the reference's own sources are fixed at 3 bounces,
simd_path_tracer.cpp:195; their 3-bounce rate on the same sample is reported beside it).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RESULT_OUT = sys.stdout
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md "FALLBACK"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # development overrides (the defaults are the BASELINE configuration)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--bounces", type=int, default=5)
    ap.add_argument("--math", type=int, default=0, help="0: double-rounded libm stand-ins (parity mode), 1: CUDA f32")
    ap.add_argument("--workload", default="c3", choices=["c3", "c5", "c5u"],
                    help="c3: BASELINE configs[2] (default, the bench line); c5: configs[4] instanced stress scene; "
                         "c5u: the same with every sphere its own mesh (BVH larger than L2)")
    ap.add_argument("--render-mode", type=int, default=0, help="0: wavefront kernels (default), 1: per-pixel kernel")
    ap.add_argument("--samples-per-pass", type=int, default=0)
    ap.add_argument("--strip-rows", type=int, default=8, help="granularity of strip boundaries in pixel rows (multiple of 4)")
    ap.add_argument("--sort-bounces", type=int, default=0, help="development: direction-sort the rays leaving this many bounces")
    ap.add_argument("--refill", default="", help="development: refill thresholds primary,sorted,other")
    ap.add_argument("--evict", default="", help="development: straggler eviction thresholds sorted,other (0 = off)")
    ap.add_argument("--no-candidates", action="store_true", help="development: every primary ray walks the tree")
    ap.add_argument("--no-ray-sorting", action="store_true", help="development: bounce rays in hit-queue order")
    ap.add_argument("--paths-per-pass", type=int, default=0, help="development: paths in flight per pass (0 = library default)")
    ap.add_argument("--device-builder", action="store_true",
                    help="development: mesh trees from the device LBVH builder (sp_b200_SetMeshBuilder); default: host SAH")
    ap.add_argument("--quick", action="store_true", help="development: value only (no e2e, roofline, cpu baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the whole-frame parity block")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C5 secondary block")
    ap.add_argument("--no-copy-overlap", action="store_true", help="development: sp_b200_SetCopyOverlap(0)")
    ap.add_argument("--fuse-miss", type=int, default=-1, help="development: sp_b200_SetMissFusion(0 / 1)")
    ap.add_argument("--nccl-channels", type=int, default=0,
                    help="development: NCCL_MAX_NCHANNELS for this run (the env-map all-gather of the e2e steps runs beside the kernels)")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="development: one sp_b200_RenderRows call per step instead of RenderRowsBegin / End with two frames in flight")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload_config(args):
    name = ("C3: bunny.obj (4968 tris) + kiara-like 4096x2048 equirect env, "
            f"{args.width}x{args.height}, {args.spp} spp, {args.bounces} bounces (BASELINE.json configs[2])")
    if getattr(args, "workload", "c3") != "c3":
        name = (f"C5: 64 bunnies + 118 level-6 icospheres (9.98 M triangles, "
                f"{'unique sphere meshes' if args.workload == 'c5u' else 'instanced'}), "
                f"{args.width}x{args.height}, {args.spp} spp, {args.bounces} bounces (BASELINE.json configs[4])")
    return {
        "workload": name,
        "width": args.width, "height": args.height, "spp": args.spp, "bounces": args.bounces,
        "rng": "XorShift32 stream per (pixel, sample, frame), seed = sp_b200_Seed",
        "math_mode": "f64-rounded sin/cos/atan2/pow (bit-parity mode)" if args.math == 0 else "CUDA f32 libm",
        "env_filter": "nearest (reference image.h:3-18)",
        "l2": "env map 134 MB + framebuffer 133 MB touched per step exceed the 126 MB L2; the "
              "1.3 MB BVH is re-read by every ray inside a step by design; frame index advances per step",
        "partition": f"strips of whole {args.strip_rows}-pixel rows, re-cut every warm-up frame from the measured per-row cost (sp_b200_PartitionRows: min-max optimal contiguous cut)",
        "scheduler": "wavefront (trace / shade-miss / shade-hit / accumulate kernels over device queues)"
                     if args.render_mode == 0 else "per-pixel kernel",
    }


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(path))
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
        for v in d.values():
            if isinstance(v, dict) and "hbm_gbs" in v:
                return float(v["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.file = None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "25"],
                stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.file.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.file.name)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                   "power_w": float(np.median(pw)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------------
# CPU arms (oracle/ is test infrastructure: only the cpu_baseline and --impl reference legs use it)

def cpu_sample_tiles(width, height, stride):
    """Every `stride`-th 64x64 tile of the frame in ComputeTiles order (evenly spread)."""
    tw = th = 64
    tx, ty = (width + tw - 1) // tw, (height + th - 1) // th
    while stride > 1 and math.gcd(stride, tx) != 1:   # never sample whole tile columns
        stride += 1
    tiles = []
    for i in range(0, tx * ty, stride):
        x, y = (i % tx) * tw, (i // tx) * th
        tiles.append((x, y, min(x + tw, width), min(y + th, height)))
    return np.asarray(tiles, dtype=np.uint32), tx * ty


def cpu_time_sample(args, wl, seconds, kind):
    """Render an evenly spread subset of the frame's tiles with the CPU checker's native tile
    scheduler (all host threads), sized by a calibration pass to take about `seconds`."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ora
    cores = os.cpu_count() or 1
    use_ref = kind == "reference" and ora.have_ref()
    lib = ora.load_ref() if use_ref else ora.load_port()
    bounces = 3 if use_ref else args.bounces
    s = lib.scene().load_workload(wl)
    image = np.zeros((wl.height, wl.width, 4), np.float32)
    cal_tiles, total = cpu_sample_tiles(wl.width, wl.height, 97)
    _, m, secs = s.render_tile_list(cal_tiles, spp=args.spp, bounces=bounces, threads=cores, image=image)
    per_tile = secs / max(1, len(cal_tiles))
    want = int(max(cores, min(total, seconds / max(per_tile, 1e-6))))
    stride = max(1, total // want)
    tiles, _ = cpu_sample_tiles(wl.width, wl.height, stride)
    _, m, secs = s.render_tile_list(tiles, spp=args.spp, bounces=bounces, threads=cores, image=image)
    s.close()
    rays = int(m[2])
    return {"value": rays / secs / 1e6, "unit": "Mrays/s", "cores": cores,
            "kind": "reference" if use_ref else "port",
            "sample": f"{len(tiles)} of {total} 64x64 tiles (every {stride}th, whole frame spread) at "
                      f"{args.spp} spp, {bounces} bounces, native tile queue, {cores} threads, "
                      f"{rays} rays in {secs:.2f} s",
            "seconds": secs, "rays": rays, "tiles": tiles}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # workloads.py alone, by path: importing the package would map libspb200.so into a process that
    # must not contain any of the product (the driver records the .so files each arm loaded)
    import importlib.util
    spec = importlib.util.spec_from_file_location("spb_workloads", os.path.join(ROOT, "vk_cinematic_b200", "workloads.py"))
    W = importlib.util.module_from_spec(spec)
    sys.modules["spb_workloads"] = W
    spec.loader.exec_module(W)
    wl = W.config3(args.width, args.height, spp=args.spp, bounces=args.bounces)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ora
    cores = os.cpu_count() or 1
    lib = ora.load_port()
    s = lib.scene().load_workload(wl)
    image = np.zeros((wl.height, wl.width, 4), np.float32)
    cal_tiles, total = cpu_sample_tiles(wl.width, wl.height, 97)
    _, m, secs = s.render_tile_list(cal_tiles, spp=args.spp, bounces=args.bounces, threads=cores, image=image)
    per_tile = secs / max(1, len(cal_tiles))
    budget = 150.0 / max(1, args.steps + args.warmup)        # whole run within a few minutes
    want = int(max(cores, min(total, min(budget, 20.0) / max(per_tile, 1e-6))))
    stride = max(1, total // want)
    tiles, _ = cpu_sample_tiles(wl.width, wl.height, stride)
    for _ in range(args.warmup):
        s.render_tile_list(tiles, spp=args.spp, bounces=args.bounces, threads=cores, image=image)
    rays, secs = 0, 0.0
    for _ in range(args.steps):
        _, m, t = s.render_tile_list(tiles, spp=args.spp, bounces=args.bounces, threads=cores, image=image)
        rays += int(m[2])
        secs += t
    s.close()
    value = rays / secs / 1e6
    sample = (f"each step = {len(tiles)} of {total} 64x64 tiles (every {stride}th) at {args.spp} spp, "
              f"{args.bounces} bounces, native tile queue (main.cpp:731-759), {cores} threads")
    out = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    # the reference's own sources on the same sample (3 bounces: literal at simd_path_tracer.cpp:195)
    if ora.have_ref():
        r = ora.load_ref().scene().load_workload(wl)
        _, m, t = r.render_tile_list(tiles, spp=args.spp, bounces=3, threads=cores, image=image)
        r.close()
        out["reference_verbatim_3_bounces"] = {"value": int(m[2]) / t / 1e6, "unit": "Mrays/s",
                                               "kind": "reference", "cores": cores}
    print(json.dumps(out), file=RESULT_OUT, flush=True)


# ------------------------------------------------------------------------------------------------

def claim_stdout():
    """stdout must carry exactly one JSON line, but libraries print there too (NCCL's version
    banner ignores NCCL_DEBUG_FILE).  Keep the real stdout for the result and point fd 1 at
    stderr for everything else."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def load_json(*parts):
    try:
        return json.load(open(os.path.join(ROOT, *parts)))
    except Exception:
        return None


def on_chip_peaks(bvh_bytes):
    """L2 / L1 read peaks of this pool's B200 measured by tools/l2_peak.cu (committed sweep,
    profiles/r2/l2_peak_sweep.json): the L2 figure for the smallest swept working set that holds the
    scene's BVH (what the traversal re-reads), the plateau over the large sets, the L1 figure."""
    lp = load_json("profiles", "r2", "l2_peak_sweep.json")
    if not lp:
        return None
    sweep = {int(k.replace("MiB", "")): float(v) for k, v in lp["l2_cg_16B_GBs"].items()}
    sizes = sorted(sweep)
    fit = next((m for m in sizes if m * (1 << 20) >= bvh_bytes), sizes[-1])
    plateau = min(sweep[m] for m in sizes if m >= 16)
    return {"l2_read_peak": sweep[fit], "l2_working_set_mib": fit, "l2_sweep_GBs": sweep, "l2_plateau_16_to_96_mib": plateau,
            "l1_read_peak": float(lp["l1_64KiB_per_cta_GBs"]), "hbm_read_peak": float(lp["hbm_1GiB_GBs"]),
            "source": "profiles/r2/l2_peak_sweep.json (tools/l2_peak.cu: ld.global.cg read sweeps, 2-96 MiB working sets, measured)"}


def roofline_block(sp, args, r, bounds, rank, image_ptr, frame, timed_trace, world, kernel_ms, workload, scene_bytes):
    """Roofline of the dominant kernel (k_trace: BVH traversal + triangle tests; every launch of a step is
    timed with CUDA events inside the library).  A stats launch of one frame counts I and L."""
    trace_ms = float(np.mean([t[0] for t in timed_trace])) if timed_trace else 0.0
    trace_launches = int(np.mean([t[1] for t in timed_trace])) if timed_trace else 0
    traced_rays = float(np.mean([t[2] for t in timed_trace])) if timed_trace else 0.0
    sp.lib.sp_b200_EnableStats(1)
    b, e = bounds[rank]
    st = None
    if e > b:
        r.render_rows(b, e, frame=frame, host=False, device_ptr=image_ptr)
        st = sp.last_stats()
    sp.lib.sp_b200_EnableStats(0)
    if st is None or trace_ms <= 0:
        return None
    srays = max(1, int(st.tracedRays))
    I, L = st.nodeVisits / srays, st.triangleTests / srays
    # traversal: per traced ray 128 I + 48 L + 64; sky-kernel samples never enter the
    # traversal kernel and are not counted
    per_ray = 128.0 * I + 48.0 * L + 64.0
    bytes_per_step = per_ray * traced_rays
    achieved = bytes_per_step / (trace_ms * 1e-3) / 1e9
    hbm_peak, which = measured_peak()
    chip = on_chip_peaks(scene_bytes)
    # DRAM traffic per launch from the committed ncu capture of this workload (1 GPU, whole frame);
    # with several ranks a launch covers a strip: scaled by traced rays (per-ray DRAM bytes are a
    # property of the kernel and the queues, not of the strip)
    traffic, traffic_how = None, None
    tr = load_json("profiles", "r2", f"trace_traffic_{workload}.json")
    if tr and tr.get("spp") == args.spp and tr.get("width") == args.width and tr.get("launches") and tr.get("traced_rays_per_step"):
        per_traced_ray = float(tr["dram_bytes_per_step"]) / float(tr["traced_rays_per_step"])
        traffic = per_traced_ray * traced_rays / max(1, trace_launches)
        traffic_how = ("ncu dram__bytes_read+write of every k_trace launch of one frame (profiles/r2/trace_traffic_%s.json), "
                       "per launch%s" % (workload, "" if world == 1 else "; scaled to this rank's traced rays"))
    out = {"bound": "l2", "achieved": achieved, "unit": "GB/s",
           "peak": chip["l2_read_peak"] if chip else None,
           "frac": achieved / chip["l2_read_peak"] if chip else None,
           "peak_source": ("measured L2 read peak (tools/l2_peak.cu) for a %d MiB working set, the smallest swept set that holds "
                           "the scene's %.1f MB of nodes and triangles" % (chip["l2_working_set_mib"], scene_bytes / 1e6)) if chip else "unavailable",
           "traffic": traffic, "traffic_source": traffic_how,
           "kernel": "k_trace (BVH traversal + triangle tests)",
           "launches_per_step": trace_launches, "kernel_ms_per_step": trace_ms,
           "step_kernels_ms": float(np.mean(kernel_ms)) if kernel_ms else 0.0,
           "avg_launch_ms": trace_ms / max(1, trace_launches),
           "algorithmic_bytes_per_launch": bytes_per_step / max(1, trace_launches),
           "traced_rays_per_step": traced_rays, "traversal_grays_per_s": traced_rays / (trace_ms * 1e-3) / 1e9,
           "node_visits_per_ray": I, "triangle_tests_per_ray": L, "algorithmic_bytes_per_ray": per_ray,
           "hbm": {"peak": hbm_peak, "frac": achieved / hbm_peak, "peak_source": which,
                   "note": "context only: the BVH is served from L1/L2; DRAM sees the ray and hit-record queues (traffic)"},
           "on_chip": chip,
           "note": "achieved = algorithmic bytes (128 I + 48 L + 64 per traced ray, I and L counted by the kernel) of the step's "
                   "k_trace launches / their summed CUDA-event time.  The kernel is bound by instruction issue, not by a memory "
                   "level (profiles/r2/README.md: lanes per instruction x issue-slot utilisation); the L2 fraction is what "
                   "SURVEY.md §8d asks to be reported"}
    if chip:
        out["frac_of_l2_plateau"] = achieved / chip["l2_plateau_16_to_96_mib"]
        out["frac_of_l1_read_peak"] = achieved / chip["l1_read_peak"]
    return out


def parity_block(sp, W, args, render_full_frame, rank):
    """Whole-frame parity of what this run's library renders, against the committed fingerprints of the CPU
    checkers (tests/golden/g5, g6: made by tools/make_full_frame_golden.py from the unmodified reference
    and the restatement; no checker is called here).  Every rank renders its strip; rank 0 compares."""
    from vk_cinematic_b200.fixtures import lattice, relative_error_report, tile_crcs
    gold = np.load(os.path.join(ROOT, "tests", "golden", "g5_c3_full_frame.npz"))
    out = {}

    def compare(image, key, ties):
        crc = tile_crcs(image)
        want = gold["tile_crc_" + key]
        bad = np.nonzero(crc != want)[0]
        tx = (image.shape[1] + 63) // 64
        tie_tiles = set(int(y // 64) * tx + int(x // 64) for x, y in gold[ties])
        return {"tiles": int(len(want)), "tiles_differing": int(len(bad)),
                "tiles_differing_without_a_tie_pixel": int(sum(1 for t in bad if int(t) not in tie_tiles)),
                "tie_pixels_in_frame": int(len(gold[ties])), "pixels": int(image.shape[0] * image.shape[1])}

    img5, m5 = render_full_frame(5, 0)
    if rank == 0:
        out["c3_5_bounces_vs_port_dm"] = compare(img5, "port_dm_5b", "ties_5b")
        out["c3_5_bounces_vs_port_dm"]["rays"] = [int(m5), int(gold["metrics_port_dm_5b"][1])]
    img3, m3 = render_full_frame(3, 0)
    if rank == 0:
        out["c3_3_bounces_vs_reference_dm"] = compare(img3, "ref_dm_3b", "ties_3b")
        out["c3_3_bounces_vs_reference_dm"]["rays"] = [int(m3), int(gold["metrics_ref_dm_3b"][1])]
        rep = relative_error_report(lattice(img3), gold["lattice_ref_3b"])
        rep["what"] = ("deterministic-math render vs the unmodified reference with glibc libm, 3 bounces, on the fixture's lattice "
                       "of every 16th pixel; stated tolerance: <= 2 % of pixels off by > 1e-3 relative, RMSE <= 2 % of mean radiance")
        out["c3_3_bounces_vs_plain_reference"] = rep
    return out


def c2_parity(sp, W, local):
    """BASELINE configs[1] (monkey, 1920x1080, primary rays): closest-hit triangle ids, distances and object
    ids against the fingerprint of the unmodified reference's (tests/golden/g6_c2_full_frame.npz)."""
    from vk_cinematic_b200.fixtures import tile_crcs
    gold = np.load(os.path.join(ROOT, "tests", "golden", "g6_c2_full_frame.npz"))
    wl = W.config2()
    r = sp.Renderer(local).load_workload(wl)
    t0 = time.perf_counter()
    g = r.primary_hits(sample=0, frame=0)
    secs = time.perf_counter() - t0
    st = sp.last_stats()
    r.close()
    tri_off = int((tile_crcs(g["tri"]) != gold["tile_crc_tri"]).sum())
    return {"rays": int(g["tri"].size), "hit_pixels": [int((g["tri"] >= 0).sum()), int(gold["hit_pixels"])],
            "tiles": int(len(gold["tile_crc_tri"])), "tiles_with_differing_triangle_ids": tri_off,
            "tiles_with_differing_t_bits": int((tile_crcs(g["t"]) != gold["tile_crc_t"]).sum()),
            "tiles_with_differing_object_ids": int((tile_crcs(g["obj"]) != gold["tile_crc_obj"]).sum()),
            "triangle_id_mismatch_rate_bound": tri_off * 4096 / g["tri"].size,
            "stated_bound": 1e-4, "kernel_ms": st.kernelMs, "primary_mrays_per_s": g["tri"].size / max(st.kernelMs, 1e-9) / 1e3,
            "call_seconds": secs}


def main():
    args = parse_args()
    global RESULT_OUT
    RESULT_OUT = claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from vk_cinematic_b200 import sp, strips, workloads as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA devices; libspb200 has no CPU path"
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: whatever NCCL_DEBUG asks for goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if args.nccl_channels:
            os.environ["NCCL_MAX_NCHANNELS"] = str(args.nccl_channels)
            os.environ["NCCL_MIN_NCHANNELS"] = str(min(args.nccl_channels, 2))
        dist.init_process_group("nccl", device_id=dev)
        # "the frame is complete in host memory on every rank" is a statement about the hosts: the end-to-end steps
        # mark it with a CPU barrier (gloo).  An NCCL barrier is a kernel on the render stream -- it would run behind
        # the NEXT frame, already enqueued there (two frames in flight), and make every step wait for it.
        host_group = dist.new_group(backend="gloo")
    assert sp.lib.sp_b200_Init(local) == 0
    stream = torch.cuda.current_stream()
    sp.lib.sp_b200_SetStream(stream.cuda_stream)

    def make_workload(name, width, height, spp, bounces):
        if name == "c3":
            return W.config3(width, height, spp=spp, bounces=bounces)
        return W.config5(width, height, spp=spp, bounces=bounces, unique_spheres=name == "c5u")

    wl = make_workload(args.workload, args.width, args.height, args.spp, args.bounces)
    H, Wd = wl.height, wl.width
    # pinned host buffers: environment map (input) and the image plane (output).  With several ranks the
    # image plane is ONE pinned buffer shared by all processes (a file in /dev/shm mapped by every rank
    # and registered with cudaHostRegister): each GPU copies its own rows there over its own PCIe link.
    env_key = W.IMAGE_ENV
    env_pinned = torch.from_numpy(wl.textures[env_key]).pin_memory()
    wl.textures[env_key] = env_pinned.numpy()
    shared_path = None
    if world > 1:
        name = [f"/dev/shm/spb200_frame_{os.getpid()}"] if rank == 0 else [None]
        dist.broadcast_object_list(name, src=0)
        shared_path = name[0]
        if rank == 0:
            with open(shared_path, "wb") as f:
                f.truncate(2 * H * Wd * 16)
        dist.barrier()
        # (two images: with two frames in flight -- sp_b200_RenderRowsBegin / End -- consecutive frames must not
        # share their destination)
        host_np = np.memmap(shared_path, dtype=np.float32, mode="r+", shape=(2, H, Wd, 4))
        host_both = torch.from_numpy(host_np)
        rc = torch.cuda.cudart().cudaHostRegister(host_both.data_ptr(), host_both.numel() * 4, 0)
        assert int(rc) == 0, f"cudaHostRegister failed: {rc}"
        host_images = [host_both[0], host_both[1]]
    else:
        host_images = [torch.zeros((H, Wd, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    host_image = host_images[0]
    if args.device_builder:
        sp.lib.sp_b200_SetMeshBuilder(sp.BUILDER_DEVICE_LBVH)
    r = sp.Renderer(local).load_workload(wl, pixels=host_image.numpy())
    sp.lib.sp_b200_SetMeshBuilder(sp.BUILDER_HOST_SAH)

    def set_render_params(spp, bounces):
        sp.set_params(samplesPerPixel=spp, bounceCount=bounces, cullByDistance=1,
                      mathMode=args.math, envFilter=0, radianceClamp=10.0, tileWidth=64, tileHeight=args.strip_rows,
                      renderMode=args.render_mode, samplesPerPass=args.samples_per_pass)
    set_render_params(args.spp, args.bounces)
    if args.paths_per_pass:
        sp.lib.sp_b200_SetPathsPerPass(args.paths_per_pass)
    if args.no_ray_sorting:
        sp.lib.sp_b200_SetRaySorting(0)
    if args.no_candidates:
        sp.lib.sp_b200_SetPrimaryCandidates(0)
    if args.sort_bounces:
        sp.lib.sp_b200_SetRaySorting(args.sort_bounces)
    if args.refill:
        sp.lib.sp_b200_SetRefillThresholds(*[int(x) for x in args.refill.split(",")])
    if args.evict:
        sp.lib.sp_b200_SetStragglerEviction(*[int(x) for x in args.evict.split(",")])
    if args.fuse_miss >= 0:
        sp.lib.sp_b200_SetMissFusion(args.fuse_miss)
    if args.no_copy_overlap:
        sp.lib.sp_b200_SetCopyOverlap(0)
    TH = args.strip_rows   # strip boundaries and cost accounting: rows of TH pixels
    # Two device images, used alternately: the strips of frame k are on their way to rank 0 (NCCL, on
    # its own stream) while frame k + 1 renders into the other one.
    pipelined = not args.no_pipeline
    images = [torch.zeros((H, Wd, 4), dtype=torch.float32, device=dev) for _ in range(2 if (world > 1 or pipelined) else 1)]
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    gathered = [torch.cuda.Event() for _ in images]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    gather_ev = []
    trace_log = []
    step_counter = [0]

    def gather_async(image, bounds, slot, after=None):
        """The one exchange step: strips to rank 0, one batched send/recv group on the
        communication stream; the render stream goes on with the next frame."""
        if after is not None:
            comm.wait_event(after)               # (the frame's last kernel; later frames may already be enqueued)
        else:
            comm.wait_stream(stream)
        with torch.cuda.stream(comm):
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(comm)
            strips.gather_strips(image, bounds, dist)
            g1.record(comm)
            gathered[slot].record(comm)
        gather_ev.append((g0, g1))

    # The gather of frame k is ISSUED by a helper thread while the main thread is already inside the
    # library with frame k + 1 (ctypes releases the GIL there): building the send / recv group costs
    # 0.1-0.2 ms of host time per frame, which at 8 ms per frame is 2 % of a GPU left idle.
    from concurrent.futures import ThreadPoolExecutor
    issuer = ThreadPoolExecutor(max_workers=1) if world > 1 else None
    pending = [None for _ in images]

    def issue_gather(image, bounds, slot, after=None):
        torch.cuda.set_device(local)
        gather_async(image, bounds, slot, after)

    def drain_gathers():
        for i, f in enumerate(pending):
            if f is not None:
                f.result()
                pending[i] = None

    def render_step(frame, bounds, want_cost=False, wait_gather=False):
        slot = step_counter[0] % len(images)
        step_counter[0] += 1
        image = images[slot]
        if pending[slot] is not None:            # this buffer's previous gather has been issued ...
            pending[slot].result()
            pending[slot] = None
        stream.wait_event(gathered[slot])        # ... and its strips have left before we overwrite them
        b, e = bounds[rank]
        m, cost = np.zeros(12, np.uint64), None
        if e > b:
            m, cost = r.render_rows(b, e, frame=frame, host=False, device_ptr=image.data_ptr(),
                                    want_cost=want_cost)
        st = sp.last_stats()
        if world > 1:
            if wait_gather:
                drain_gathers()
                gather_async(image, bounds, slot)
                stream.wait_stream(comm)
            else:
                pending[slot] = issuer.submit(issue_gather, image, bounds, slot)
        trace_log.append((st.traceMs, st.traceLaunches, int(st.tracedRays)) if e > b else (0.0, 0, 0))
        return m, cost, (st.kernelMs if e > b else 0.0), image

    # The same step in two halves (sp_b200_RenderRowsBegin / End, two frames in flight): begin frame k + 1, then
    # end frame k -- the host-side epilogue of k (event times, metrics) and the prologue of k + 1 run under
    # k + 1's kernels, and the device does not idle across the frame boundary.
    def step_begin(frame, bounds):
        slot = step_counter[0] % len(images)
        step_counter[0] += 1
        image = images[slot]
        if pending[slot] is not None:
            pending[slot].result()
            pending[slot] = None
        stream.wait_event(gathered[slot])
        b, e = bounds[rank]
        h = r.render_rows_begin(b, e, frame=frame, host=False, device_ptr=image.data_ptr()) if e > b else None
        if world > 1:
            done = torch.cuda.Event()
            done.record(stream)
            pending[slot] = issuer.submit(issue_gather, image, bounds, slot, done)
        return h

    def step_end(h):
        if h is None:
            trace_log.append((0.0, 0, 0))
            return np.zeros(12, np.uint64), 0.0
        m, _ = r.render_rows_end(h)
        st = sp.last_stats()
        trace_log.append((st.traceMs, st.traceLaunches, int(st.tracedRays)))
        return m, st.kernelMs

    # ---- warm-up.  With several ranks every warm-up frame but the first is also one iteration of the
    # strip rebalancing: the cut of the next frame is made from this frame's per-row cost (nanoseconds,
    # sp_b200_RenderRows) scaled to what the rank's step took on the host clock -- kernels plus the
    # rank's own launch and gather bookkeeping, which is what bounds its frame rate (NOT the time to
    # the end of the gather, which waits for the slowest rank and would hide the imbalance).  The
    # first frame allocates the working set and is not a measurement.
    bounds = strips.partition_rows(H, TH, world)
    frame = 0
    warmups = max(args.warmup, 3)
    history = []
    for w in range(warmups):
        barrier()
        t_step = time.perf_counter()
        m, cost, kms, _ = render_step(frame, bounds, want_cost=True, wait_gather=False)
        step_s = time.perf_counter() - t_step
        if world > 1:
            drain_gathers()
        torch.cuda.synchronize()
        frame += 1
        if world > 1:
            row_cost, _secs = strips.gather_row_costs(
                cost if cost is not None else np.zeros(0), step_s, bounds, H, TH, dist, dev)
            history.append({"strips": [int(b[1]) for b in bounds], "step_ms": [round(float(x) * 1e3, 3) for x in _secs],
                            "kernel_ms_rank0": round(kms, 3)})
            if w >= 1:
                bounds = strips.partition_rows(H, TH, world, row_cost)

    # ---- timed region: exactly K steps
    launches0 = sp.lib.sp_b200_KernelLaunchCount()
    gather_ev.clear()
    trace_log.clear()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    rays = 0
    kernel_ms = []
    barrier()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    if pipelined:
        handle = None
        for k in range(args.steps + 1):
            nxt = step_begin(frame, bounds) if k < args.steps else None
            frame += 1 if k < args.steps else 0
            if k > 0:
                m, kms = step_end(handle)
                rays += int(m[sp.sp_Metric_RaysTraced])
                kernel_ms.append(kms)
            handle = nxt
    else:
        for k in range(args.steps):
            m, _, kms, _ = render_step(frame, bounds)
            rays += int(m[sp.sp_Metric_RaysTraced])
            kernel_ms.append(kms)
            frame += 1
    if world > 1:
        drain_gathers()
        stream.wait_stream(comm)                 # the last strips have arrived
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = sp.lib.sp_b200_KernelLaunchCount() - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    timed_trace = list(trace_log[:args.steps])
    traced = float(np.sum([t[2] for t in timed_trace]))
    stat = torch.tensor([elapsed_ms, float(rays), float(launches), float(np.sum(kernel_ms)), traced],
                        dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        gather_ms = float(np.mean([a.elapsed_time(b) for a, b in gather_ev[:args.steps]])) if gather_ev else 0.0
        mine = torch.tensor([float(np.mean(kernel_ms)), gather_ms, float(rays) / args.steps], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"kernel_ms": float(t[0]), "gather_ms": float(t[1]), "rays_per_step": float(t[2])} for t in allr]
        mx = stat.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stat.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        elapsed_ms, rays, launches, traced = float(mx[0]), float(sm[1]), float(sm[2]), float(sm[4])
    value = rays / (elapsed_ms * 1e-3) / 1e6

    if args.quick:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": elapsed_ms / args.steps, "kernel_ms": float(np.mean(kernel_ms)),
                              "trace_ms": float(np.mean([t[0] for t in timed_trace])) if timed_trace else 0.0,
                              "rays_per_step": rays / args.steps, "traced_rays_per_step": traced / args.steps,
                              "launches": int(launches), "strips": [list(map(int, b)) for b in bounds],
                              "ranks": per_rank, "rebalance_history": history}), file=RESULT_OUT, flush=True)
        r.close()
        if world > 1:
            issuer.shutdown()
            dist.barrier()
            dist.destroy_process_group()
            if rank == 0:
                os.unlink(shared_path)
        return

    # ---- e2e: host buffers, copies inside the timed region.  Every step: the scene is flattened and
    # uploaded again (sp_BuildSceneBroadphase), the environment map and the material table go to the
    # device again (texture cache flushed), the strip is rendered, and its rows are copied to the pinned
    # host image (band by band on the copy stream while later bands render).  With several ranks every
    # rank uploads one slice of the environment map and an NCCL all-gather over NVLink assembles the
    # copies (sp_b200_SetDeviceTexture), and every rank writes its rows into the shared host image.
    env_flat = env_pinned.view(-1)
    # (several ranks: two device copies of the map, filled alternately -- the slice upload and the
    # all-gather of step k + 1 run on a side stream while step k renders from the other copy)
    env_dev = [torch.empty_like(env_flat, device=dev) for _ in range(2)] if world > 1 else None
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    env_ready = [torch.cuda.Event() for _ in range(2)] if world > 1 else None
    env_issued = [-1, -1]
    eh, ew = wl.textures[env_key].shape[0], wl.textures[env_key].shape[1]

    def upload_env(step):
        """This rank's slice of the environment map to the device and the all-gather of the other ranks'
        slices over NVLink, for e2e step `step`, on the side stream."""
        buf = step % 2
        if env_issued[buf] == step:
            return
        n = env_flat.numel() // world
        side.wait_stream(stream)                      # the render that last read this copy is enqueued before us
        with torch.cuda.stream(side):
            lo, hi = rank * n, (rank + 1) * n
            env_dev[buf][lo:hi].copy_(env_flat[lo:hi], non_blocking=True)
            dist.all_gather_into_tensor(env_dev[buf][:n * world], env_dev[buf][lo:hi])
            if n * world < env_flat.numel():
                env_dev[buf][n * world:].copy_(env_flat[n * world:], non_blocking=True)
            env_ready[buf].record(side)
        env_issued[buf] = step

    e2e_counter = [0]
    e2e_phases = np.zeros(4)                          # host seconds: inputs, scene build, render call, barrier
    e2e_device = np.zeros(2)                          # device ms inside the render call: kernels, whole call

    def e2e_begin(frame):
        """First half of an end-to-end step: inputs to the device, scene re-built, the strip enqueued with its rows
        going to host image k % 2.  With --no-pipeline the frame is also waited for here."""
        b, e = bounds[rank]
        k = e2e_counter[0]
        e2e_counter[0] += 1
        t_a = time.perf_counter()
        if world > 1:
            upload_env(k)                             # (already in flight since the previous step, except for the first)
            sp.lib.sp_b200_SetDeviceTexture(env_pinned.data_ptr(), env_dev[k % 2].data_ptr(), ew, eh, env_ready[k % 2].cuda_event)
        else:
            sp.lib.sp_b200_FlushTextureCache()       # env map + materials re-uploaded
        t_b = time.perf_counter()
        r.build()                                     # scene flattened and re-uploaded
        t_c = time.perf_counter()
        if world > 1:
            upload_env(k + 1)                         # next step's copy goes up under this step's kernels
        h = None
        if e > b:
            h = r.render_rows_begin(b, e, frame=frame, host_ptr=host_images[k % 2].data_ptr())   # rows -> pinned host image
        t_d = time.perf_counter()
        e2e_phases[:3] += (t_b - t_a, t_c - t_b, t_d - t_c)
        return h

    def e2e_end(h):
        """Second half: the frame is complete in host memory on every rank."""
        t_a = time.perf_counter()
        m = np.zeros(12, np.uint64)
        if h is not None:
            m, _ = r.render_rows_end(h)
            st_ = sp.last_stats()
            e2e_device[:] += (st_.kernelMs, st_.totalMs)
        t_b = time.perf_counter()
        if world > 1:
            dist.barrier(group=host_group)            # the frame is complete in host memory
        t_c = time.perf_counter()
        e2e_phases[2:] += (t_b - t_a, t_c - t_b)
        return m

    def e2e_run(steps, frame):
        """`steps` end-to-end frames, frame k + 1 begun before frame k is ended (two in flight) unless --no-pipeline."""
        total = 0
        handle = None
        for k in range(steps + (1 if pipelined else 0)):
            nxt = e2e_begin(frame + k) if k < steps else None
            if not pipelined:
                total += int(e2e_end(nxt)[sp.sp_Metric_RaysTraced])
                continue
            if k > 0:
                total += int(e2e_end(handle)[sp.sp_Metric_RaysTraced])
            handle = nxt
        return total
    e2e_run(2, frame)
    frame += 2
    barrier()
    e2e_phases[:] = 0
    e2e_device[:] = 0
    t0 = time.perf_counter()
    e2e_rays = e2e_run(args.steps, frame)
    frame += args.steps
    barrier()
    e2e_secs = time.perf_counter() - t0
    if world > 1:
        torch.cuda.synchronize()
        sp.lib.sp_b200_SetDeviceTexture(env_pinned.data_ptr(), None, 0, 0, None)
        sp.lib.sp_b200_FlushTextureCache()
    est = torch.tensor([e2e_secs, float(e2e_rays)], dtype=torch.float64, device=dev)
    if world > 1:
        a = est.clone()
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b_ = est.clone()
        dist.all_reduce(b_, op=dist.ReduceOp.SUM)
        e2e_secs, e2e_rays = float(a[0]), float(b_[1])
    e2e_value = e2e_rays / e2e_secs / 1e6
    e2e_ranks = None
    if world > 1:
        mine = torch.tensor([e2e_device[0] / args.steps, e2e_device[1] / args.steps, e2e_phases[3] / args.steps * 1e3],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        e2e_ranks = [{"kernels_ms": round(float(t[0]), 3), "call_device_ms": round(float(t[1]), 3),
                      "host_barrier_wait_ms": round(float(t[2]), 3)} for t in allr]
    scene_bytes = int(sp.lib.sp_b200_SceneDeviceBytes(r.scene))
    env_bytes = int(env_pinned.numel() * 4)
    h2d = (scene_bytes + 2048) * world + env_bytes * (1 if world > 1 else 1)

    # ---- roofline of the dominant kernel
    roofline = roofline_block(sp, args, r, bounds, rank, images[0].data_ptr(), frame - 1, timed_trace, world, kernel_ms,
                              args.workload, scene_bytes) if rank == 0 else None
    if world > 1 and rank != 0:
        # (the stats frame above is rank 0's alone; nothing to do here)
        pass

    # ---- parity of this run's frames against the committed fingerprints (untimed)
    parity = None
    if not args.no_parity and args.workload == "c3" and (args.width, args.height, args.spp, args.math) == (3840, 2160, 64, 0):
        def render_full_frame(bounces, fr):
            set_render_params(args.spp, bounces)
            m, _, _, image = render_step(fr, bounds, wait_gather=True)
            torch.cuda.synchronize()
            total = torch.tensor([float(m[sp.sp_Metric_RaysTraced])], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(total)
            return (image.cpu().numpy() if rank == 0 else None), float(total[0])
        try:
            parity = parity_block(sp, W, args, render_full_frame, rank)
            set_render_params(args.spp, args.bounces)
            if rank == 0:
                parity["c2_primary_rays_vs_reference"] = c2_parity(sp, W, local)
        except Exception as ex:   # a missing fixture must not cost the bench line
            parity = {"error": repr(ex)}
        set_render_params(args.spp, args.bounces)

    # ---- secondary: BASELINE configs[4] (C5), time-bounded, same strips machinery
    secondary = {}
    if not args.no_secondary and args.workload == "c3":
        try:
            secondary["c5"] = secondary_c5(sp, W, strips, args, rank, world, local, dev, dist if world > 1 else None, torch)
        except Exception as ex:
            secondary["c5"] = {"error": repr(ex)}
    if rank == 0 and not args.no_secondary:
        try:
            secondary["reference_perf_tests"] = secondary_perf_tests(sp, W, local, world == 1 and not args.no_cpu_baseline)
        except Exception as ex:
            secondary["reference_perf_tests"] = {"error": repr(ex)}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_time_sample(args, wl, args.cpu_seconds, "port")
        cpu.pop("tiles", None)
        cpu.pop("rays", None)
        cpu.pop("seconds", None)

    if rank == 0:
        ms = elapsed_ms / args.steps
        out = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmups, "ms_per_step": ms,
            "s_per_frame": ms * 1e-3,
            "value_traced_only": traced / (elapsed_ms * 1e-3) / 1e6,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args),
            "rays_per_step": rays / args.steps, "traced_rays_per_step": traced / args.steps,
            "rays_accounting": "value counts the reference's rays for the same frame (sp_Metric_RaysTraced: one per "
                               "sp_RayIntersectScene call it makes, simd_path_tracer.cpp:242), which is what makes it comparable "
                               "with the CPU arm: both divide the same count by their own time.  Camera rays of pixels the "
                               "coverage pass proves empty are settled by the sky kernel (bit-identical result) and never "
                               "enter the traversal kernel: value_traced_only counts only rays that did; s_per_frame does not "
                               "depend on the accounting",
            "e2e": {"value": e2e_value, "unit": "Mrays/s",
                    "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": H * Wd * 16, "ms_per_step": e2e_secs / args.steps * 1e3,
                    "rank0_host_phases_ms": dict(zip(("inputs", "scene_build", "render_call", "barrier"),
                                                     [round(float(x) / args.steps * 1e3, 4) for x in e2e_phases])),
                    "rank0_render_call_device_ms": {"kernels": round(float(e2e_device[0]) / args.steps, 4), "call": round(float(e2e_device[1]) / args.steps, 4)},
                    "ranks": e2e_ranks,
                    "how": "per step: texture cache flushed, scene re-built and re-uploaded, environment map re-uploaded "
                           + ("(one slice per rank + NCCL all-gather over NVLink, double-buffered: step k + 1's copy goes up under step k's "
                              "kernels), rows copied by every rank into one shared pinned host image, barrier"
                              if world > 1 else "(copy stream, overlapped with coverage / candidates / first primary trace), "
                              "rows copied to the pinned host image band by band while later bands render")
                           + ("; two frames in flight (sp_b200_RenderRowsBegin / End): step k + 1's inputs, scene and launches are issued before "
                              "step k's frame is waited for, consecutive frames go to two pinned host images"
                              + ("; frame completion on all ranks marked by a host (gloo) barrier" if world > 1 else "") if pipelined else "")},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "strips": [list(map(int, b)) for b in bounds],
        }
        if per_rank is not None:
            out["ranks"] = per_rank
            out["rebalance_history"] = history
        if parity is not None:
            out["parity"] = parity
        if secondary:
            out["secondary"] = secondary
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), file=RESULT_OUT, flush=True)
    r.close()
    if world > 1:
        issuer.shutdown()
        dist.barrier()
        dist.destroy_process_group()
        if rank == 0:
            try:
                os.unlink(shared_path)
            except OSError:
                pass


def secondary_perf_tests(sp, W, local, with_cpu):
    """The reference's own performance tests (perf_tests/perf_tests.cpp:51-118 TestBvh -- the only
    performance number the repository states: < 3100 cycles per ray; :212-305 TestMeshMidphase) with the
    reference's seeded inputs restated draw for draw, as batched queries through the C ABI.  GPU figures:
    CUDA-event time of the query kernel, best of 5 (rays resident).  with_cpu (N = 1, the cpu_baseline leg):
    the same loops over the CPU checker, single-threaded as the reference runs them, and the two results
    compared (leaf sets per ray; t bit for bit)."""
    out = {}
    mn, mx, o, d = W.perf_bvh_inputs(sp.xorshift_bilateral_stream(0x1A34C249))
    r = sp.Renderer(local)
    boxes = W.boxes_as_triangles(mn, mx)
    t0 = time.perf_counter()
    m = r.add_mesh(boxes.vertices, boxes.indices, False)
    build_s = time.perf_counter() - t0
    best, got = 1e30, None
    for _ in range(6):
        got, ms = sp.mesh_leaves_batch(r.meshes[m], o, d)
        best = min(best, ms)
    out["TestBvh"] = {"boxes": int(len(mn)), "rays": int(len(o)), "gpu_kernel_us": best * 1e3, "gpu_ns_per_ray": best * 1e6 / len(o),
                      "gpu_tree_build_ms": build_s * 1e3, "leaves_per_ray": float(got[:, 0].mean()),
                      "reference_assert": "cyclesPerRay < 3100 on one CPU core (perf_tests.cpp:115)"}
    rays = 1024 * 8192
    mesh, o2, d2 = W.perf_mesh_inputs(sp.xorshift_bilateral_stream(0x1A34C249), rays)
    m2 = r.add_mesh(mesh.vertices, mesh.indices, False)
    best2, t_got = 1e30, None
    for _ in range(4):
        t_got, tri_got, ms = sp.mesh_intersect_batch(r.meshes[m2], o2, d2)
        best2 = min(best2, ms)
    out["TestMeshMidphase"] = {"triangles": int(mesh.triangle_count), "rays": rays, "gpu_kernel_ms": best2,
                               "gpu_ns_per_ray": best2 * 1e6 / rays, "gpu_mrays_per_s": rays / best2 / 1e3,
                               "hits": int((t_got >= 0).sum())}
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import ora
        lib = ora.load_ref() if ora.have_ref() else ora.load_port()
        want, cpu_build, cpu_q = lib.perf_bvh(mn, mx, o, d, 2048)
        # (half the reference's boxes have min > max, which makes ITS leaf sets depend on its tree; equality is
        # checked on the same draws with |radius|, containment on the input as it is: workloads.perf_bvh_inputs)
        mnp, mxp, op, dp = W.perf_bvh_inputs(sp.xorshift_bilateral_stream(0x1A34C249), proper=True)
        want_p, _, _ = lib.perf_bvh(mnp, mxp, op, dp, 2048)
        bp = W.boxes_as_triangles(mnp, mxp)
        got_p, _ = sp.mesh_leaves_batch(r.meshes[r.add_mesh(bp.vertices, bp.indices, False)], op, dp)
        out["TestBvh"].update({"cpu_ns_per_ray": cpu_q * 1e9 / len(o), "cpu_tree_build_ms": cpu_build * 1e3, "cpu_kind": lib.name,
                               "gpu_leaf_sets_contain_reference": bool(np.all(got[:, 0] >= want[:, 0])),
                               "leaf_sets_equal_on_proper_boxes": bool(np.array_equal(want_p, got_p))})
        n = 1 << 20   # bounded sample of the 8.4 M rays (about a second of one core)
        s = lib.scene()
        s.add_mesh(mesh.vertices, mesh.indices, False)
        t_want, _tri, secs = s.perf_mesh(0, o2[:n], d2[:n])
        s.close()
        out["TestMeshMidphase"].update({"cpu_ns_per_ray": secs * 1e9 / n, "cpu_sample_rays": n, "cpu_kind": lib.name,
                                        "t_bit_equal_on_sample": bool(np.array_equal(t_want.view(np.uint32), t_got[:n].view(np.uint32)))})
    r.close()
    return out


def secondary_c5(sp, W, strips, args, rank, world, local, dev, dist, torch):
    """BASELINE configs[4]: 64 bunnies + 118 level-6 icospheres (9.98 M instanced triangles), 3840x2160,
    16 spp, 5 bounces -- the configuration the reference cannot hold (32-object table, sp_scene.h:15).
    Same strips, same kernels; 2 warm-up + 3 timed frames per rank, device time, max over ranks."""
    wl = W.config5(3840, 2160, spp=16, bounces=5)
    H, Wd, TH = wl.height, wl.width, args.strip_rows
    t0 = time.perf_counter()
    r = sp.Renderer(local).load_workload(wl)
    build_s = time.perf_counter() - t0
    sp.set_params(samplesPerPixel=16, bounceCount=5, cullByDistance=1, mathMode=args.math, envFilter=0, radianceClamp=10.0,
                  tileWidth=64, tileHeight=TH, renderMode=0, samplesPerPass=0)
    image = torch.zeros((H, Wd, 4), dtype=torch.float32, device=dev)
    bounds = strips.partition_rows(H, TH, world)
    frame = 100

    def step(want_cost):
        b, e = bounds[rank]
        m, cost, kms, st = np.zeros(12, np.uint64), None, 0.0, None
        if e > b:
            m, cost = r.render_rows(b, e, frame=frame, host=False, device_ptr=image.data_ptr(), want_cost=want_cost)
            st = sp.last_stats()
            kms = st.kernelMs
        return m, cost, kms, st
    for _ in range(2):
        m, cost, kms, _ = step(True)
        frame += 1
        if world > 1:
            row_cost, _ = strips.gather_row_costs(cost if cost is not None else np.zeros(0), kms * 1e-3, bounds, H, TH, dist, dev)
            bounds = strips.partition_rows(H, TH, world, row_cost)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    rays, kms_all, trace_ms, traced, launches = 0, [], [], 0, []
    for _ in range(3):
        m, _, kms, st = step(False)
        rays += int(m[sp.sp_Metric_RaysTraced])
        kms_all.append(kms)
        if st is not None:
            trace_ms.append(st.traceMs)
            traced += int(st.tracedRays)
            launches.append(st.traceLaunches)
        frame += 1
    ev1.record()
    torch.cuda.synchronize()
    elapsed = ev0.elapsed_time(ev1)
    stat = torch.tensor([elapsed, float(rays), float(np.mean(kms_all))], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        allr = [torch.zeros_like(stat) for _ in range(world)]
        dist.all_gather(allr, stat)
        per_rank = [float(t[2]) for t in allr]
        elapsed = max(float(t[0]) for t in allr)
        rays = sum(float(t[1]) for t in allr)
    out = None
    if rank == 0:
        scene_bytes = int(sp.lib.sp_b200_SceneDeviceBytes(r.scene))
        sp.lib.sp_b200_EnableStats(1)
        b, e = bounds[rank]
        r.render_rows(b, e, frame=frame - 1, host=False, device_ptr=image.data_ptr())
        st = sp.last_stats()
        sp.lib.sp_b200_EnableStats(0)
        srays = max(1, int(st.tracedRays))
        I, L = st.nodeVisits / srays, st.triangleTests / srays
        per_ray = 128.0 * I + 48.0 * L + 64.0
        tms = float(np.mean(trace_ms)) if trace_ms else 0.0
        tr = traced / 3.0
        achieved = per_ray * tr / (tms * 1e-3) / 1e9 if tms > 0 else 0.0
        chip = on_chip_peaks(scene_bytes)
        hbm, which = measured_peak()
        traffic = None
        tj = load_json("profiles", "r2", "trace_traffic_c5.json")
        if tj and tj.get("traced_rays_per_step"):
            traffic = float(tj["dram_bytes_per_step"]) / float(tj["traced_rays_per_step"]) * tr / max(1, int(np.mean(launches)))
        out = {"workload": "C5: 64 bunnies + 118 level-6 icospheres (9.98 M instanced triangles, 182 objects), 3840x2160, 16 spp, "
                           "5 bounces (BASELINE.json configs[4])",
               "value": rays / (elapsed * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": elapsed / 3.0, "steps": 3, "warmup": 2,
               "n_gpus": world, "rank_kernel_ms": per_rank, "strips": [list(map(int, b)) for b in bounds],
               "scene_build_seconds": build_s, "scene_device_bytes": scene_bytes,
               "what": "device time of 3 frames (CUDA events, max over ranks); strips rendered, not gathered",
               "roofline": {"bound": "l2", "achieved": achieved, "unit": "GB/s",
                            "peak": chip["l2_read_peak"] if chip else None,
                            "frac": achieved / chip["l2_read_peak"] if chip else None,
                            "l2_working_set_mib": chip["l2_working_set_mib"] if chip else None,
                            "hbm": {"peak": hbm, "frac": achieved / hbm, "peak_source": which},
                            "traffic": traffic, "kernel": "k_trace", "kernel_ms_per_step": tms,
                            "launches_per_step": int(np.mean(launches)) if launches else 0,
                            "traced_rays_per_step": tr, "traversal_grays_per_s": tr / (tms * 1e-3) / 1e9 if tms > 0 else 0.0,
                            "node_visits_per_ray": I, "triangle_tests_per_ray": L, "object_entries_per_ray": st.objectTests / srays,
                            "algorithmic_bytes_per_ray": per_ray}}
    r.close()
    return out


if __name__ == "__main__":
    main()
