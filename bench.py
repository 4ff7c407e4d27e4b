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
    ap.add_argument("--strip-rows", type=int, default=16, help="granularity of strip boundaries in pixel rows (multiple of 4)")
    ap.add_argument("--sort-bounces", type=int, default=0, help="development: direction-sort the rays leaving this many bounces")
    ap.add_argument("--refill", default="", help="development: refill thresholds primary,sorted,other")
    ap.add_argument("--no-candidates", action="store_true", help="development: every primary ray walks the tree")
    ap.add_argument("--no-ray-sorting", action="store_true", help="development: bounce rays in hit-queue order")
    ap.add_argument("--paths-per-pass", type=int, default=0, help="development: paths in flight per pass (0 = library default)")
    ap.add_argument("--device-builder", action="store_true",
                    help="development: mesh trees from the device LBVH builder (sp_b200_SetMeshBuilder); default: host SAH")
    ap.add_argument("--quick", action="store_true", help="development: value only (no e2e, roofline, cpu baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        "partition": "strips of whole 16-pixel rows (quarter tiles), rebalanced from measured per-row cost",
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
    from vk_cinematic_b200 import workloads as W
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
        dist.init_process_group("nccl", device_id=dev)
    assert sp.lib.sp_b200_Init(local) == 0
    stream = torch.cuda.current_stream()
    sp.lib.sp_b200_SetStream(stream.cuda_stream)

    if args.workload == "c3":
        wl = W.config3(args.width, args.height, spp=args.spp, bounces=args.bounces)
    else:
        wl = W.config5(args.width, args.height, spp=args.spp, bounces=args.bounces,
                       unique_spheres=args.workload == "c5u")
    H, Wd = wl.height, wl.width
    # pinned host buffers: environment map (input) and the image plane (output)
    env_key = W.IMAGE_ENV
    env_pinned = torch.from_numpy(wl.textures[env_key]).pin_memory()
    wl.textures[env_key] = env_pinned.numpy()
    host_image = torch.zeros((H, Wd, 4), dtype=torch.float32).pin_memory()
    if args.device_builder:
        sp.lib.sp_b200_SetMeshBuilder(sp.BUILDER_DEVICE_LBVH)
    r = sp.Renderer(local).load_workload(wl, pixels=host_image.numpy())
    sp.lib.sp_b200_SetMeshBuilder(sp.BUILDER_HOST_SAH)
    sp.set_params(samplesPerPixel=args.spp, bounceCount=args.bounces, cullByDistance=1,
                  mathMode=args.math, envFilter=0, radianceClamp=10.0, tileWidth=64, tileHeight=args.strip_rows,
                  renderMode=args.render_mode, samplesPerPass=args.samples_per_pass)
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
    TH = args.strip_rows   # strip boundaries and cost accounting: rows of 16 pixels (a quarter tile)
    image = torch.zeros((H, Wd, 4), dtype=torch.float32, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    gather_ev = []
    trace_log = []

    def render_step(frame, bounds, want_cost=False):
        b, e = bounds[rank]
        m, cost = np.zeros(12, np.uint64), None
        if e > b:
            m, cost = r.render_rows(b, e, frame=frame, host=False, device_ptr=image.data_ptr(),
                                    want_cost=want_cost)
        st = sp.last_stats()
        if world > 1:
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            strips.gather_strips(image, bounds, dist)
            g1.record()
            gather_ev.append((g0, g1))
        trace_log.append((st.traceMs, st.traceLaunches, int(st.tracedRays)) if e > b else (0.0, 0, 0))
        return m, cost, (st.kernelMs if e > b else 0.0)

    # ---- warm-up (also measures per-tile-row cost and re-cuts the strips)
    bounds = strips.partition_rows(H, TH, world)
    frame = 0
    # with several ranks every warm-up frame is also one iteration of the strip rebalancing (re-cut
    # from measured cost); 3 iterations from an even split did not converge at N = 8 (ranks at
    # 3.7-11.0 ms, profiles/r02r_bench_n8.json), so multi-rank runs take at least 8
    warmups = max(args.warmup, 3) if world == 1 else max(args.warmup, 8)
    for w in range(warmups):
        barrier()
        m, cost, kms = render_step(frame, bounds, want_cost=True)
        torch.cuda.synchronize()
        frame += 1
        if world > 1:
            # seconds = this rank's own render time (CUDA events around its kernels), NOT the step
            # time: the gather waits for the slowest rank and would hide the imbalance
            row_cost, _secs = strips.gather_row_costs(
                cost if cost is not None else np.zeros(0), kms * 1e-3, bounds, H, TH, dist, dev)
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
    for k in range(args.steps):
        m, _, kms = render_step(frame, bounds)
        rays += int(m[sp.sp_Metric_RaysTraced])
        kernel_ms.append(kms)
        frame += 1
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = sp.lib.sp_b200_KernelLaunchCount() - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    stat = torch.tensor([elapsed_ms, float(rays), float(launches), float(np.sum(kernel_ms))],
                        dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        gather_ms = float(np.mean([a.elapsed_time(b) for a, b in gather_ev])) if gather_ev else 0.0
        mine = torch.tensor([float(np.mean(kernel_ms)), gather_ms, float(rays) / args.steps], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"kernel_ms": float(t[0]), "gather_ms": float(t[1]), "rays_per_step": float(t[2])} for t in allr]
        mx = stat.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stat.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        elapsed_ms, rays, launches = float(mx[0]), float(sm[1]), float(sm[2])
    value = rays / (elapsed_ms * 1e-3) / 1e6

    # ---- e2e: host buffers, copies inside the timed region
    def e2e_step(frame):
        b, e = bounds[rank]
        sp.lib.sp_b200_FlushTextureCache()           # env map + materials re-uploaded
        r.build()                                     # scene flattened and re-uploaded
        m = np.zeros(12, np.uint64)
        if world == 1:
            _, m = r.render_frame(frame=frame)        # D2H of the finished frame into pinned host
        else:
            if e > b:
                m, _ = r.render_rows(b, e, frame=frame, host=False, device_ptr=image.data_ptr())
            strips.gather_strips(image, bounds, dist)
            if rank == 0:
                host_image.copy_(image, non_blocking=True)
                torch.cuda.synchronize()
        return m
    if args.quick:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": elapsed_ms / args.steps, "kernel_ms": float(np.mean(kernel_ms)),
                              "rays_per_step": rays / args.steps, "launches": int(launches), "strips": [list(map(int, b)) for b in bounds],
                              "ranks": per_rank}), file=RESULT_OUT, flush=True)
        r.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    scene_bytes = int(env_pinned.numel() * 4) + int(sp.lib.sp_b200_SceneDeviceBytes(r.scene)) + 2048
    for _ in range(2):
        e2e_step(frame)
        frame += 1
    barrier()
    t0 = time.perf_counter()
    e2e_rays = 0
    for k in range(args.steps):
        m = e2e_step(frame)
        e2e_rays += int(m[sp.sp_Metric_RaysTraced])
        frame += 1
    barrier()
    e2e_secs = time.perf_counter() - t0
    est = torch.tensor([e2e_secs, float(e2e_rays)], dtype=torch.float64, device=dev)
    if world > 1:
        a = est.clone()
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b_ = est.clone()
        dist.all_reduce(b_, op=dist.ReduceOp.SUM)
        e2e_secs, e2e_rays = float(a[0]), float(b_[1])
    e2e_value = e2e_rays / e2e_secs / 1e6

    # ---- roofline of the dominant kernel (k_trace, the traversal kernel; all its launches of a
    # step are timed with CUDA events inside the library): a stats launch of one frame counts I, L
    roofline = None
    cpu = None
    if rank == 0:
        timed_trace = trace_log[:args.steps]
        trace_ms = float(np.mean([t[0] for t in timed_trace])) if timed_trace else 0.0
        trace_launches = int(np.mean([t[1] for t in timed_trace])) if timed_trace else 0
        traced_rays = float(np.mean([t[2] for t in timed_trace])) if timed_trace else 0.0
        sp.lib.sp_b200_EnableStats(1)
        b, e = bounds[rank]
        m, _ = r.render_rows(b, e, frame=frame - 1, host=False, device_ptr=image.data_ptr())
        st = sp.last_stats()
        sp.lib.sp_b200_EnableStats(0)
        srays = max(1, int(st.tracedRays))
        I, L = st.nodeVisits / srays, st.triangleTests / srays
        # traversal: per traced ray 128 I + 48 L + 64; sky-kernel samples never
        # enter the traversal kernel and are not counted
        per_ray = 128.0 * I + 48.0 * L + 64.0
        bytes_per_step = per_ray * traced_rays
        achieved = bytes_per_step / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else 0.0
        peak, which = measured_peak()
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "trace_traffic.json")))
            if tr.get("workload") == args.workload and tr.get("spp") == args.spp and tr.get("width") == args.width \
                    and world == 1 and tr.get("launches"):
                traffic = float(tr["dram_bytes_per_step"]) / float(tr["launches"])
        except Exception:
            pass
        # on-chip read peaks of this pool's B200 (tools/l2_peak.cu, committed result): the levels the
        # resident BVH is actually served from
        onchip = {}
        try:
            lp = json.load(open(os.path.join(ROOT, "profiles", "r03b_l2_peak.json")))
            l2_peak = max(float(lp["l2_16MiB_GBs"]), float(lp["l2_48MiB_GBs"]))
            onchip = {"l2_read_peak": l2_peak, "frac_of_l2_read_peak": achieved / l2_peak,
                      "l1_read_peak": float(lp["l1_64KiB_per_cta_GBs"]),
                      "frac_of_l1_read_peak": achieved / float(lp["l1_64KiB_per_cta_GBs"]),
                      "source": "profiles/r03b_l2_peak.json (tools/l2_peak.cu: 16-byte read sweeps, measured)"}
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "on_chip": onchip,
                    "frac": achieved / peak, "traffic": traffic, "peak_source": which,
                    "kernel": "k_trace (BVH traversal + triangle tests)",
                    "launches_per_step": trace_launches,
                    "kernel_ms_per_step": trace_ms, "step_kernels_ms": float(np.mean(kernel_ms)) if kernel_ms else 0.0,
                    "avg_launch_ms": trace_ms / max(1, trace_launches),
                    "algorithmic_bytes_per_launch": bytes_per_step / max(1, trace_launches),
                    "traced_rays_per_step": traced_rays,
                    "node_visits_per_ray": I, "triangle_tests_per_ray": L,
                    "algorithmic_bytes_per_ray": per_ray,
                    "note": "achieved = algorithmic bytes of the step's k_trace launches / their summed CUDA-event time; "
                            "traffic = DRAM bytes per launch from the ncu capture under profiles/ (the 1.3 MB BVH is "
                            "L1/L2-resident, so DRAM traffic is the ray and hit-record queues); the fraction is of the "
                            "HBM copy peak although the bytes are served from L1/L2"}
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_time_sample(args, wl, args.cpu_seconds, "port")
            cpu.pop("tiles", None)
            cpu.pop("rays", None)
            cpu.pop("seconds", None)

    if rank == 0:
        out = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmups, "ms_per_step": elapsed_ms / args.steps,
            "s_per_frame": elapsed_ms / args.steps * 1e-3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args),
            "rays_per_step": rays / args.steps,
            "rays_accounting": "rays = the reference's sp_Metric_RaysTraced for the same frame (one per "
                               "sp_RayIntersectScene call it would make, simd_path_tracer.cpp:242); camera rays of "
                               "pixels the coverage pass proves empty are counted but never traced (sky kernel); "
                               "roofline.traced_rays_per_step is what went through the traversal kernel",
            "e2e": {"value": e2e_value, "unit": "Mrays/s",
                    "h2d_bytes_per_step": scene_bytes * world,
                    "d2h_bytes_per_step": H * Wd * 16, "ms_per_step": e2e_secs / args.steps * 1e3},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "strips": [list(map(int, b)) for b in bounds],
        }
        if per_rank is not None:
            out["ranks"] = per_rank
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), file=RESULT_OUT, flush=True)
    r.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
