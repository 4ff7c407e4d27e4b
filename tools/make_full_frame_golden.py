#!/usr/bin/env python
"""Full-size fixtures for the bench configurations ("C3 full-size parity covers about
1 232 of 8.29 M pixels"): tests/golden/g5_c3_full_frame.npz and g6_c2_full_frame.npz.

Run HERE, where /root/reference is mounted and oracle/_ref builds (several minutes of CPU):

    python tools/make_full_frame_golden.py [--threads N]

C3 = BASELINE configs[2], 3840x2160, 64 spp, frame 0, the bench's own workload (W.config3):

  tile_crc_port_dm_5b   2040 x u32   CRC-32 of every 64x64 tile (ComputeTiles order, tile.h:11-42;
                                     the tile's RGBA f32 rows, top to bottom) of the frame rendered
                                     by the restatement in deterministic-math mode at the
                                     workload's 5 bounces -- what the GPU must equal bit for bit;
  tile_crc_ref_dm_3b    2040 x u32   the same from the UNMODIFIED reference sources
                                     (oracle/_ref/libspref_dm.so) at the 3 bounces they are fixed
                                     at (simd_path_tracer.cpp:195): the full 4K frame pinned to the
                                     reference itself;
  lattice_ref_3b        135 x 240 x 3 f32  every 16th pixel (x % 16 == 8, y % 16 == 8) of the frame
                                     rendered by the unmodified reference with glibc's libm (the
                                     "plain" reference): per-pixel relative error and RMSE of the
                                     GPU's dm / f32 modes against what a user of the reference sees;
  tile_sum_ref_3b       2040 x 3 f64 per-tile sums of R, G, B of that plain frame (all pixels);
  metrics_*             paths, rays, hits, misses of each render;
  ties_5b / ties_3b     n x 2 u16  (x, y) of every pixel on whose paths some scene query ended with
                                     two candidates of bit-equal closest t (ora_tie_mask, port only):
                                     the one case where the winner depends on the order of visits,
                                     i.e. on tree topology, which the reference's algorithm does not
                                     fix.  A tile may differ from its CRC only if it holds such a pixel.

C2 = BASELINE configs[1], monkey 1920x1080 primary rays (sample 0, frame 0):

  tile_crc_tri / tile_crc_t / tile_crc_obj   510 x u32 per-tile CRCs of the closest-hit triangle
                                     ids, distances (f32 bits) and object ids from the unmodified
                                     reference: the bench's parity block compares the GPU's ids
                                     against them without calling the checker.
"""
import argparse
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ora  # noqa: E402
from vk_cinematic_b200 import workloads as W  # noqa: E402
from vk_cinematic_b200.fixtures import tile_crcs, LATTICE  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--only", default="c2,c3,ties")
    args = ap.parse_args()
    if not ora.have_ref():
        ora.build_oracles()
    if "c3" in args.only:
        wl = W.config3()
        out = {"width": wl.width, "height": wl.height, "spp": wl.spp, "frame": 0}
        for key, lib, bounces in (("port_dm_5b", ora.load_port_dm(), 5), ("ref_dm_3b", ora.load_ref_dm(), 3),
                                  ("ref_3b", ora.load_ref(), 3)):
            s = lib.scene().load_workload(wl)
            t0 = time.time()
            img, m = s.render_seeded(spp=wl.spp, bounces=bounces, frame=0, threads=args.threads)
            s.close()
            print(f"{key}: {time.time() - t0:.1f} s, rays {int(m[2])}", flush=True)
            out["metrics_" + key] = m[1:5].copy()
            if key == "ref_3b":
                out["lattice_ref_3b"] = img[LATTICE // 2::LATTICE, LATTICE // 2::LATTICE, 0:3].copy()
                tw = th = 64
                tx, ty = (wl.width + tw - 1) // tw, (wl.height + th - 1) // th
                sums = np.zeros((tx * ty, 3), np.float64)
                for j in range(ty):
                    for i in range(tx):
                        sums[j * tx + i] = img[j * th:(j + 1) * th, i * tw:(i + 1) * tw, 0:3].astype(np.float64).sum(axis=(0, 1))
                out["tile_sum_ref_3b"] = sums
            else:
                out["tile_crc_" + key] = tile_crcs(img)
            del img
        np.savez_compressed(os.path.join(OUT, "g5_c3_full_frame.npz"), **out)
    if "ties" in args.only:
        # added to an existing g5 file (the renders above take minutes)
        path = os.path.join(OUT, "g5_c3_full_frame.npz")
        out = dict(np.load(path))
        wl = W.config3()
        s = ora.load_port_dm().scene().load_workload(wl)
        for key, bounces in (("ties_5b", 5), ("ties_3b", 3)):
            t0 = time.time()
            mask = s.tie_mask(spp=wl.spp, bounces=bounces, frame=0, threads=args.threads)
            ys, xs = np.nonzero(mask)
            out[key] = np.stack([xs, ys], axis=1).astype(np.uint16)
            print(f"{key}: {time.time() - t0:.1f} s, tie pixels {len(xs)}", flush=True)
        s.close()
        np.savez_compressed(path, **out)
    if "c2" in args.only:
        wl = W.config2()
        s = ora.load_ref().scene().load_workload(wl)
        hits = s.primary_hits(sample=0, frame=0)
        s.close()
        out = {"width": wl.width, "height": wl.height,
               "tile_crc_tri": tile_crcs(hits["tri"]), "tile_crc_t": tile_crcs(hits["t"]),
               "tile_crc_obj": tile_crcs(hits["obj"]), "hit_pixels": int((hits["tri"] >= 0).sum())}
        np.savez_compressed(os.path.join(OUT, "g6_c2_full_frame.npz"), **out)
        print("c2: hit pixels", out["hit_pixels"], flush=True)


if __name__ == "__main__":
    main()
