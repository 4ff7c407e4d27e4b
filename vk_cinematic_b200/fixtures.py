"""Compact whole-frame fingerprints: per-tile CRC-32s in the reference's tile order.

A 3840x2160 RGBA-f32 frame is 133 MB; its 2040 64x64 tiles (ComputeTiles order, reference
src/tile.h:11-42) hash to 8 KB.  tools/make_full_frame_golden.py stores the CPU checkers' tile
CRCs under tests/golden/; the GPU tests and bench.py's `parity` block hash the frame the library
rendered and compare -- every pixel of the bench configuration is pinned bit for bit without
shipping the frame or calling the checker at run time (nothing of the CPU checkers is used here).
"""
import zlib

import numpy as np

TILE = 64
LATTICE = 16   # the plain-reference lattice fixture keeps pixel (8 + 16 i, 8 + 16 j)


def tile_crcs(image, tile=TILE):
    """CRC-32 of every tile x tile block of image[H, W(, C)] (rows top to bottom inside a tile, tiles
    row-major, edge tiles clamped to the image like ComputeTiles does)."""
    a = np.ascontiguousarray(image)
    h, w = a.shape[0], a.shape[1]
    tx, ty = (w + tile - 1) // tile, (h + tile - 1) // tile
    out = np.zeros(tx * ty, np.uint32)
    for j in range(ty):
        band = a[j * tile:(j + 1) * tile]
        for i in range(tx):
            out[j * tx + i] = zlib.crc32(np.ascontiguousarray(band[:, i * tile:(i + 1) * tile]).tobytes())
    return out


def lattice(image, step=LATTICE):
    return np.ascontiguousarray(image)[step // 2::step, step // 2::step, 0:3]


def relative_error_report(image_rgb, reference_rgb, tolerances=(1e-6, 1e-5, 1e-4, 1e-3, 1e-2)):
    """Per-pixel relative error max_c |a - b| / max(|b|, 1e-3 mean(b)) of two (h, w, 3) arrays, its
    histogram against `tolerances`, and the RMSE relative to the reference's mean radiance."""
    a = image_rgb.astype(np.float64)
    b = reference_rgb.astype(np.float64)
    floor = 1e-3 * max(float(np.abs(b).mean()), 1e-30)
    rel = (np.abs(a - b) / np.maximum(np.abs(b), floor)).max(axis=-1)
    rmse = float(np.sqrt(((a - b) ** 2).mean()))
    mean = float(np.abs(b).mean())
    return {"pixels": int(rel.size), "identical": int((rel == 0).sum()),
            "fraction_above": {f"{t:g}": float((rel > t).mean()) for t in tolerances},
            "max_relative": float(rel.max()), "rmse": rmse, "rmse_over_mean": rmse / max(mean, 1e-30)}
