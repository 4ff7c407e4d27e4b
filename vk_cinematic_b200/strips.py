"""Tile-row strips across the GPUs of one box.

The reference's only parallelism is 64x64 tiles popped from an atomic queue by 16 threads over a
shared read-only scene (reference src/main.cpp:731-759,819-844, src/tile.h:11-42).  ComputeTiles
orders tiles row-major, so a contiguous range of tile ROWS is a contiguous byte range of the
RGBA-f32 image.  Here every rank (one process per GPU) holds the whole scene, renders one strip
of whole tile rows into its own image buffer and the finished strips are gathered on rank 0 --
the only exchange step of the path.  Strip boundaries are re-cut between frames from the
measured per-tile-row cost.

Host logic only: the cut is made by the library (sp_b200_PartitionRows); the transport is
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np


def tile_row_count(height, tile_h):
    return (height + tile_h - 1) // tile_h


def partition_rows(height, tile_h, world, cost=None):
    """Cut `height` pixel rows into `world` contiguous strips of whole tile rows.

    cost: optional per-tile-row cost (length tile_row_count): the cut then minimises the largest
    strip's summed cost (the step takes as long as the slowest rank); every strip keeps at least
    one tile row while rows last.  The arithmetic is the library's (sp_b200_PartitionRows,
    csrc/spb_strips.cpp): the same cut the single-process multi-device path of libspb200 makes.
    Returns a list of (row_begin, row_end) pixel rows, one per rank; trailing ranks may be empty
    when there are fewer tile rows than ranks."""
    from . import sp
    bounds = np.zeros(world + 1, np.uint32)
    c = None
    if cost is not None:
        c = np.ascontiguousarray(cost, dtype=np.float64)
        assert len(c) == tile_row_count(height, tile_h)
    sp.lib.sp_b200_PartitionRows(height, tile_h, world, c.ctypes.data if c is not None else None, bounds.ctypes.data)
    return [(int(bounds[i]), int(bounds[i + 1])) for i in range(world)]


def strip_cost_to_row_cost(bounds, strip_costs, height, tile_h):
    """Concatenate the per-strip tile-row cost arrays (what sp_b200_RenderRows returns on each
    rank) into one per-tile-row array for the whole image."""
    rows = tile_row_count(height, tile_h)
    out = np.zeros(rows, dtype=np.float64)
    for (b, e), c in zip(bounds, strip_costs):
        if e <= b:
            continue
        first = b // tile_h
        c = np.asarray(c, dtype=np.float64)
        out[first:first + len(c)] += c
    return out


def gather_row_costs(local_cost, seconds, bounds, height, tile_h, dist, device):
    """All ranks learn every tile row's cost in SECONDS: rays per tile row scaled by the rank's
    measured seconds per ray (tiny all-gather of metadata; the pixel gather is gather_strips)."""
    import torch
    world = dist.get_world_size()
    rows = tile_row_count(height, tile_h)
    mine = torch.zeros(rows + 1, dtype=torch.float64, device=device)
    b, e = bounds[dist.get_rank()]
    if e > b:
        first = b // tile_h
        c = torch.as_tensor(np.asarray(local_cost, dtype=np.float64), device=device)
        rays = float(c.sum().item())
        scale = seconds / rays if rays > 0 else 0.0
        mine[first:first + len(c)] = c * scale
    mine[rows] = seconds
    everyone = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(everyone, mine)
    stacked = torch.stack(everyone).cpu().numpy()
    return stacked[:, :rows].sum(axis=0), stacked[:, rows]


def gather_strips(image, bounds, dist, dst=0):
    """image: (H, W, 4) float32 tensor on every rank (only the rank's strip rows are valid).
    After the call rank `dst` holds every strip.  One batched send/recv group: strips are unequal
    after rebalancing, so this is point-to-point rather than a padded all-gather."""
    rank, world = dist.get_rank(), dist.get_world_size()
    ops = []
    if rank == dst:
        for r in range(world):
            b, e = bounds[r]
            if r != dst and e > b:
                ops.append(dist.P2POp(dist.irecv, image[b:e], r))
    else:
        b, e = bounds[rank]
        if e > b:
            ops.append(dist.P2POp(dist.isend, image[b:e], dst))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return image
