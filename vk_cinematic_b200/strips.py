"""Tile-row strips across the GPUs of one box.

The reference's only parallelism is 64x64 tiles popped from an atomic queue by 16 threads over a
shared read-only scene (reference src/main.cpp:731-759,819-844, src/tile.h:11-42).  ComputeTiles
orders tiles row-major, so a contiguous range of tile ROWS is a contiguous byte range of the
RGBA-f32 image.  Here every rank (one process per GPU) holds the whole scene, renders one strip
of whole tile rows into its own image buffer and the finished strips are gathered on rank 0 --
the only exchange step of the path.  Strip boundaries are re-cut between frames from the
measured per-tile-row cost.

Host logic only; the transport is torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU
tests).
"""
import numpy as np


def tile_row_count(height, tile_h):
    return (height + tile_h - 1) // tile_h


def partition_rows(height, tile_h, world, cost=None):
    """Cut `height` pixel rows into `world` contiguous strips of whole tile rows.

    cost: optional per-tile-row cost (length tile_row_count); strips get near-equal cost sums
    (prefix-sum cuts), every strip keeps at least one tile row while rows last.
    Returns a list of (row_begin, row_end) pixel rows, one per rank; trailing ranks may be empty
    when there are fewer tile rows than ranks."""
    rows = tile_row_count(height, tile_h)
    if cost is not None and rows > world > 1:
        cuts = _minmax_cuts(np.asarray(cost, dtype=np.float64), world)
        return [(min(cuts[i] * tile_h, height), min(cuts[i + 1] * tile_h, height)) for i in range(world)]
    if cost is None:
        cost = np.ones(rows, dtype=np.float64)
    cost = np.asarray(cost, dtype=np.float64)
    assert len(cost) == rows
    cost = np.maximum(cost, 1e-9 * max(1.0, float(cost.max()) if rows else 1.0))
    prefix = np.concatenate([[0.0], np.cumsum(cost)])
    total = prefix[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        # first boundary whose prefix is closest to the target
        k = int(np.searchsorted(prefix, target))
        if k > 0 and abs(prefix[k - 1] - target) <= abs(prefix[min(k, rows)] - target):
            k -= 1
        lo = min(cuts[-1] + 1, rows)              # at least one tile row per strip ...
        hi = max(lo, rows - (world - r))          # ... and leave one for each later strip
        cuts.append(int(min(max(k, lo), hi)))
    cuts.append(rows)
    return [(min(cuts[i] * tile_h, height), min(cuts[i + 1] * tile_h, height)) for i in range(world)]


def _minmax_cuts(cost, world):
    """Contiguous partition of `cost` into `world` non-empty strips with the smallest possible
    largest strip sum (the step time is the slowest rank's): dynamic programme over (strips used,
    rows covered), vectorised over the position of the last cut.  The prefix-closest-to-target
    cuts used before sit up to half a row off per boundary, which at 8 strips of ~17 rows each was
    a 10 % imbalance on a centre-heavy frame; this is the optimum for the given granularity."""
    rows = len(cost)
    assert rows >= world
    cost = np.maximum(cost, 1e-9 * max(1.0, float(cost.max())))
    prefix = np.concatenate([[0.0], np.cumsum(cost)])
    best = np.full((world + 1, rows + 1), np.inf)
    arg = np.zeros((world + 1, rows + 1), dtype=np.int64)
    best[0, 0] = 0.0
    for k in range(1, world + 1):
        for i in range(k, rows - (world - k) + 1):
            j = np.arange(k - 1, i)
            v = np.maximum(best[k - 1, k - 1:i], prefix[i] - prefix[j])
            m = int(np.argmin(v))
            best[k, i] = v[m]
            arg[k, i] = j[m]
    cuts = [rows]
    i = rows
    for k in range(world, 0, -1):
        i = int(arg[k, i])
        cuts.append(i)
    return cuts[::-1]


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
