"""Worker of tests/test_strips.py: 2 ranks over gloo, CPU tensors.  Each rank fills its strip
with a rank-dependent pattern; rank 0 gathers and checks, then both re-cut from a skewed cost."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vk_cinematic_b200 import strips  # noqa: E402  (loads libspb200.so for the cut; no compute call, no GPU)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    H, W, TH = 200, 33, 64
    bounds = strips.partition_rows(H, TH, world)
    image = torch.zeros(H, W, 4)
    b, e = bounds[rank]
    ys = torch.arange(b, e, dtype=torch.float32)
    image[b:e] = (ys[:, None, None] * 1000 + torch.arange(W)[None, :, None] + rank * 0.25)
    strips.gather_strips(image, bounds, dist)
    if rank == 0:
        for r, (b, e) in enumerate(bounds):
            ys = torch.arange(b, e, dtype=torch.float32)
            exp = (ys[:, None, None] * 1000 + torch.arange(W)[None, :, None] + r * 0.25).expand(e - b, W, 4)
            assert torch.equal(image[b:e], exp), r
    # cost feedback: rank 0's rows are 3x as expensive per ray
    rows = strips.tile_row_count(H, TH)
    b, e = bounds[rank]
    first = b // TH
    local = np.full((bounds[rank][1] - 1) // TH - first + 1, 100.0)
    cost, secs = strips.gather_row_costs(local, 3.0 if rank == 0 else 1.0, bounds, H, TH, dist, "cpu")
    assert len(cost) == rows and abs(cost.sum() - 4.0) < 1e-9 and list(secs) == [3.0, 1.0]
    new = strips.partition_rows(H, TH, world, cost)
    assert new[0][1] <= bounds[0][1] and new[0][1] == new[1][0] and new[-1][1] == H
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("STRIPS_OK")


if __name__ == "__main__":
    main()
