"""Tile-row strip partition (host logic) and the strip gather over a 2-rank gloo group (CPU)."""
import os
import socket
import subprocess
import sys

import numpy as np

from vk_cinematic_b200 import strips

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check_cover(bounds, height, tile_h):
    assert bounds[0][0] == 0 and bounds[-1][1] == height
    for (b0, e0), (b1, e1) in zip(bounds, bounds[1:]):
        assert e0 == b1
    for b, e in bounds:
        assert b <= e and (b % tile_h == 0 or b == height) and (e % tile_h == 0 or e == height)


def test_partition_uniform():
    # C3: 2160 rows, 64-row tiles -> 34 tile rows
    for world in (1, 2, 4, 8):
        b = strips.partition_rows(2160, 64, world)
        check_cover(b, 2160, 64)
        rows = [(e - s + 63) // 64 for s, e in b]
        assert sum(rows) == 34 and max(rows) - min(rows) <= 1
    b = strips.partition_rows(100, 64, 4)  # fewer tile rows than ranks: trailing strips empty
    check_cover(b, 100, 64)
    assert [e - s for s, e in b] == [64, 36, 0, 0]
    assert strips.partition_rows(0, 64, 2) == [(0, 0), (0, 0)]


def test_partition_by_cost_balances():
    rng = np.random.RandomState(5)
    for _ in range(50):
        rows = int(rng.randint(1, 60))
        world = int(rng.choice([1, 2, 4, 8]))
        height = rows * 64 - int(rng.randint(0, 63))
        cost = rng.uniform(0.1, 10.0, rows) ** 2
        b = strips.partition_rows(height, 64, world, cost)
        check_cover(b, height, 64)
        if rows >= 4 * world:
            sums = [cost[s // 64:(e + 63) // 64].sum() for s, e in b]
            uniform = strips.partition_rows(height, 64, world)
            usums = [cost[s // 64:(e + 63) // 64].sum() for s, e in uniform]
            assert max(sums) <= max(usums) + cost.max()
    # a heavy band in the middle (an object in the centre of the frame) shrinks the middle strips
    cost = np.ones(34)
    cost[12:22] = 10
    b = strips.partition_rows(2160, 64, 4, cost)
    assert (b[1][1] - b[1][0]) < (b[0][1] - b[0][0])


def test_partition_is_the_minmax_optimum():
    """sp_b200_PartitionRows against brute force over every cut position (small cases)."""
    import itertools
    rng = np.random.RandomState(11)
    for _ in range(40):
        rows, world = int(rng.randint(4, 11)), int(rng.choice([2, 3, 4]))
        cost = rng.uniform(0.0, 5.0, rows) ** 3
        b = strips.partition_rows(rows * 16, 16, world, cost)
        got = max(cost[s // 16:e // 16].sum() for s, e in b)
        best = min(max(cost[a:c].sum() for a, c in zip((0,) + cuts, cuts + (rows,)))
                   for cuts in itertools.combinations(range(1, rows), world - 1))
        assert got <= best * (1 + 1e-12) + 1e-12


def test_strip_cost_roundtrip():
    bounds = strips.partition_rows(2160, 64, 4)
    per = [np.arange((e - 1) // 64 - s // 64 + 1) + 10 * r for r, (s, e) in enumerate(bounds)]
    rc = strips.strip_cost_to_row_cost(bounds, per, 2160, 64)
    assert len(rc) == 34 and rc[0] == 0 and rc[bounds[1][0] // 64] == 10


def test_gather_two_ranks_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "_strip_worker.py")]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert p.returncode == 0 and "STRIPS_OK" in p.stdout, p.stdout + p.stderr


def test_partition_is_minmax_optimal():
    """With a cost array the cut is the contiguous partition whose largest strip sum is minimal
    (checked against brute force on small cases) -- the step time is the slowest rank's."""
    import itertools
    rng = np.random.RandomState(11)
    for _ in range(40):
        rows = int(rng.randint(3, 11))
        world = int(rng.randint(2, min(rows, 4) + 1))
        if rows <= world:
            continue
        cost = rng.uniform(0.05, 5.0, rows) ** 2
        b = strips.partition_rows(rows * 16, 16, world, cost)
        check_cover(b, rows * 16, 16)
        assert all(e > s for s, e in b)
        got = max(cost[s // 16:e // 16].sum() for s, e in b)
        brute = min(max(cost[a:c].sum() for a, c in zip((0,) + cuts, cuts + (rows,)))
                    for cuts in itertools.combinations(range(1, rows), world - 1))
        assert got <= brute * (1 + 1e-12)


def test_rebalancing_converges_at_eight_ranks():
    """The bench's loop in miniature: per-row cost units with a model error (cheap sky rows
    overrated), scaled per rank by its measured seconds (strips.gather_row_costs), re-cut every
    frame.  C3-like profile (expensive rows in the middle of 2160 / 16 = 135 rows), 8 ranks, fixed
    per-rank overhead: after the bench's 8 warm-up frames the slowest rank is within 8 % of the
    mean, and never worse than the even split it started from."""
    H, TH, world = 2160, 16, 8
    rows = H // TH
    y = (np.arange(rows) + 0.5) / rows
    obj = np.exp(-((y - 0.5) / 0.18) ** 2)
    true_ms = 0.08 + obj
    for overrate in (1.0, 1.3, 5.0):
        units = true_ms * np.where(obj < 0.3, overrate, 1.0)
        bounds = strips.partition_rows(H, TH, world)
        history = []
        for _ in range(8):
            secs = [true_ms[b // TH:e // TH].sum() + 0.15 for b, e in bounds]
            history.append(max(secs) / np.mean(secs))
            cost = np.zeros(rows)
            for (b, e), s in zip(bounds, secs):
                seg = units[b // TH:e // TH]
                cost[b // TH:e // TH] = seg * (s / seg.sum())
            bounds = strips.partition_rows(H, TH, world, cost)
            check_cover(bounds, H, TH)
        secs = [true_ms[b // TH:e // TH].sum() + 0.15 for b, e in bounds]
        final = max(secs) / np.mean(secs)
        assert final <= 1.08 and final <= history[0], (overrate, history, final)
