"""Host-side logic of the multi-GPU path on CPU: the partitions of a frame (computed in libgrt), the world_size-2
exchange of the communicator handles and the feedback loop of the contiguous partition (gloo).  The delivery itself - peer stores into the owner's frame - is tested with -m gpu."""
import os
import socket

import numpy as np
import pytest


def test_interleaved_stripes_partition_the_frame(grt):
    for h in (2160, 1080, 144, 100):
        for world in (1, 2, 4, 8):
            for rows in (54, 16, 30):
                seen = np.zeros(h, int)
                for r in range(world):
                    for (y0, y1) in grt.interleaved_stripes(h, r, world, rows):
                        assert y0 % 2 == 0 and (y1 % 2 == 0 or y1 == h) and y0 < y1 <= h
                        seen[y0:y1] += 1
                assert (seen == 1).all(), (h, world, rows)
    assert grt.interleaved_stripes(2160, 0, 8, 54)[:3] == [(0, 54), (702, 756), (972, 1026)]      # position 0, then 5, then 2: rotating by three
    assert all(len(grt.interleaved_stripes(2160, r, 8, 54)) == 5 for r in range(8))
    assert grt.interleaved_stripes(2160, 1, 2, 54)[:3] == [(54, 108), (108, 162), (270, 324)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench

    # the host side of the multi-GPU path: every rank contributes its 64-byte communicator handle and receives all of them in rank
    # order (what grt_comm_connect takes), and computes its own stripes; the frame delivery itself is device code (tests -m gpu)
    handle = bytes([rank * 16 + (i % 16) for i in range(bench.grt.GRT_COMM_HANDLE_BYTES)])
    for _ in range(2):
        got = bench.exchange_handles(handle, rank, world, torch.device("cpu"), dist)
    want = b"".join(bytes([r * 16 + (i % 16) for i in range(bench.grt.GRT_COMM_HANDLE_BYTES)]) for r in range(world))
    stripes = bench.grt.interleaved_stripes(2160, rank, world, bench.STRIPE_ROWS)
    rows = torch.tensor([sum(b - a for a, b in stripes)], dtype=torch.int64)
    dist.all_reduce(rows)
    # the feedback loop of the contiguous partition as bench.py runs it: every rank "measures" its own block (here: a synthetic cost per
    # row), the numbers are all-gathered, and every rank must arrive at the same boundaries
    yy = (np.arange(2160) + 0.5) / 2160
    row = 0.5 + 4.0 * np.exp(-((yy - 0.36) / 0.02) ** 2) + 4.0 * np.exp(-((yy - 0.64) / 0.02) ** 2) + (0.3 if world == 2 else 0.0) * yy
    halves = [(i * 54 + (0 if k == 0 else 26), i * 54 + (26 if k == 0 else 54)) for i in range(40) for k in (0, 1)]
    profile = [float(row[a:b].sum()) for a, b in halves]
    bounds = bench.grt.contiguous_bounds(2160, world, profile, 54)
    for _ in range(3):
        mine = torch.tensor([float(row[bounds[rank]:bounds[rank + 1]].sum())], dtype=torch.float64)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        bounds = bench.grt.refine_bounds(2160, bounds, [float(p.item()) for p in parts], profile, 54)
    loads = [float(row[bounds[r]:bounds[r + 1]].sum()) for r in range(world)]
    all_bounds = [None] * world
    dist.all_gather_object(all_bounds, bounds)
    balanced = all(b == bounds for b in all_bounds) and max(loads) / (sum(loads) / world) < 1.03
    q.put((rank, got == want and balanced, int(rows.item())))
    dist.destroy_process_group()


def test_handle_exchange_and_stripes_world_size_2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = sorted(q.get(timeout=5) for _ in range(2))
    assert results == [(0, True, 2160), (1, True, 2160)]


def test_c_stripe_function_refuses_bad_requests(grt):
    import ctypes as C
    y0, y1 = (C.c_int * 4)(), (C.c_int * 4)()
    L = grt.lib()
    assert L.grt_interleaved_stripes(2160, 0, 8, 54, y0, y1, 4) < 0          # five stripes do not fit four slots
    assert L.grt_interleaved_stripes(2160, 8, 8, 54, y0, y1, 4) < 0          # rank outside the world
    assert L.grt_interleaved_stripes(100, 0, 1, 54, y0, y1, 4) == 1 and (y0[0], y1[0]) == (0, 100)


def test_communicator_calls_fail_loudly_without_a_context(grt):
    L = grt.lib()
    assert L.grt_comm_init(None, 0, 2, 64, 64, 1, 2, None) < 0
    assert L.grt_comm_render(None, None, None, 0, 0, None) < 0
    assert L.grt_comm_frame(None, 0) is None


def test_balanced_stripes_deal_by_cost():
    """grt_balanced_stripes: the ranks' stripes tile the frame exactly once, the heaviest rank carries far less than with plain
    interleaving (stripes that dominate a rank's share are dealt as halves), the same input gives the same answer on every rank, a
    missing cost vector falls back to the interleaved pattern."""
    import ctypes as C

    import numpy as np

    import geodesic_raytracing_b200 as grt
    h, rows = 2160, 54
    n = h // rows
    y = (np.arange(2 * n) + 0.5) / (2 * n)                     # half stripes
    # a ring-like profile: cheap sky, two narrow expensive bands where the rows are tangent to the shadow's edge, cheap culled centre
    cost = 0.5 + 2.5 * np.exp(-((y - 0.36) / 0.03) ** 2) + 2.5 * np.exp(-((y - 0.64) / 0.03) ** 2)    # heaviest stripe: 0.7 of a fair share of 8
    row_cost = np.repeat(cost / 27.0, 27)[:h]                   # per row, for scoring any partition
    for world in (2, 4, 8):
        owned = [grt.balanced_stripes(h, r, world, cost.tolist(), rows) for r in range(world)]
        covered = np.zeros(h, int)
        for o in owned:
            assert o == sorted(o)
            for y0, y1 in o:
                assert y0 % 2 == 0 and y1 % 2 == 0 and y1 > y0
                covered[y0:y1] += 1
        assert (covered == 1).all()
        load = [sum(row_cost[y0:y1].sum() for y0, y1 in o) for o in owned]
        plain = [sum(row_cost[y0:y1].sum() for y0, y1 in grt.interleaved_stripes(h, r, world, rows)) for r in range(world)]
        assert max(load) / np.mean(load) < 1.06, (world, load)
        if world == 8:
            assert max(load) < 0.85 * max(plain)
        assert owned == [grt.balanced_stripes(h, r, world, cost.tolist(), rows) for r in range(world)]
    L = grt.lib()
    y0, y1 = (C.c_int * 64)(), (C.c_int * 64)()
    k = L.grt_balanced_stripes(h, 3, 8, rows, None, 0, y0, y1, 64)
    assert [(y0[i], y1[i]) for i in range(k)] == grt.interleaved_stripes(h, 3, 8, rows)
    bad = (C.c_double * 3)(1.0, 2.0, 3.0)
    assert L.grt_balanced_stripes(h, 0, 8, rows, bad, 3, y0, y1, 64) < 0          # cost vector of another frame shape
    assert grt.balanced_stripes(100, 0, 1, [5.0, 1.0, 2.0, 2.0], rows) == [(0, 100)]


def test_contiguous_bounds_and_feedback():
    """grt_contiguous_bounds / grt_refine_bounds: even, ascending boundaries from 0 to the height; equal-cost blocks for a peaked
    profile; the feedback rounds bring the real loads of a finer-grained truth within a few percent; bad input is refused."""
    import ctypes as C

    import numpy as np

    import geodesic_raytracing_b200 as grt
    h, rows = 2160, 54
    n = h // rows
    yy = (np.arange(h) + 0.5) / h
    row = 0.5 + 4.0 * np.exp(-((yy - 0.36) / 0.015) ** 2) + 4.0 * np.exp(-((yy - 0.64) / 0.015) ** 2)       # true cost per row, sharper than the profile resolves
    halves = [(i * rows + (0 if k == 0 else 26), i * rows + (26 if k == 0 else rows)) for i in range(n) for k in (0, 1)]
    profile = [row[a:b].sum() for a, b in halves]
    assert grt.contiguous_bounds(h, 8) == [270 * r for r in range(9)]
    for world in (2, 3, 4, 8):
        b = grt.contiguous_bounds(h, world, profile, rows)
        for _ in range(3):
            assert b[0] == 0 and b[-1] == h and all(x % 2 == 0 for x in b) and all(b[r + 1] - b[r] >= 2 for r in range(world))
            load = [row[b[r]:b[r + 1]].sum() for r in range(world)]
            b = grt.refine_bounds(h, b, load, profile, rows)
        load = [row[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(load) / np.mean(load) < 1.05, (world, b, load)
    # a degenerate profile (all the cost in one half stripe) still gives every rank two rows
    spike = [0.0] * (2 * n)
    spike[40] = 1.0
    b = grt.contiguous_bounds(h, 8, spike, rows)
    assert b[0] == 0 and b[-1] == h and all(b[r + 1] - b[r] >= 2 for r in range(8))
    L = grt.lib()
    out = (C.c_int * 9)()
    bad = (C.c_double * 3)(1.0, 2.0, 3.0)
    assert L.grt_contiguous_bounds(h, 8, rows, bad, 3, out) < 0
    assert L.grt_contiguous_bounds(10, 8, rows, None, 0, out) < 0                     # fewer than two rows per rank
    desc = (C.c_int * 3)(0, 1200, 1000)
    m = (C.c_double * 2)(1.0, 1.0)
    assert L.grt_refine_bounds(h, 2, rows, None, 0, desc, m, out) < 0
    assert grt.refine_bounds(h, [0, 1000, h], [3.0, 1.0])[1] < 1000                    # without a profile: uniform inside a block
