"""Host-side logic of the multi-GPU path on CPU: the stripe partition (computed in libgrt) and the world_size-2
exchange of the communicator handles (gloo).  The delivery itself - peer stores into the owner's frame - is tested with -m gpu."""
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
    q.put((rank, got == want, int(rows.item())))
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
