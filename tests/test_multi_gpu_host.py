"""Host-side logic of the multi-GPU path on CPU: stripe partition and the world_size-2 gather (gloo)."""
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
    assert grt.interleaved_stripes(2160, 0, 8, 54)[:2] == [(0, 54), (432, 486)]
    assert len(grt.interleaved_stripes(2160, 3, 8, 54)) == 5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, h, w, rows, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench

    full = torch.arange(h * w * 4, dtype=torch.float32).reshape(h, w, 4)       # the frame every rank would agree on
    stripes = bench.grt.interleaved_stripes(h, rank, world, rows)
    local = torch.cat([full[a:b] for a, b in stripes], 0)
    bufs = {}
    for _ in range(2):      # twice: the second call reuses the persistent buffers
        frame = bench.gather_frame(local, h, w, rank, world, rows, torch.device("cpu"), bufs)
    if rank == 0:
        q.put(bool(torch.equal(frame, full)))
    dist.destroy_process_group()


@pytest.mark.parametrize("h, rows", [(100, 16), (128, 16)])   # ragged last stripe / regular interleave (single strided copy)
def test_gather_reassembles_the_frame_world_size_2(h, rows):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, h, 24, rows, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
