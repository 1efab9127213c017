#!/usr/bin/env python3
"""Multi-process check of the frame delivery (run under torchrun on N GPUs): every rank renders its stripes into rank 0's frame;
rank 0 compares the assembled frames with its own single-GPU frames, bit for bit.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import geodesic_raytracing_b200 as grt  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    content = grt.Content(grt.CONTENT_DIR)
    ok = True
    for name, (w, h), n_frames in (("kerr_schild", (640, 360), 2), ("double_unequal_kerr", (512, 288), 3)):
        metric = content.load_by_name(name)
        ctx = grt.Context(local)
        ctx.set_metric(metric)
        ctx.set_features(max_acceleration_change=metric.info.max_acceleration_change)
        ctx.set_background(grt.synthetic_background(1024, 512))
        ctx.set_frames_in_flight(3)
        cams = [(0.0, 0.0, -4.0 - 0.25 * i, 0.1 * i) for i in range(6)]
        wants = []
        if rank == 0:
            for c in cams:
                ctx.set_camera(c)
                img, _ = ctx.render(w, h, grt.GRT_FORMAT_RGBA8)
                wants.append(img.copy())
        handle = ctx.comm_init(rank, world, w, h, grt.GRT_FORMAT_RGBA8, n_frames)
        ctx.comm_connect(bench.exchange_handles(handle, rank, world, device, dist))
        if name == "kerr_schild":
            # dealt by measured cost: every rank measures the profile of the first view itself (deterministic) and must arrive at one partition
            ctx.set_camera(cams[0])
            costs = ctx.measure_stripe_costs(w, h, 18)
            stripes = grt.balanced_stripes(h, rank, world, costs, 18)
            mine = torch.zeros(h, dtype=torch.int32, device=device)
            for y0, y1 in stripes:
                mine[y0:y1] += 1
            dist.all_reduce(mine)
            ok &= bool((mine == 1).all().item())          # the ranks' stripes tile the frame exactly once
        elif world <= 4:
            # one contiguous block per rank with boundaries at equal measured cost, moved once by the blocks' real cost
            ctx.set_camera(cams[0])
            costs = ctx.measure_stripe_costs(w, h, 18)
            bounds = grt.contiguous_bounds(h, world, costs, 18)
            scratch = torch.empty((h, w, 4), dtype=torch.uint8, device=device)
            _, st = ctx.render(w, h, grt.GRT_FORMAT_RGBA8, out=scratch, stats=True, stripes=[(bounds[rank], bounds[rank + 1])])
            mine = torch.tensor([float(st.trial_steps)], dtype=torch.float64, device=device)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            bounds = grt.refine_bounds(h, bounds, [float(p.item()) for p in parts], costs, 18)
            stripes = [(bounds[rank], bounds[rank + 1])]
            if rank == 0:
                print(f"{name} world {world}: contiguous bounds {bounds}", flush=True)
        else:
            stripes = grt.interleaved_stripes(h, rank, world, 18)
        host = [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() for _ in cams] if rank == 0 else None
        for seq, c in enumerate(cams):
            ctx.set_camera(c)
            ctx.comm_render(stripes, seq, host[seq] if rank == 0 else None)
        ctx.synchronize()
        dist.barrier()
        if rank == 0:
            for seq in range(len(cams)):
                same = np.array_equal(host[seq].numpy(), wants[seq])
                ok &= same
                print(f"{name} world {world} frame {seq}: assembled == single-GPU frame: {same}", flush=True)
        ctx.comm_disconnect()
        dist.barrier()
        ctx.close()
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_CHECK", "OK" if ok else "FAILED", flush=True)
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
