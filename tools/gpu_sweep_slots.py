#!/usr/bin/env python3
"""Sustained ms/frame of the 4K headline frame against frames in flight and blocks per SM (24 frames per point)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import geodesic_raytracing_b200 as grt
content = grt.Content(grt.CONTENT_DIR)
metric = content.load_by_name("kerr_schild")
ctx = grt.Context(0)
ctx.set_metric(metric)
ctx.set_background(grt.synthetic_background(4096, 2048))
ctx.set_camera()
outs = [torch.empty((2160, 3840, 4), dtype=torch.uint8, device="cuda:0") for _ in range(8)]
K = int(sys.argv[1]) if len(sys.argv) > 1 else 24
for slots in (2, 3, 4, 5, 6, 8):
    ctx.set_frames_in_flight(slots)
    row = []
    for blocks in (1, 2, 3, 4):
        ctx.set_tuning(blocks, 0)
        best = 1e9
        for rep in range(2):
            ctx.synchronize(); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e0.record()
            ends = []
            for k in range(K):
                ctx.render(3840, 2160, grt.GRT_FORMAT_RGBA8, out=outs[k % slots], stats=False)
                if k >= K - slots:
                    e = torch.cuda.Event(enable_timing=True); e.record(torch.cuda.ExternalStream(ctx.stream)); ends.append(e)
            ctx.synchronize(); torch.cuda.synchronize()
            best = min(best, max(e0.elapsed_time(e) for e in ends) / K)
        row.append(best)
    print(f"slots {slots}: " + "  ".join(f"blocks {b}: {t:6.2f}" for b, t in zip((1, 2, 3, 4), row)), flush=True)
