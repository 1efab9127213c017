#!/usr/bin/env python3
"""Frames per second of a few bundled metrics at 1920x1080 (4 frame slots, reference default mode), to compare the occupancy
tier chosen at program build (default) with the free compile (GRT_OCCUPANCY_TIER=0 in the environment)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import geodesic_raytracing_b200 as grt  # noqa: E402


def main():
    content = grt.Content(grt.CONTENT_DIR)
    ctx = grt.Context(0)
    ctx.set_background(grt.synthetic_background(4096, 2048))
    ctx.set_frames_in_flight(4)
    ctx.set_tuning(3, 0)
    w, h = 1920, 1080
    outs = [torch.empty((h, w, 4), dtype=torch.uint8, device="cuda:0") for _ in range(4)]
    for name in sys.argv[1:] or ["schwarzschild", "alcubierre", "kerr_boyer", "double_schwarzschild", "kerr_schild"]:
        m = content.load_by_name(name)
        ctx.set_metric(m)
        ctx.set_camera()
        ctx.set_features(max_acceleration_change=m.info.max_acceleration_change, redshift=1 if name == "alcubierre" else 0)
        _, st = ctx.render(w, h, grt.GRT_FORMAT_RGBA8, out=outs[0])
        best = 1e9
        for rep in range(3):
            ctx.synchronize(); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e0.record()
            ends = []
            n = 16
            for k in range(n):
                ctx.render(w, h, grt.GRT_FORMAT_RGBA8, out=outs[k % 4], stats=False)
                if k >= n - 4:
                    e = torch.cuda.Event(enable_timing=True); e.record(torch.cuda.ExternalStream(ctx.stream)); ends.append(e)
            ctx.synchronize(); torch.cuda.synchronize()
            best = min(best, max(e0.elapsed_time(e) for e in ends) / n)
        info = ctx.kernel_info()
        print(f"tier={os.environ.get('GRT_OCCUPANCY_TIER', 'auto'):4s} {name:24s} regs {info['k_integrate'][0]:3d} blocks/SM {info['integrate_blocks_per_sm'][0]:2d}  "
              f"{best:7.3f} ms/frame  {st.trial_steps / best / 1e6:9.1f} G ray-steps/s  one frame {st.ms_total:6.2f} ms", flush=True)


if __name__ == "__main__":
    main()
