#!/usr/bin/env python3
"""Stage times of one frame alone and sustained ms/frame with frames in flight, for any bundled metric and size.
   python tools/gpu_frame_probe.py --metric kerr_schild --size 3840 2160 --slots 4 --blocks 3 [--static] [--adaptive 0]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import geodesic_raytracing_b200 as grt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--metric", default="kerr_schild")
    ap.add_argument("--size", type=int, nargs=2, default=[3840, 2160])
    ap.add_argument("--slots", type=int, nargs="+", default=[4])
    ap.add_argument("--blocks", type=int, nargs="+", default=[3])
    ap.add_argument("--static", action="store_true")
    ap.add_argument("--adaptive", type=int, default=1)
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    w, h = a.size
    content = grt.Content(grt.CONTENT_DIR)
    metric = content.load_by_name(a.metric)
    ctx = grt.Context(0)
    ctx.set_metric(metric, is_static=a.static)
    ctx.set_features(max_acceleration_change=metric.info.max_acceleration_change, adaptive_sampling=a.adaptive)
    ctx.set_background(grt.synthetic_background(4096, 2048))
    ctx.set_camera()
    info = ctx.kernel_info()
    print(f"[{a.tag}] {a.metric} {w}x{h} k_integrate {info.get('k_integrate')} blocks/SM {info.get('integrate_blocks_per_sm')}", flush=True)
    outs = [torch.empty((h, w, 4), dtype=torch.uint8, device="cuda:0") for _ in range(max(a.slots))]
    for blocks in a.blocks:
        ctx.set_tuning(blocks, 0)
        best = None
        for _ in range(3):
            _, st = ctx.render(w, h, grt.GRT_FORMAT_RGBA8, out=outs[0])
            if best is None or st.ms_total < best[0].ms_total:
                best = (st, ctx.launch_log())
        st, log = best
        print(f"[{a.tag}] blocks {blocks}: alone total {st.ms_total:7.2f} ms prepass {log['prepass']:6.2f} integrate {log['integrate']:6.2f} extra {log['integrate_extra']:6.2f} "
              f"render {log['render']:5.2f} steps {st.trial_steps} rays {st.rays_integrated} extra_rays {st.extra_rays}", flush=True)
        for slots in a.slots:
            ctx.set_frames_in_flight(slots)
            for rep in range(2):
                ctx.synchronize(); torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e0.record()
                ends = []
                for k in range(a.frames):
                    ctx.render(w, h, grt.GRT_FORMAT_RGBA8, out=outs[k % slots], stats=False)
                    if k >= a.frames - slots:
                        e = torch.cuda.Event(enable_timing=True); e.record(torch.cuda.ExternalStream(ctx.stream)); ends.append(e)
                ctx.synchronize(); torch.cuda.synchronize()
                ms = max(e0.elapsed_time(e) for e in ends) / a.frames
            print(f"[{a.tag}] blocks {blocks} slots {slots}: sustained {ms:7.2f} ms/frame = {1e3 / ms:6.1f} fps, {st.trial_steps / ms / 1e6:8.2f} G ray-steps/s", flush=True)
        ctx.set_frames_in_flight(2)


if __name__ == "__main__":
    main()
