#!/usr/bin/env python3
"""Sweep the integration kernel's scheduling knobs on the 4K headline frame."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import geodesic_raytracing_b200 as grt  # noqa: E402


def main():
    content = grt.Content(grt.CONTENT_DIR)
    metric = content.load_by_name("kerr_schild")
    ctx = grt.Context(0)
    ctx.set_metric(metric, is_static=bool(int(os.environ.get("GRT_STATIC", "0"))))
    ctx.set_background(grt.synthetic_background(4096, 2048))
    ctx.set_camera()
    out = torch.empty((2160, 3840, 4), dtype=torch.uint8, device="cuda:0")
    ref = None
    combos = [(1, 0)] + [(p, t) for p in (4, 8, 12, 16, 24, 32) for t in (8, 16, 20, 24, 28, 31)]
    if len(sys.argv) > 1:
        combos = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
    for passes, thr in combos:
        ctx.set_tuning(passes, thr)
        best = None
        for _ in range(3):
            _, st = ctx.render(3840, 2160, grt.GRT_FORMAT_RGBA8, out=out)
            if best is None or st.ms_total < best.ms_total:
                best = st
        # pipelined throughput: 8 frames queued back to back, two in flight
        outs = [out] + [torch.empty_like(out) for _ in range(3)]
        ctx.synchronize(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        ends = []
        nfl = int(os.environ.get('GRT_FRAME_SLOTS', '2'))
        for k in range(12):
            ctx.render(3840, 2160, grt.GRT_FORMAT_RGBA8, out=outs[k % 4], stats=False)
            if k >= 12 - nfl:
                e = torch.cuda.Event(enable_timing=True); e.record(torch.cuda.ExternalStream(ctx.stream)); ends.append(e)
        ctx.synchronize(); torch.cuda.synchronize()
        pipelined = max(e0.elapsed_time(e) for e in ends) / 12
        img = out.cpu()
        same = True if ref is None else bool(torch.equal(img, ref))
        if ref is None:
            ref = img
        log = ctx.launch_log()
        print(f"passes {passes:2d} thr {thr:2d}: total {best.ms_total:7.2f} ms  prepass {log['prepass']:6.2f} integrate {log['integrate']:6.2f} extra {log['integrate_extra']:6.2f} "
              f"steps {best.trial_steps} identical={same} pipelined {pipelined:6.2f} ms/frame", flush=True)


if __name__ == "__main__":
    main()
