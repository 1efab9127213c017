#!/usr/bin/env python3
"""First-contact GPU check: small frames per bundled metric against the CPU oracle, then the 4K headline timing."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import geodesic_raytracing_b200 as grt  # noqa: E402
import oracle  # noqa: E402


def main():
    content = grt.Content(grt.CONTENT_DIR)
    ctx = grt.Context(0)
    bg = grt.synthetic_background(2048, 1024)
    ctx.set_background(bg)
    ctx.set_camera()
    for fast in (1, 0):
        ctx.set_fast_math(bool(fast))
        for name, (w, h) in [("minkowski", (128, 128)), ("schwarzschild_fast", (256, 144)), ("kerr_schild", (256, 144)), ("alcubierre", (256, 144)),
                             ("double_schwarzschild", (256, 144))]:
            metric = content.load_by_name(name)
            t = time.time()
            ctx.set_metric(metric)
            t_build = time.time() - t
            for adaptive in (0, 1):
                ctx.set_features(adaptive_sampling=adaptive)
                img, st = ctx.render(w, h)
                orc = oracle.for_metric(metric)
                f = orc.features(max_acceleration_change=metric.info.max_acceleration_change, adaptive_sampling=adaptive)
                ref = orc.frame(w, h, f, background=bg)
                rd = ctx.read_render_data(w, h).reshape(h, w)
                agree = (rd["terminated"] == ref["rdata"]["terminated"]).mean()
                both = (rd["terminated"] == 1) & (ref["rdata"]["terminated"] == 1)
                duv = np.abs(rd["tex_coord"] - ref["rdata"]["tex_coord"])
                duv = np.minimum(duv, 1 - duv)[both]
                close = (np.abs(img[..., :3] - ref["image"][..., :3]).max(axis=-1) <= 2 / 255).mean()
                print(f"fast={fast} {name:22s} {w}x{h} adaptive={adaptive} build {t_build:.2f}s steps gpu {st.trial_steps} cpu {ref['trial_steps']} extra gpu {st.extra_rays} cpu {ref['extra_rays']} "
                      f"term-agree {agree:.4f} uv-med {np.median(duv) if duv.size else 0:.2e} uv-p99 {np.percentile(duv, 99) if duv.size else 0:.2e} px<=2/255 {close:.4f} "
                      f"ms {st.ms_total:.3f} int {st.ms_integrate:.3f}", flush=True)
        print(ctx.kernel_info(), flush=True)
    # headline
    metric = content.load_by_name("kerr_schild")
    ctx.set_fast_math(True)
    ctx.set_metric(metric)
    ctx.set_background(grt.synthetic_background(4096, 2048))
    import torch
    out = torch.empty((2160, 3840, 4), dtype=torch.float32, device="cuda:0")
    for adaptive, prepass in [(1, -1), (0, -1), (0, 0)]:
        ctx.set_features(adaptive_sampling=adaptive, use_prepass=prepass)
        for it in range(3):
            _, st = ctx.render(3840, 2160, out=out)
        print(f"4K kerr_schild adaptive={adaptive} prepass={prepass}: total {st.ms_total:.2f} ms, integrate {st.ms_integrate:.2f} ms, steps {st.trial_steps} "
              f"({st.trial_steps / st.ms_integrate / 1e6:.1f} G ray-steps/s), rays {st.rays_integrated} finished {st.rays_finished} extra {st.extra_rays}", flush=True)
        print("   ", ctx.launch_log(), flush=True)


if __name__ == "__main__":
    main()
