#!/usr/bin/env python3
"""Parity figures of every bundled metric and sampling mode against the CPU oracle (GPU box): the measured values the
bars in tests/test_gpu_parity.py are set from.  python tools/gpu_parity_report.py [W H]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import geodesic_raytracing_b200 as grt  # noqa: E402
import oracle  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from test_gpu_parity import METRICS, bind, camera_of, oracle_features  # noqa: E402


def main():
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 144)
    content = grt.Content(grt.CONTENT_DIR)
    ctx = grt.Context(0)
    bg = grt.synthetic_background(1024, 512)
    for name in METRICS:
        metric = content.load_by_name(name)
        orc = oracle.for_metric(metric)
        for adaptive in (0, 1):
            base = bind(ctx, metric, bg, adaptive_sampling=adaptive)
            img, st = ctx.render(w, h)
            ref = orc.frame(w, h, oracle_features(orc, base), camera_pos=camera_of(metric), background=bg, use_prepass=base["use_prepass"], max_probes=base["max_probes"])
            rd = ctx.read_render_data(w, h).reshape(h, w)
            o = ref["rdata"]
            agree = (rd["terminated"] == o["terminated"]).mean()
            both = (rd["terminated"] == 1) & (o["terminated"] == 1)
            duv = np.abs(rd["tex_coord"] - o["tex_coord"])
            duv = np.minimum(duv, 1 - duv)[both].max(axis=-1)
            bad = np.abs(img[..., :3] - ref["image"][..., :3]).max(axis=-1) > 2.0 / 255.0
            # where are the differing pixels: on refined/interpolated cells? near unfinished pixels?
            near_edge = np.zeros_like(bad)
            unf = o["terminated"] != 1
            for dy in (-2, -1, 0, 1, 2):
                for dx in (-2, -1, 0, 1, 2):
                    near_edge |= np.roll(np.roll(unf, dy, 0), dx, 1)
            print(f"{name:28s} adaptive {adaptive}: terminated {agree:.5f} uv<1e-3 {(duv < 1e-3).mean() if duv.size else 1:.5f} px {1 - bad.mean():.5f} "
                  f"(bad {int(bad.sum())}, of which within 2 px of an unfinished pixel {int((bad & near_edge).sum())}, uv>1e-3 there {int(((np.minimum(np.abs(rd['tex_coord'] - o['tex_coord']), 1 - np.abs(rd['tex_coord'] - o['tex_coord'])).max(axis=-1) > 1e-3) & both & bad).sum())}) "
                  f"steps {int(st.trial_steps)} vs {ref['trial_steps']} extra {int(st.extra_rays)} vs {ref['extra_rays']}", flush=True)


if __name__ == "__main__":
    main()
