#!/usr/bin/env python3
"""Catalogue conformance on the CPU (test infrastructure): for every metric script of a content folder, run one small frame
through (i) the reference's own kernels - /root/reference/cl.cl compiled as C++ (oracle/ref_build.py) - and (ii) the oracle
restatement, and report whether render_data, image and step counts are bit-identical.
usage: catalogue_reference_conformance.py <content folder> [width height]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import geodesic_raytracing_b200 as grt  # noqa: E402
import oracle  # noqa: E402


def main():
    folder = sys.argv[1]
    w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (48, 32)
    content = grt.Content(folder)
    bg = grt.synthetic_background(256, 128)
    print(f"{'metric':36s} {'big':>3s} {'cth':>3s} {'steps(ref kernels)':>18s} {'steps(oracle)':>14s} {'finished':>8s} {'rdata':>6s} {'image':>6s} {'s':>6s}")
    bad = 0
    for name in content.metric_names():
        t = time.time()
        m = content.load_by_name(name)
        ref, orc = oracle.reference_for_metric(m), oracle.for_metric(m)
        feat = orc.features(max_acceleration_change=m.info.max_acceleration_change, redshift=1)
        a = ref.frame(w, h, feat, background=bg, background2=bg[::-1].copy())
        b = orc.frame(w, h, feat, background=bg, background2=bg[::-1].copy())
        same_rd = all(np.array_equal(a["rdata"][k], b["rdata"][k], equal_nan=True) for k in ("tex_coord", "z_shift", "sx", "sy", "terminated", "side"))
        same_img = np.array_equal(a["image"], b["image"], equal_nan=True)
        same = same_rd and same_img and a["trial_steps"] == b["trial_steps"]
        bad += 0 if same else 1
        print(f"{name:36s} {m.info.is_big:3d} {m.info.is_constant_theta:3d} {a['trial_steps']:18d} {b['trial_steps']:14d} {float((a['rdata']['terminated'] == 1).mean()):8.3f} "
              f"{'same' if same_rd else 'DIFF':>6s} {'same' if same_img else 'DIFF':>6s} {time.time() - t:6.1f}", flush=True)
    print(f"{bad} metric(s) differ")


if __name__ == "__main__":
    main()
