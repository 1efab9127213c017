#!/usr/bin/env python3
"""BASELINE.json configs 1-3 through the reference's OWN kernels on host cores (test infrastructure, no GPU): /root/reference/cl.cl
compiled as C++ (oracle/ref_build.py, the -O3 x86-64-v3 unsafe-math build), whole frames in the reference's kernel order, with the
reference's own background (backgrounds/nasa.png through grt_image_load_png) when it is mounted.
usage: cpu_reference_configs.py [out_dir for PPM frames]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import geodesic_raytracing_b200 as grt  # noqa: E402
import oracle  # noqa: E402

CONFIGS = [
    ("config 1", "minkowski", 256, 256, dict(adaptive_sampling=0, redshift=0)),
    ("config 2", "schwarzschild", 1920, 1080, dict()),
    ("config 3", "kerr_schild", 3840, 2160, dict()),
]


def lin_to_srgb(x):
    x = np.clip(x, 0, 1)
    return np.where(x <= 0.0031308, x * 12.92, 1.055 * np.power(x, 1 / 2.4) - 0.055)


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else None
    nasa = "/root/reference/backgrounds/nasa.png"
    bg = grt.load_png(nasa) if os.path.exists(nasa) else grt.synthetic_background(4096, 2048)
    print(f"background: {'reference nasa.png' if os.path.exists(nasa) else 'synthetic'} {bg.shape[1]}x{bg.shape[0]}; host threads: {os.cpu_count()}")
    content = grt.Content(grt.CONTENT_DIR)
    print(f"{'config':9s} {'metric':14s} {'size':>10s} {'rays traced':>12s} {'ray-steps':>12s} {'integrate s':>11s} {'frame s':>8s} {'Mray-steps/s':>12s} {'per core':>9s} {'fps':>8s}")
    for tag, name, w, h, feat_kw in CONFIGS:
        m = content.load_by_name(name)
        ref = oracle.reference_for_metric(m, native=True)
        feat = ref.features(max_acceleration_change=m.info.max_acceleration_change, **feat_kw)
        t = time.time()
        r = ref.frame(w, h, feat, background=bg)
        wall = time.time() - t
        threads = ref.info["threads"]
        rate = r["trial_steps"] / r["t_integrate_s"] / 1e6
        print(f"{tag:9s} {name:14s} {w:5d}x{h:<4d} {r['rays_traced']:12d} {r['trial_steps']:12d} {r['t_integrate_s']:11.2f} {r['t_total_s']:8.2f} {rate:12.1f} {rate / threads:9.2f} {1.0 / r['t_total_s']:8.3f}",
              flush=True)
        if out_dir:
            os.makedirs(out_dir, exist_ok=True)
            rgb = (lin_to_srgb(r["image"][..., :3]) * 255 + 0.5).astype(np.uint8)
            with open(os.path.join(out_dir, f"{name}_{w}x{h}.ppm"), "wb") as f:
                f.write(f"P6\n{w} {h}\n255\n".encode() + rgb.tobytes())
        del r
        _ = wall


if __name__ == "__main__":
    main()
