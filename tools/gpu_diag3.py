#!/usr/bin/env python3
"""Divergence of the GPU geodesic-camera path from the oracle's along the whole path (to set tolerances that grow with arc length)."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import geodesic_raytracing_b200 as grt  # noqa: E402
import oracle  # noqa: E402
from test_gpu_parity import bind, oracle_features  # noqa: E402

content = grt.Content(grt.CONTENT_DIR)
ctx = grt.Context(0)
bg = grt.synthetic_background(256, 128)
q = np.array([-0.70710678, 0, 0, 0.70710678])
fwd = np.array([2 * (q[0] * q[2] + q[3] * q[1]), 2 * (q[1] * q[2] - q[3] * q[0]), 1 - 2 * (q[0] ** 2 + q[1] ** 2)])
for name, speed, cam in [("alcubierre", 0.2, None), ("schwarzschild", 0.0, (0.0, 0.5, -6.0, 1.0)), ("schwarzschild", 0.35, (0.0, 0.5, -6.0, 1.0)), ("kerr_boyer", 0.2, (0.0, 0.5, -6.0, 1.0)),
                         ("kerr_schild", -0.3, (0.0, 0.5, -6.0, 1.0))]:
    metric = content.load_by_name(name)
    orc = oracle.for_metric(metric)
    base = bind(ctx, metric, bg, adaptive_sampling=0)
    if cam:
        ctx.set_camera(cam)
    cs = ctx.camera_state()
    n = ctx.snapshot_geodesic(speed, True)
    pos, vel, ds, tr = ctx.read_geodesic(n)
    opos, ovel, ods, otr, _ = orc.geodesic_snapshot(np.array(list(cs.pos_generic)), np.array(list(cs.tetrad)).reshape(4, 4), fwd * speed, oracle_features(orc, base))
    m = min(n, len(opos))
    s = np.cumsum(ds[:m]); so = np.cumsum(ods[:m])
    print(f"{name} speed {speed}: n gpu {n} oracle {len(opos)}, total proper time gpu {s[-1]:.3f} oracle {so[-1]:.3f}")
    for frac in (0.1, 0.25, 0.5, 0.75, 0.9, 1.0):
        k = max(1, int(m * frac))
        scale = np.abs(opos[:k]).max(axis=0) + 1e-3
        dpos = (np.abs(pos[:k] - opos[:k]) / scale).max()
        dvel = (np.abs(vel[:k] - ovel[:k]) / (np.abs(ovel[:k]).max(axis=0) + 1e-3)).max()
        dtr = np.abs(tr[:, :k] - otr[:, :k]).max()
        dds = (np.abs(ds[:k] - ods[:k]) / (np.abs(ods[:k]) + 1e-9)).max()
        print(f"   first {frac:4.2f} of the path ({k} points): rel dpos {dpos:.2e} rel dvel {dvel:.2e} abs dtetrad {dtr:.2e} rel dds {dds:.2e}")
    # interpolated camera at three proper times
    for frac in (0.2, 0.6, 0.95):
        tau = float(s[-1] * frac)
        ctx.set_camera_on_geodesic(tau)
        cs2 = ctx.camera_state()
        ocam, otet, _ = orc.geodesic_interpolate(opos, ovel, ods, otr, tau, True, fwd * speed)
        print(f"   camera at tau = {frac:.2f} of the path: abs dpos {np.abs(np.array(list(cs2.pos_generic)) - ocam).max():.2e} abs dtetrad {np.abs(np.array(list(cs2.tetrad)).reshape(4, 4) - otet).max():.2e}")
    ctx.clear_camera_geodesic()
    ctx.set_camera()
