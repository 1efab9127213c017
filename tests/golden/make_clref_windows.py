#!/usr/bin/env python3
"""Generate tests/golden/clref_windows.npz: windows of frames at the REAL SIZES of BASELINE.json configs 2-5, produced by the
REFERENCE'S OWN KERNELS (test infrastructure; run where /root/reference is mounted).

/root/reference/cl.cl compiled as C++ for the host (oracle/ref_build.py) runs the reference's frame sequence (main.cpp:2356-2526) in the
reference's default operating mode; `frame_window` integrates only the rays a window depends on (the 1/16 prepass runs in full) and is
checked to equal the full frame inside the window bit for bit (tests/test_reference_kernels.py), so a 3840x2160 comparison costs seconds
instead of minutes.  tests/test_gpu_parity.py compares the CUDA path's full-size frames against these windows with the bars of BASELINE.md.

Windows per config (128x64 pixels each): two across the shadow edge, one in the smooth sky, the top-left corner and the bottom-right
corner of the frame (border handling of adaptive sampling and of the texture Jacobian)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import geodesic_raytracing_b200 as grt  # noqa: E402
import oracle  # noqa: E402

BG_W, BG_H = 1024, 512
WW, WH = 128, 64
QUAT = (-0.70710678, 0.0, 0.0, 0.70710678)
# BASELINE.json configs[1..4]
CONFIGS = {
    "config2": dict(metric="schwarzschild", size=(1920, 1080), features={}),
    "config3": dict(metric="kerr_schild", size=(3840, 2160), features={}),
    "config4": dict(metric="alcubierre", size=(1920, 1080), features={"redshift": 1}, geodesic=dict(speed=0.3, tau=2.0)),
    "config5": dict(metric="double_unequal_kerr", size=(3840, 2160), features={}),
}


def forward(q):
    q = np.asarray(q, np.float64)
    return np.array([2 * (q[0] * q[2] + q[3] * q[1]), 2 * (q[1] * q[2] - q[3] * q[0]), 1 - 2 * (q[0] ** 2 + q[1] ** 2)])


def pick_windows(ref, w, h, feat, kw):
    """Window rectangles (even origins, inside the frame) chosen on a 1/16-size frame of the reference kernels: two places on the
    shadow edge (the one nearest the frame centre and the one farthest from it), the finished place farthest from any unfinished
    pixel, and two frame corners."""
    small = ref.frame(w // 16, h // 16, ref.features(**{**feat, "adaptive_sampling": 0}), use_prepass=0, **kw)
    fin = small["rdata"]["terminated"] == 1
    sh, sw = fin.shape
    ys, xs = np.mgrid[0:sh, 0:sw]
    pad = np.pad(fin, 1, mode="edge")
    mixed = np.zeros_like(fin)
    for dy in (0, 1, 2):
        for dx in (0, 1, 2):
            mixed |= pad[dy:dy + sh, dx:dx + sw] != fin
    picks = []
    if mixed.any():
        ey, ex = ys[mixed], xs[mixed]
        d = (ey - sh / 2) ** 2 + (ex - sw / 2) ** 2
        k = int(np.argmin(d))
        picks.append((ex[k], ey[k]))
        k2 = int(np.argmax((ey - ey[k]) ** 2 + (ex - ex[k]) ** 2))
        picks.append((ex[k2], ey[k2]))
    else:
        picks.append((sw // 2, sh // 2))                     # no shadow in this frame: the frame centre instead
    if fin.any() and (~fin).any():
        uy, ux = ys[~fin].ravel(), xs[~fin].ravel()
        fy, fx = ys[fin].ravel(), xs[fin].ravel()
        dist = np.array([((uy - y) ** 2 + (ux - x) ** 2).min() for y, x in zip(fy, fx)])
        k = int(np.argmax(dist))
        picks.append((fx[k], fy[k]))
    else:
        picks.append((sw // 4, sh // 4))
    wins = []
    for (px, py) in picks:
        x0 = int(np.clip(px * 16 + 8 - WW // 2, 0, w - WW)) & ~1
        y0 = int(np.clip(py * 16 + 8 - WH // 2, 0, h - WH)) & ~1
        wins.append((x0, y0, x0 + WW, y0 + WH))
    wins.append((0, 0, WW, WH))
    wins.append((w - WW, h - WH, w, h))
    return wins


def main():
    content = grt.Content(grt.CONTENT_DIR)
    bg = grt.synthetic_background(BG_W, BG_H)
    out, meta = {}, {}
    for key, cfg in CONFIGS.items():
        m = content.load_by_name(cfg["metric"])
        ref = oracle.reference_for_metric(m)
        w, h = cfg["size"]
        feat = dict(max_acceleration_change=m.info.max_acceleration_change, **cfg["features"])
        kw = {}
        rec = {"metric": cfg["metric"], "size": [w, h], "features": cfg["features"], "background": [BG_W, BG_H]}
        if "geodesic" in cfg:
            # camera attached to a geodesic (main.cpp:2675-2760, 2264-2296): snapshot from the default camera, interpolate at tau
            g = cfg["geodesic"]
            small = ref.frame(16, 16, ref.features(**{**feat, "adaptive_sampling": 0}))
            speed = forward(QUAT) * g["speed"]
            pos, vel, ds, tr, _ = ref.geodesic_snapshot(small["camera_generic"], small["tetrad"], speed, ref.features(**feat))
            cam, tet, _ = ref.geodesic_interpolate(pos, vel, ds, tr, g["tau"], 1, speed)
            kw["given_camera"] = (cam, tet)
            out[f"{key}|camera"], out[f"{key}|tetrad"] = cam, tet
            rec["geodesic"] = g
        wins = pick_windows(ref, w, h, feat, kw)
        rec["windows"] = wins
        for i, win in enumerate(wins):
            r = ref.frame_window(w, h, win, ref.features(**feat), background=bg, background2=bg[::-1].copy(), **kw)
            rd = r["rdata"]
            out[f"{key}|{i}|image"] = r["image"]
            out[f"{key}|{i}|tex_coord"] = np.ascontiguousarray(rd["tex_coord"])
            out[f"{key}|{i}|z_shift"] = np.ascontiguousarray(rd["z_shift"])
            out[f"{key}|{i}|terminated"] = np.ascontiguousarray(rd["terminated"]).astype(np.int8)
            out[f"{key}|{i}|side"] = np.ascontiguousarray(rd["side"]).astype(np.int8)
            print(key, win, "finished", float((rd["terminated"] == 1).mean()), "steps", r["trial_steps"], f"{r['t_total_s']:.1f}s", flush=True)
        meta[key] = rec
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "clref_windows.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
