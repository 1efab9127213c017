#!/usr/bin/env python3
"""One-off diagnostics (GPU box): where the CUDA frame and the oracle differ."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import geodesic_raytracing_b200 as grt  # noqa: E402
import oracle  # noqa: E402
from test_gpu_parity import bind, oracle_features  # noqa: E402
from test_reference_kernels import load_windows  # noqa: E402

content = grt.Content(grt.CONTENT_DIR)
ctx = grt.Context(0)


def confusion(a, b):
    return {f"{i}{j}": int(((a == i) & (b == j)).sum()) for i in range(3) for j in range(3) if ((a == i) & (b == j)).any()}


def windows(key):
    z, meta = load_windows()
    rec = meta[key]
    metric = content.load_by_name(rec["metric"])
    w, h = rec["size"]
    bg = grt.synthetic_background(*rec["background"])
    bind(ctx, metric, bg, **rec["features"])
    ctx.set_background(bg[::-1].copy(), which=1)
    img, st = ctx.render(w, h)
    rd = ctx.read_render_data(w, h).reshape(h, w)
    mask = ctx.read_prepass_mask()
    print(key, "steps", int(st.trial_steps), "extra", int(st.extra_rays), "mask dead", None if mask is None else int(mask.sum()))
    for i, (x0, y0, x1, y1) in enumerate(rec["windows"]):
        k = f"{key}|{i}|"
        term = z[k + "terminated"].astype(np.int32)
        got = rd[y0:y1, x0:x1]
        both = (got["terminated"] == 1) & (term == 1)
        duv = np.abs(got["tex_coord"] - z[k + "tex_coord"])
        duv = np.minimum(duv, 1 - duv).max(axis=-1)
        px = np.abs(img[y0:y1, x0:x1, :3] - z[k + "image"][..., :3]).max(axis=-1) <= 2.0 / 255.0
        print(f"  window {i} {(x0, y0, x1, y1)}: raw terminated same {(got['terminated'] == term).mean():.4f} finished-flag same {((got['terminated'] == 1) == (term == 1)).mean():.4f} "
              f"confusion(gpu,ref) {confusion(got['terminated'], term)} uv<1e-3 {(duv[both] < 1e-3).mean() if both.any() else 1:.4f} px {px.mean():.4f}")
        if mask is not None:
            print("    gpu prepass mask rows", (y0 // 16, (y1 + 15) // 16), "cols", (x0 // 16, (x1 + 15) // 16))
            print("    " + "\n    ".join("".join(str(v) for v in row) for row in mask[max(0, y0 // 16 - 1):(y1 + 15) // 16 + 1, max(0, x0 // 16 - 1):(x1 + 15) // 16 + 1]))


def frame_diff(name, w, h, adaptive):
    metric = content.load_by_name(name)
    orc = oracle.for_metric(metric)
    bg = grt.synthetic_background(1024, 512)
    base = bind(ctx, metric, bg, adaptive_sampling=adaptive)
    img, st = ctx.render(w, h)
    ref = orc.frame(w, h, oracle_features(orc, base), background=bg, use_prepass=base["use_prepass"], max_probes=base["max_probes"], want_rays=True)
    rd = ctx.read_render_data(w, h).reshape(h, w)
    o = ref["rdata"]
    bad = np.abs(img[..., :3] - ref["image"][..., :3]).max(axis=-1) > 2.0 / 255.0
    duv = np.abs(rd["tex_coord"] - o["tex_coord"])
    duv = np.minimum(duv, 1 - duv).max(axis=-1)
    ys, xs = np.where(bad)
    print(name, "adaptive", adaptive, "bad", int(bad.sum()), "parity of bad pixels (x%2,y%2):", {(a, b): int(((xs % 2 == a) & (ys % 2 == b)).sum()) for a in (0, 1) for b in (0, 1)})
    print("  confusion(gpu,ref)", confusion(rd["terminated"], o["terminated"]), "extra", int(st.extra_rays), ref["extra_rays"])
    print("  duv at bad pixels: median %.2e max %.2e ; duv overall max %.2e" % (np.median(duv[bad]) if bad.any() else 0, duv[bad].max() if bad.any() else 0, duv.max()))
    print("  colour diff at bad pixels: median %.4f max %.4f" % (np.median(np.abs(img[..., :3] - ref["image"][..., :3]).max(axis=-1)[bad]) if bad.any() else 0,
                                                                   np.abs(img[..., :3] - ref["image"][..., :3]).max(axis=-1).max()))
    for y, x in list(zip(ys, xs))[:12]:
        print(f"   ({x},{y}) gpu uv {rd['tex_coord'][y, x]} ref uv {o['tex_coord'][y, x]} term {rd['terminated'][y, x]}/{o['terminated'][y, x]} gpu rgb {img[y, x, :3]} ref rgb {ref['image'][y, x, :3]}"
              f" right uv gpu {rd['tex_coord'][y, min(x + 1, w - 1)]} ref {o['tex_coord'][y, min(x + 1, w - 1)]}")
    # rows of bad pixels
    print("  bad rows histogram (row: count):", {int(r): int(c) for r, c in zip(*np.unique(ys, return_counts=True))} if bad.sum() < 2000 else "many")
    if adaptive:
        # refined cells: a cell is refined iff its odd pixels carry traced (not interpolated) data; compare via extra ray pixel set
        rays = ctx.read_rays()
        first, extra = ctx.ray_count()
        ex = rays[first:first + extra]
        g = np.zeros((h, w), bool); g[ex["sy"], ex["sx"]] = True
        print("  refined cells gpu", int(g[1::2, 1::2].sum()), "oracle", ref["extra_rays"] // 3)


windows("config3")
frame_diff("schwarzschild_fast", 256, 144, 1)
frame_diff("schwarzschild", 256, 144, 1)
frame_diff("schwarzschild", 256, 144, 0)
