#!/usr/bin/env python3
"""One-off diagnostics (GPU box): initial rays and the equatorial row of the constant-theta metrics."""
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
np.set_printoptions(linewidth=250, precision=7, suppress=False)


def rays_by_pixel(rays, w, h):
    out = np.zeros((h, w), rays.dtype)
    out[rays["sy"], rays["sx"]] = rays
    return out


def initial(name, w, h):
    metric = content.load_by_name(name)
    orc = oracle.for_metric(metric)
    bg = grt.synthetic_background(256, 128)
    base = bind(ctx, metric, bg, adaptive_sampling=0, universe_size=0.5, use_prepass=0)      # every ray "finishes" at once: final state = initial state
    ctx.render(w, h)
    g = rays_by_pixel(ctx.read_rays(), w, h)
    ref = orc.frame(w, h, oracle_features(orc, base), background=bg, use_prepass=0, want_rays=True)
    o = rays_by_pixel(ref["rays"], w, h)
    for f in ("position", "velocity", "initial_quat"):
        d = np.abs(g[f].astype(np.float64) - o[f].astype(np.float64)).max(axis=-1)
        scale = np.abs(o[f]).max(axis=-1) + 1e-6
        rel = d / scale
        print(name, f, "max rel diff %.3e at %s ; rows with rel > 1e-4: %s" % (np.nanmax(rel), np.unravel_index(np.nanargmax(rel), rel.shape), sorted(set(np.where(rel > 1e-4)[0].tolist()))[:20]),
              "nan gpu", int(np.isnan(g[f]).sum()), "nan ref", int(np.isnan(o[f]).sum()))
    print(" term gpu", np.unique(g["terminated"], return_counts=True), "ref", np.unique(o["terminated"], return_counts=True))
    r = h // 2
    for x in (0, 1, w // 2, w - 1):
        print(f"  row {r} x {x}: gpu pos {g['position'][r, x]} vel {g['velocity'][r, x]} quat {g['initial_quat'][r, x]}\n              ref pos {o['position'][r, x]} vel {o['velocity'][r, x]} quat {o['initial_quat'][r, x]}")


def equator(name, w, h):
    metric = content.load_by_name(name)
    orc = oracle.for_metric(metric)
    bg = grt.synthetic_background(1024, 512)
    base = bind(ctx, metric, bg, adaptive_sampling=0)
    img, st = ctx.render(w, h)
    g = rays_by_pixel(ctx.read_rays(), w, h)
    ref = orc.frame(w, h, oracle_features(orc, base), background=bg, use_prepass=base["use_prepass"], want_rays=True)
    o = rays_by_pixel(ref["rays"], w, h)
    mism = g["terminated"] != o["terminated"]
    ys, xs = np.where(mism)
    print(name, "terminated mismatches", int(mism.sum()), "rows", {int(r): int(c) for r, c in zip(*np.unique(ys, return_counts=True))})
    for y, x in list(zip(ys, xs))[:6]:
        print(f"  ({x},{y}) gpu term {g['terminated'][y, x]} pos {g['position'][y, x]} vel {g['velocity'][y, x]} acc {g['acceleration'][y, x]}\n           ref term {o['terminated'][y, x]} pos {o['position'][y, x]} vel {o['velocity'][y, x]}")
    for r in (h // 2 - 2, h // 2 - 1, h // 2, h // 2 + 1):
        print(f"  row {r}: finished gpu {int((g['terminated'][r] == 1).sum())} ref {int((o['terminated'][r] == 1).sum())}")


def midrow_bad(name, w, h):
    metric = content.load_by_name(name)
    orc = oracle.for_metric(metric)
    bg = grt.synthetic_background(1024, 512)
    base = bind(ctx, metric, bg, adaptive_sampling=1)
    img, st = ctx.render(w, h)
    rd = ctx.read_render_data(w, h).reshape(h, w)
    ref = orc.frame(w, h, oracle_features(orc, base), background=bg, use_prepass=base["use_prepass"], max_probes=base["max_probes"])
    o = ref["rdata"]
    bad = np.abs(img[..., :3] - ref["image"][..., :3]).max(axis=-1) > 2.0 / 255.0
    ys, xs = np.where(bad & (np.arange(h)[:, None] > 30) & (np.arange(h)[:, None] < 120))
    rays = ctx.read_rays()
    first, extra = ctx.ray_count()
    ex = rays[first:first + extra]
    refined = np.zeros((h, w), bool)
    refined[ex["sy"], ex["sx"]] = True
    print(name, "mid-row bad pixels", len(ys), "of which refined (traced) pixels", int(refined[ys, xs].sum()))
    for y, x in list(zip(ys, xs))[:10]:
        print(f"  ({x},{y}) refined {refined[y, x]}")
        for (yy, xx) in ((y, x), (y, x + 1), (y + 1, x)):
            print(f"      ({xx},{yy}) gpu uv {rd['tex_coord'][yy, xx]} z {rd['z_shift'][yy, xx]:.6f} t {rd['terminated'][yy, xx]} s {rd['side'][yy, xx]} | ref uv {o['tex_coord'][yy, xx]} z {o['z_shift'][yy, xx]:.6f} t {o['terminated'][yy, xx]} s {o['side'][yy, xx]}")
        print(f"      gpu rgb {img[y, x, :3]} ref rgb {ref['image'][y, x, :3]}")


initial("schwarzschild", 256, 144)
initial("kerr_schild", 256, 144)
equator("schwarzschild", 256, 144)
midrow_bad("schwarzschild_fast", 256, 144)
