"""Pins for the CPU oracle. The reference ships no golden vectors for this path, so the
restatement is anchored on analytic identities: flat-space rays, the Schwarzschild shadow, null-norm and
Killing-energy conservation, partials vs finite differences, and the three acceleration formulations."""
import numpy as np
import pytest

CAM_POS = (0.0, 0.0, -4.0, 0.0)
CAM_QUAT = (-0.70710678, 0.0, 0.0, 0.70710678)   # main.cpp:669-673


def rot(q, v):
    q = np.asarray(q, np.float64) / np.linalg.norm(q)
    qv, w = q[:3], q[3]
    t = 2 * np.cross(qv, v)
    return v + w * t + np.cross(qv, t)


def test_minkowski_rays_are_straight_and_hit_the_analytic_direction(oracles):
    m, o = oracles("minkowski")
    w = h = 64
    f = o.features(adaptive_sampling=0, max_acceleration_change=m.info.max_acceleration_change)
    r = o.frame(w, h, f, want_rays=True)
    rays = r["rays"]
    assert (rays["terminated"] == 1).all()
    assert np.abs(rays["acceleration"]).max() == 0.0
    assert r["trial_steps"] == r["accepted_steps"]
    rd = r["rdata"]
    # hit direction on the universe sphere for a straight line from the camera, analytically
    for (sx, sy) in [(0, 0), (13, 40), (32, 32), (63, 5), (20, 63)]:
        d = np.array([sx - w / 2, sy - h / 2, (w / 2) / np.tan(np.pi / 4)])
        d = rot(CAM_QUAT, d / np.linalg.norm(d))
        p0 = np.array(CAM_POS[1:])
        b = 2 * d.dot(p0)
        c = p0.dot(p0) - 20.0 ** 2
        t = (-b + np.sqrt(b * b - 4 * c)) / 2
        hit = p0 + t * d
        theta = np.arccos(hit[2] / 20.0)
        phi = np.arctan2(hit[1], hit[0])
        u = np.fmod(phi, 2 * np.pi) / (2 * np.pi) + 0.5
        v = theta / np.pi
        got = rd[sy, sx]["tex_coord"]
        assert abs(got[0] - u) < 2e-4 and abs(got[1] - v) < 2e-4, (sx, sy, got, u, v)
    assert np.abs(rd["z_shift"]).max() < 1e-5   # static observer, static emitter, flat space


def test_schwarzschild_shadow_radius(oracles):
    """Critical impact parameter b_c = 3 sqrt(3)/2 r_s: from r0 = 4 the shadow's angular radius alpha satisfies
    sin(alpha) = b_c sqrt(1 - r_s/r0) / r0 for a static observer."""
    m, o = oracles("schwarzschild_fast")
    w, h = 256, 256
    f = o.features(adaptive_sampling=0, max_acceleration_change=m.info.max_acceleration_change)
    r = o.frame(w, h, f)
    term = r["rdata"]["terminated"]
    u = r["rdata"]["tex_coord"]
    # captured rays end at the singular terminator: they are "finished" too (singular metrics); separate them by hit radius
    rays = o.frame(w, h, f, want_rays=True)["rays"].reshape(h, w)
    captured = np.abs(rays["position"][..., 1]) < 2.0
    assert term.min() >= 0 and captured.any() and (~captured).any()
    bc = 1.5 * np.sqrt(3.0)
    alpha = np.arcsin(bc * np.sqrt(1 - 1 / 4.0) / 4.0)
    fstop = (w / 2) / np.tan(np.pi / 4)
    expect_px = fstop * np.tan(alpha)
    row = captured[h // 2]
    xs = np.where(row)[0]
    radius_px = (xs.max() - xs.min() + 1) / 2
    assert abs(radius_px - expect_px) < 2.0, (radius_px, expect_px)
    col = captured[:, w // 2]
    ys = np.where(col)[0]
    assert abs((ys.max() - ys.min() + 1) / 2 - expect_px) < 2.0
    assert u.shape == (h, w, 2)


def lower(g, v):
    return g.reshape(4, 4).astype(np.float64) @ np.asarray(v, np.float64)


def test_kerr_schild_null_norm_and_energy_are_conserved(oracles):
    m, o = oracles("kerr_schild")
    f = o.features(max_acceleration_change=m.info.max_acceleration_change)
    cam, tet = o.camera(CAM_POS)
    # a ray that passes the hole at a safe distance: tetrad direction mostly sideways
    d = np.array([0.55, 0.1, 0.8])
    d /= np.linalg.norm(d)
    k = d[0] * tet[1] + d[1] * tet[2] + d[2] * tet[3] - tet[0]
    g0, _ = o.metric_eval(cam)
    assert abs(lower(g0, k).dot(k)) < 1e-5                      # null at emission
    P, V, term = o.trace_path(cam, k, f)
    assert term == 1 and len(P) > 50
    E0 = lower(g0, k)[0]
    norms, energies = [], []
    for p, v in zip(P[::10], V[::10]):
        g, _ = o.metric_eval(p)
        kl = lower(g, v)
        norms.append(kl.dot(v) / max(1.0, np.abs(v).max() ** 2))
        energies.append(kl[0])
    assert np.abs(norms).max() < 5e-4
    assert np.abs(np.array(energies) / E0 - 1).max() < 5e-3   # d/dt is Killing: k_t is conserved


@pytest.mark.parametrize("name, pos", [
    ("schwarzschild_fast", [0.3, 3.1, 1.1, 0.4]), ("kerr_schild", [0.2, 2.5, -1.4, 0.8]), ("alcubierre", [0.4, 1.2, 0.7, -0.5]),
    ("double_schwarzschild", [0.1, 1.7, 0.6, 0.9]),
])
def test_partials_match_finite_differences(oracles, name, pos):
    """The reference's own (disabled) cross-check: NUMERICAL_DIFFERENTIATION, cl.cl:1055-1088."""
    _, o = oracles(name)
    pos = np.array(pos, np.float64)
    _, dg = o.metric_eval(pos)
    for k in range(4):
        h = 2e-3
        a, b = pos.copy(), pos.copy()
        a[k] -= h
        b[k] += h
        fd = (o.metric_eval(b)[0].astype(np.float64) - o.metric_eval(a)[0]) / (2 * h)
        assert np.allclose(dg[k * 16:(k + 1) * 16], fd, rtol=2e-2, atol=3e-3), (name, k)


@pytest.mark.parametrize("name, pos", [
    ("schwarzschild_fast", [0.3, 3.1, 1.1, 0.4]), ("kerr_schild", [0.2, 2.5, -1.4, 0.8]), ("alcubierre", [0.4, 1.2, 0.7, -0.5]),
    ("double_schwarzschild", [0.1, 1.7, 0.6, 0.9]),
])
def test_three_acceleration_formulations_agree(oracles, name, pos):
    """GEO_ACCEL (symbolic Christoffel, metric.hpp:184-244) == numeric Christoffel (cl.cl:1443-1537) == the lowered form the CUDA kernels use."""
    m, o = oracles(name)
    vel = np.array([-1.0, 0.3, -0.2, 0.5])
    a = o.accel(pos, vel)
    if m.info.is_constant_theta:
        # the emitted accelerations are specialised to the theta = pi/2 plane (cl.cl:3294-3306)
        pos = list(pos)
        pos[2] = np.pi / 2
        vel[2] = 0
        a = o.accel(pos, vel)
        a["numeric"][2] = 0
    scale = np.abs(a["numeric"]).max() + 1e-6
    assert np.abs(a["ref"] - a["numeric"]).max() / scale < 2e-4
    assert np.abs(a["lowered"] - a["numeric"]).max() / scale < 2e-4


def test_coordinate_maps_round_trip(oracles):
    for name in ("kerr_schild", "double_schwarzschild", "schwarzschild_fast"):
        _, o = oracles(name)
        pos = np.array([0.5, 1.5, 0.9, 0.7], np.float32)
        vel = np.array([1.0, 0.2, -0.1, 0.3], np.float32)
        c = o.coords(pos, vel)
        back = o.coords(c["to_polar"], c["to_polar_vel"])
        assert np.allclose(back["from_polar"], pos, atol=1e-5), name
        assert np.allclose(back["from_polar_vel"], vel, atol=1e-5), name


def test_camera_tetrad_is_orthonormal(oracles):
    # schwarzschild_skewed keeps time in coordinate slot 1: the frame construction has to find and swap the timelike leg (cl.cl:1761-1861)
    for name in ("minkowski", "schwarzschild_fast", "schwarzschild_skewed", "kerr_schild", "kerr_boyer", "alcubierre", "double_schwarzschild"):
        _, o = oracles(name)
        cam, tet = o.camera(CAM_POS)
        g, _ = o.metric_eval(cam)
        G = g.reshape(4, 4).astype(np.float64)
        gram = tet.astype(np.float64) @ G @ tet.astype(np.float64).T
        assert np.allclose(gram, np.diag([-1, 1, 1, 1]), atol=2e-4), (name, gram)


def test_mip_pyramid_shape_and_requantisation(oracles, grt):
    _, o = oracles("minkowski")
    bg = grt.synthetic_background(256, 128)
    l0, levels = o.mip_level(bg, 0)
    assert levels == 8 and l0.shape == (128, 256, 4)      # floor(log2(128)) + 1, capped at 10 (graphics_settings.cpp:159-163)
    # (uchar)((b / 255.f) * 255) truncates: never above the input, at most one below
    d = bg.astype(np.int32) - l0.astype(np.int32)
    assert d.min() >= 0 and d.max() <= 1
    l3, _ = o.mip_level(bg, 3)
    assert l3.shape == (16, 32, 4)


def test_prepass_only_culls_rays_that_would_be_black(oracles):
    m, o = oracles("kerr_schild")
    f = o.features(adaptive_sampling=0, max_acceleration_change=m.info.max_acceleration_change)
    w, h = 192, 128
    a = o.frame(w, h, f, use_prepass=0)
    b = o.frame(w, h, f, use_prepass=1)
    culled = b["rdata"]["terminated"] == 2
    assert culled.any()
    # culled pixels are overwhelmingly ones the full trace leaves unfinished
    assert (a["rdata"]["terminated"][culled] == 0).mean() > 0.97
    assert b["trial_steps"] < a["trial_steps"] * 0.7
    same = (a["rdata"]["terminated"] == 1) & (b["rdata"]["terminated"] == 1)
    assert np.array_equal(a["rdata"]["tex_coord"][same], b["rdata"]["tex_coord"][same])
