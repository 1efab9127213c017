// Kerr black hole in Boyer-Lindquist coordinates (t, r, theta, phi), spin a, Schwarzschild radius rs:
//   Sigma = r^2 + a^2 cos^2(theta),  Delta = r^2 - rs r + a^2
//   g_tt = -(1 - rs r / Sigma), g_rr = Sigma / Delta, g_thth = Sigma,
//   g_phph = (r^2 + a^2 + rs r a^2 sin^2(theta) / Sigma) sin^2(theta),  g_tph = -rs r a sin^2(theta) / Sigma
function kerr_bl(t, r, theta, phi) {
    $cfg.rs.$default = 1;
    $cfg.a.$default = -0.5;
    const rs = $cfg.rs, a = $cfg.a;
    const ct = CMath.cos(theta), s2 = CMath.sin(theta) * CMath.sin(theta);
    const Sigma = r * r + a * a * ct * ct;
    const Delta = r * r - rs * r + a * a;
    const frame_drag = -rs * r * a * s2 / Sigma;

    const g = new Array(16);
    g[0] = -(1 - rs * r / Sigma);
    g[5] = Sigma / Delta;
    g[10] = Sigma;
    g[15] = (r * r + a * a + (rs * r * a * a / Sigma) * s2) * s2;
    g[3] = frame_drag;
    g[12] = frame_drag;
    return g;
}

kerr_bl
