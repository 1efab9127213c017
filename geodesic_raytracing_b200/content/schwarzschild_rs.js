// Schwarzschild exterior with a configurable Schwarzschild radius, integrated with the adaptive step
function schwarzschild_exterior(t, r, theta, phi) {
    $cfg.rs.$default = 1;
    const f = 1 - $cfg.rs / r;
    const s = CMath.sin(theta);
    return [-f, 1 / f, r * r, r * r * s * s];
}

schwarzschild_exterior
