// Schwarzschild in ingoing Eddington-Finkelstein coordinates (v, r, theta, phi), regular across the horizon:
//   ds^2 = -(1 - rs/r) dv^2 + 2 dv dr + r^2 (dtheta^2 + sin^2(theta) dphi^2)
function ingoing_eddington_finkelstein(v, r, theta, phi) {
    $cfg.rs.$default = 1;
    const s = CMath.sin(theta);
    const g = new Array(16);
    g[0] = -(1 - $cfg.rs / r);
    g[1] = 1;
    g[4] = 1;
    g[10] = r * r;
    g[15] = r * r * s * s;
    return g;
}

ingoing_eddington_finkelstein
