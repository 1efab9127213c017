// Alcubierre warp bubble moving along x with velocity v: ds^2 = -dt^2 + (dx - v f dt)^2 + dy^2 + dz^2,
// f(rs) = (tanh(sigma (rs + R)) - tanh(sigma (rs - R))) / (2 tanh(sigma R)), rs = |(x - v t, y, z)|.
function warp_bubble(t, x, y, z) {
    $cfg.velocity.$default = 2;
    $cfg.sigma.$default = 1;
    $cfg.R.$default = 2;
    const v = $cfg.velocity, sigma = $cfg.sigma, R = $cfg.R;

    const rs = CMath.fast_length(x - v * t, y, z);
    const f = (CMath.tanh(sigma * (rs + R)) - CMath.tanh(sigma * (rs - R))) / (2 * CMath.tanh(sigma * R));

    const g = new Array(16);
    g[0] = v * v * f * f - 1;
    g[1] = g[4] = -2 * v * f * 0.5;
    g[5] = 1;
    g[10] = 1;
    g[15] = 1;
    return g;
}

warp_bubble
