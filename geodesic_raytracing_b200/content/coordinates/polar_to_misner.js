// polar (t, r, theta, phi) -> cartesian -> Misner: phi = -2 log((x - t) / 2), T = (x^2 - t^2) / 4 (arXiv:1102.0907 eqs. 8, 9)
(t, r, theta, phi) => {
    const st = CMath.sin(theta);
    const x = r * st * CMath.cos(phi), y = r * st * CMath.sin(phi), z = r * CMath.cos(theta);
    return [(x * x - t * t) / 4, -2 * CMath.log((x - t) / 2), y, z];
}
