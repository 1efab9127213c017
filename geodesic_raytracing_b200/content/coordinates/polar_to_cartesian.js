// (t, r, theta, phi) -> (t, x, y, z)
(t, r, theta, phi) => {
    const planar = r * CMath.sin(theta);
    return [t, planar * CMath.cos(phi), planar * CMath.sin(phi), r * CMath.cos(theta)];
}
