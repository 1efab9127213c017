// Ellis / Morris-Thorne wormhole with throat radius n; the proper radial coordinate p runs over both universes
// (p < 0 is the far side): ds^2 = -dt^2 + dp^2 + (p^2 + n^2) (dtheta^2 + sin^2(theta) dphi^2)
(t, p, theta, phi) => {
    $cfg.n.$default = 1;
    const area = p * p + $cfg.n * $cfg.n;
    const s = CMath.sin(theta);
    return [-1, 1, area, area * (s * s)];
}
