// Misner space extended to four dimensions, in the coordinates (T, phi, y, z) of arXiv:1102.0907 eq. (25):
//   ds^2 = -2 dT dphi - T dphi^2 + dy^2 + dz^2,
// flat space with phi identified periodically ($cfg.phi0, coordinates/misner_periodicity.js). T < 0 is the chronal region.
function misner_space(T, phi, y, z) {
    const g = [];
    g.length = 16;
    g[1] = -1;          // g_T_phi
    g[4] = -1;
    g[5] = -T;
    g[10] = 1;
    g[15] = 1;
    return g;
}

misner_space
