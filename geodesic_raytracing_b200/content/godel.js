// Goedel's rotating dust universe in cylindrical coordinates (t, r, phi, z) with length scale a:
//   ds^2 = -dt^2 + dr^2 / (1 + q^2) + r^2 (1 - q^2) dphi^2 + dz^2 - (2 r^2 / (sqrt(2) a)) dt dphi,   q = r / 2a.
// Closed timelike curves start at q = 1; the axis r = 0 is a coordinate singularity (has_cylindrical_singularity).
function goedel_universe(t, r, phi, z) {
    $cfg.a.$default = 1000;
    const scale = $cfg.a;
    const q = r / (2 * scale), q2 = q * q;
    const drag = -(r * r) * 2 * (1 / (CMath.sqrt(2) * scale));      // coefficient of dt dphi
    const g = [];
    g.length = 16;
    g[0] = -1;
    g[5] = 1 / (1 + q2);
    g[10] = r * r * (1 - q2);
    g[15] = 1;
    g[2] = 0.5 * drag;
    g[8] = 0.5 * drag;
    return g;
}

goedel_universe
