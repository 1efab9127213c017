// Two collinear Schwarzschild black holes held apart by a strut (Bach-Weyl / Israel-Khan), in Weyl
// cylindrical coordinates (t, rho, phi, z):
//   ds^2 = -e^{2U} dt^2 + e^{-2U} ( e^{2k} (drho^2 + dz^2) + rho^2 dphi^2 )
// For rods [z1 - m1, z1 + m1], [z2 - m2, z2 + m2] with R_i^+- = sqrt(rho^2 + (z - z_i -+ m_i)^2):
//   e^{2U} = prod_i (R_i^+ + R_i^- - 2 m_i) / (R_i^+ + R_i^- + 2 m_i)
//   e^{2k} = prod_i ((R_i^+ + R_i^-)^2 - 4 m_i^2) / (4 R_i^+ R_i^-) * (E(1+,2-) E(1-,2+) / (E(1+,2+) E(1-,2-)))
//   E(a, b) = R_a R_b + rho^2 + Z_a Z_b
function two_static_holes(t, rho, phi, z) {
    $cfg.m1.$default = 0.25;
    $cfg.m2.$default = 0.25;
    $cfg.separation.$default = 2;
    const m1 = $cfg.m1, m2 = $cfg.m2, half = $cfg.separation * 0.5;

    const rho2 = rho * rho;
    // rod end points (z offsets) and distances to them
    const Z = [z - (half + m1), z - (half - m1), z - (-half + m2), z - (-half - m2)];   // 1+, 1-, 2+, 2-
    const Rr = Z.map(zz => $pin(CMath.sqrt(rho2 + zz * zz)));

    const s1 = Rr[0] + Rr[1], s2 = Rr[2] + Rr[3];
    const e2U = $pin(((s1 - 2 * m1) / (s1 + 2 * m1)) * ((s2 - 2 * m2) / (s2 + 2 * m2)));

    const E = (a, b) => Rr[a] * Rr[b] + rho2 + Z[a] * Z[b];
    const self1 = (s1 * s1 - 4 * m1 * m1) / (4 * Rr[0] * Rr[1]);
    const self2 = (s2 * s2 - 4 * m2 * m2) / (4 * Rr[2] * Rr[3]);
    const cross = (E(0, 3) * E(1, 2)) / (E(0, 2) * E(1, 3));
    const e2k = $pin(self1 * self2 * cross);

    return [-e2U, e2k / e2U, rho2 / e2U, e2k / e2U];
}

two_static_holes
