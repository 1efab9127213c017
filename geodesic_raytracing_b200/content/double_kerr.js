// Two equal co-rotating Kerr black holes on the z axis held apart by a strut (the equal-mass double-Kerr solution of Manko,
// Rabadan & Ruiz) in Weyl-Papapetrou coordinates (t, rho, phi, z):
//   ds^2 = -f (dt - w dphi)^2 + ( e^{2 gamma} (drho^2 + dz^2) + rho^2 dphi^2 ) / f.
// Parameters: coordinate separation R, mass M and spin parameter a of each hole.
function two_equal_spinning_holes(t, rho, phi, z) {
    $cfg.R.$default = 3;
    $cfg.M.$default = 0.3;
    $cfg.a.$default = 0.27;
    const I = CMath.i;
    const sq = v => v * v;
    const R = $cfg.R, M = $cfg.M, a = $cfg.a;

    const span = R * R + 2 * M * R + 4 * a * a;
    const lift = R * R - 4 * M * M + 4 * a * a;
    const d = $pin(2 * M * a * lift / span);
    const sigma2 = $pin(M * M - a * a + (4 * M * M * a * a * lift) / sq(span));
    const sigma = CMath.sqrt(sigma2);
    const ia = I * a, id = I * d;

    // complex-weighted distances to the four rod ends: lower hole (+R/2 offset), upper hole (-R/2 offset)
    const lower = s => ((-M * (2 * s + R) + id) / (2 * M * M + (R + 2 * ia) * (s + ia))) * CMath.sqrt(rho * rho + sq(z + 0.5 * R + s));
    const upper = s => ((-M * (2 * s - R) + id) / (2 * M * M - (R - 2 * ia) * (s + ia))) * CMath.sqrt(rho * rho + sq(z - 0.5 * R + s));
    const Rp = $pin(lower(sigma)), Rn = $pin(lower(-sigma));
    const rp = $pin(upper(sigma)), rn = $pin(upper(-sigma));

    const K0 = $pin(4 * sigma2 * (sq(span) - 16 * M * M * a * a) / (M * M * (sq(R + 2 * M) + 4 * a * a)));
    const A = $pin(R * R * (Rp - Rn) * (rp - rn) - 4 * sigma2 * (Rp - rp) * (Rn - rn));
    const B = $pin(2 * R * sigma * ((R + 2 * sigma) * (Rn - rp) - (R - 2 * sigma) * (Rp - rn)));
    const G = -z * B + R * sigma * (2 * R * (Rn * rn - Rp * rp) + 4 * sigma * (Rp * Rn - rp * rn) - (R * R - 4 * sigma2) * (Rp - Rn - rp + rn));

    const AA = CMath.self_conjugate_multiply(A), BB = CMath.self_conjugate_multiply(B);
    const conj_sum = CMath.conjugate(A) + CMath.conjugate(B);
    const norm = CMath.Real((A + B) * conj_sum);
    const w = 4 * a - (2 * CMath.Imaginary(G * conj_sum) / (AA - BB));
    const f = (AA - BB) / norm;
    const inv_f = norm / (AA - BB);
    const e2gamma_over_f = norm / CMath.Real(K0 * K0 * Rp * Rn * rp * rn);

    const g = [];
    g.length = 16;
    g[0] = -f;
    g[2] = f * w * 2 * 0.5;
    g[8] = g[2];
    g[5] = e2gamma_over_f;
    g[10] = inv_f * rho * rho + w * w * -f;
    g[15] = e2gamma_over_f;
    return g;
}

two_equal_spinning_holes
