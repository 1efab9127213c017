// Two unequal spinning black holes on the z axis, kept apart by a strut: the five-parameter
// Manko-Ruiz solution (Phys. Lett. B 794 (2019) 36) in Weyl-Papapetrou coordinates (t, rho, phi, z),
//   ds^2 = -f (dt - w dphi)^2 + ( e^{2 gamma} (drho^2 + dz^2) + rho^2 dphi^2 ) / f.
// Parameters: masses m1, m2, spins per unit mass fa1, fa2 (a_i = fa_i m_i), coordinate separation R.
// Everything up to the rod half-lengths sigma_i depends on the parameters only; the field point enters
// through the four distances to the rod ends.
function two_spinning_holes(t, rho, phi, z) {
    $cfg.m1.$default = 0.15;
    $cfg.m2.$default = 0.3;
    $cfg.fa1.$default = 1;
    $cfg.fa2.$default = -0.3;
    $cfg.R.$default = 4;

    const I = CMath.i;
    const sq = v => v * v;
    const m1 = $cfg.m1, m2 = $cfg.m2, R = $cfg.R;
    const a1 = $cfg.fa1 * m1, a2 = $cfg.fa2 * m2;
    const M = m1 + m2, J = m1 * a1 + m2 * a2;

    // a: real root of (a1 + a2 - a)(R^2 - M^2 + a^2) / (2 (R + M)) - M a + J = 0 (Cardano)
    const k = a1 + a2, B = sq(R) - sq(M), C = 2 * (R + M);
    const P = 18 * B * k + 27 * C * J - 9 * C * k * M + 2 * k * k * k;
    const Q = 3 * B + 3 * C * M - sq(k);
    const cube = CMath.pow(CMath.sqrt(sq(P) + 4 * Q * Q * Q) + P, 1 / 3);
    const two13 = CMath.pow(2, 1 / 3);
    const a = $pin(cube / (3 * two13) - (two13 * Q) / (3 * cube) + k / 3);

    const RM = R + M;
    const S = sq(RM) + sq(a);   // (R + M)^2 + a^2
    const d1 = ((m1 * (a1 - a2 + a) + R * a) * S + m2 * a1 * sq(a)) / sq(S);
    const d2 = ((m2 * (a2 - a1 + a) + R * a) * S + m1 * a2 * sq(a)) / sq(S);
    // rod half-lengths (complex for hyperextreme sources)
    const s1 = CMath.csqrt(sq(m1) - sq(a1) + 4 * m2 * a1 * d1);
    const s2 = CMath.csqrt(sq(m2) - sq(a2) + 4 * m1 * a2 * d2);

    // distances to the rod ends: lower hole (capital R), upper hole (small r)
    const dist = offset => CMath.psqrt(sq(rho) + sq(z + offset));
    const Rsp = dist(0.5 * R + s2), Rsn = dist(0.5 * R - s2);
    const rsp = dist(-0.5 * R + s1), rsn = dist(-0.5 * R - s1);

    const mu0 = (RM - I * a) / (RM + I * a);
    const phase1 = s => ((s - m1 - I * a1) * S + 2 * a1 * (m1 * a + I * M * RM)) / ((s - m1 + I * a1) * S + 2 * a1 * (m1 * a - I * M * RM));
    const phase2 = s => ((s + m2 - I * a2) * S - 2 * a2 * (m2 * a - I * M * RM)) / ((s + m2 + I * a2) * S - 2 * a2 * (m2 * a + I * M * RM));
    const rp = (1 / mu0) * phase1(s1) * rsp, rn = (1 / mu0) * phase1(-s1) * rsn;
    const Rp = -mu0 * phase2(s2) * Rsp, Rn = -mu0 * phase2(-s2) * Rsn;

    const R2 = sq(R), s1s = sq(s1), s2s = sq(s2);
    const A = (R2 - sq(s1 + s2)) * (Rp - Rn) * (rp - rn) - 4 * s1 * s2 * (Rp - rn) * (Rn - rp);
    const Bc = 2 * s1 * (R2 - s1s + s2s) * (Rn - Rp) + 2 * s2 * (R2 + s1s - s2s) * (rn - rp) + 4 * R * s1 * s2 * (Rp + Rn - rp - rn);
    const G = -z * Bc + s1 * (R2 - s1s + s2s) * (Rn - Rp) * (rp + rn + R) + s2 * (R2 + s1s - s2s) * (rn - rp) * (Rp + Rn - R)
              - 2 * s1 * s2 * (2 * R * (rp * rn - Rp * Rn - s1 * (rn - rp) + s2 * (Rn - Rp)) + (s1s - s2s) * (rp + rn - Rp - Rn));
    const K0 = (S * (R2 - sq(m1 - m2) + sq(a)) - 4 * sq(m1) * sq(m2) * sq(a)) / (m1 * m2 * S);

    const AA = CMath.self_conjugate_multiply(A), BB = CMath.self_conjugate_multiply(Bc);
    const sum_conj = CMath.conjugate(A) + CMath.conjugate(Bc);
    const w = $pin(2 * a - (2 * CMath.Imaginary(G * sum_conj) / (AA - BB)));
    const f = $pin((AA - BB) / CMath.Real((A + Bc) * sum_conj));
    const e2gamma = $pin((AA - BB) / CMath.Real(16 * sq(CMath.fabs(s1)) * sq(CMath.fabs(s2)) * sq(K0) * Rsp * Rsn * rsp * rsn));

    const g = [];
    g.length = 16;
    g[0] = -f;
    g[2] = f * w * 2 * 0.5;
    g[8] = g[2];
    g[5] = (1 / f) * e2gamma;
    g[10] = (1 / f) * sq(rho) + sq(w) * -f;
    g[15] = g[5];
    return g;
}

two_spinning_holes
