// Charged rotating black hole (Kerr-Newman) in Kerr-Schild cartesian coordinates (t, x, y, z): g = eta + H l (x) l with the
// same null covector l as the Kerr case and H = (r_s r - Q^2) r^2 / (r^4 + a^2 z^2).
// Parameters: spin a (default -0.51), Schwarzschild radius rs (default 1), charge Q (default 0.51).
function kerr_newman_in_kerr_schild_form(t, x, y, z) {
    $cfg.a.$default = -0.51;
    $cfg.rs.$default = 1;
    $cfg.Q.$default = 0.51;
    const spin = $cfg.a, rs = $cfg.rs, charge = $cfg.Q;
    const spin2 = spin * spin;

    const rho2 = $pin(x * x + y * y + z * z);
    const eq2 = $pin(x * x + y * y - z * z);
    const r2 = $pin((CMath.sqrt(spin2 * spin2 - 2 * spin2 * eq2 + rho2 * rho2) + rho2 - spin2) / 2);
    const r = CMath.sqrt(r2);

    const l = [1, (r * x + spin * y) / (r2 + spin2), (r * y - spin * x) / (r2 + spin2), z / r];
    const H = $pin((rs * r - charge * charge) * r2 / (r2 * r2 + spin2 * z * z));

    const eta = [-1, 1, 1, 1];
    const g = [];
    for (let i = 0; i < 4; i++) {
        for (let j = 0; j < 4; j++) {
            g.push((i == j ? eta[i] : 0) + H * l[i] * l[j]);
        }
    }
    return g;
}

kerr_newman_in_kerr_schild_form
