// Kerr black hole in Kerr-Schild cartesian coordinates (t, x, y, z): g = eta + H l (x) l with the
// null covector l = (1, (r x + a y)/(r^2 + a^2), (r y - a x)/(r^2 + a^2), z / r) and
// H = r_s r^3 / (r^4 + a^2 z^2), where r solves (x^2 + y^2)/(r^2 + a^2) + z^2/r^2 = 1.
// Parameters: spin a (default -0.5), Schwarzschild radius rs (default 1).
function kerr_in_kerr_schild_form(t, x, y, z) {
    $cfg.a.$default = -0.5;
    $cfg.rs.$default = 1;
    const spin = $cfg.a, rs = $cfg.rs;
    const spin2 = spin * spin;

    const rho2 = $pin(x * x + y * y + z * z);
    const eq2 = $pin(x * x + y * y - z * z);
    // r^2 = (rho^2 - a^2 + sqrt((rho^2 - a^2)^2 + 4 a^2 z^2)) / 2; the radicand is expanded so that
    // it reads a^4 - 2 a^2 (x^2 + y^2 - z^2) + rho^4
    const r2 = $pin((CMath.sqrt(spin2 * spin2 - 2 * spin2 * eq2 + rho2 * rho2) + rho2 - spin2) / 2);
    const r = CMath.sqrt(r2);

    const l = [1, (r * x + spin * y) / (r2 + spin2), (r * y - spin * x) / (r2 + spin2), z / r];
    const H = $pin(rs * r2 * r / (r2 * r2 + spin2 * z * z));

    const eta = [-1, 1, 1, 1];
    const g = [];
    for (let i = 0; i < 4; i++) {
        for (let j = 0; j < 4; j++) {
            g.push((i == j ? eta[i] : 0) + H * l[i] * l[j]);
        }
    }
    return g;
}

kerr_in_kerr_schild_form
