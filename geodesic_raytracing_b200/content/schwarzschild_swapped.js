// Schwarzschild exterior (r_s = 1) with the first two coordinates swapped: native order is (r, t, theta, phi), so the
// timelike coordinate is not in slot 0 - a test bed for the tetrad construction's timelike-axis detection.
function swapped(r, t, theta, phi) {
    const f = 1 - 1 / r;
    const s = CMath.sin(theta);
    return [1 / f, -f, r * r, r * r * s * s];
}

swapped
