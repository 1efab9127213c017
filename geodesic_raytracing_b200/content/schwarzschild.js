// Schwarzschild exterior in Schwarzschild coordinates (t, r, theta, phi), geometric units, r_s = 1:
//   ds^2 = -(1 - 1/r) dt^2 + dr^2 / (1 - 1/r) + r^2 (dtheta^2 + sin^2(theta) dphi^2)
const schwarzschild_unit = function (t, r, theta, phi) {
    const lapse2 = 1 - 1 / r;
    const s = CMath.sin(theta);
    return [-lapse2, 1 / lapse2, r * r, r * r * s * s];
};

schwarzschild_unit
