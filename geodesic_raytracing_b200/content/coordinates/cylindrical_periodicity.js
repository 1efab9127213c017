// periods of (t, rho, phi, z): only phi wraps
(t, rho, phi, z) => [0, 0, 2 * Math.PI, 0]
