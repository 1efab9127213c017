// (t, r, theta, phi) -> (t, rho, phi, z)
(t, r, theta, phi) => [t, r * CMath.sin(theta), phi, r * CMath.cos(theta)]
