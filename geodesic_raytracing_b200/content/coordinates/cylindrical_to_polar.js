// (t, rho, phi, z) -> (t, r, theta, phi)
(t, rho, phi, z) => [t, CMath.sqrt(rho * rho + z * z), CMath.atan2(rho, z), phi]
