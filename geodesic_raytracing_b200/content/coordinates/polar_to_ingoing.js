// (t, r, theta, phi) -> (v, r, theta, phi) with v = t + r + rs ln|r - rs|
(t, r, theta, phi) => [t + r + $cfg.rs * CMath.log(CMath.fabs(r - $cfg.rs)), r, theta, phi]
