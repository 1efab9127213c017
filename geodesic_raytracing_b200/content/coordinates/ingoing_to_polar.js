// (v, r, theta, phi) -> (t, r, theta, phi) with t = v - r*, r* = r + rs ln|r - rs| (rs is the metric's own parameter)
(v, r, theta, phi) => [v - (r + $cfg.rs * CMath.log(CMath.fabs(r - $cfg.rs))), r, theta, phi]
