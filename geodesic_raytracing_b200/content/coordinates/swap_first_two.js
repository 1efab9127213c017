// (a, b, theta, phi) -> (b, a, theta, phi): its own inverse, maps (r, t, ..) to (t, r, ..) and back
(a, b, theta, phi) => [b, a, theta, phi]
