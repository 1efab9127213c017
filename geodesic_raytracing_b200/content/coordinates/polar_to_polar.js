// identity map for metrics already written in (t, r, theta, phi)
(t, r, theta, phi) => [t, r, theta, phi]
