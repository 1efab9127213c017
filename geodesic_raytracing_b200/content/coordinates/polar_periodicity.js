// periods of (t, r, theta, phi): theta wraps at pi, phi at 2 pi, 0 = not periodic
(t, r, theta, phi) => [0, 0, Math.PI, 2 * Math.PI]
