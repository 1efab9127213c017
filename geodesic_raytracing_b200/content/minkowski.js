// Flat spacetime, signature (-,+,+,+), cartesian (t, x, y, z): ds^2 = -dt^2 + dx^2 + dy^2 + dz^2
(t, x, y, z) => [-1, 1, 1, 1]
