// distance from (t, r, theta, phi) to the object at the coordinate origin
(t, r, theta, phi) => r
