// (t, x, y, z) -> (t, r, theta, phi), theta measured from +z
(t, x, y, z) => {
    const planar2 = x * x + y * y;
    return [t, CMath.sqrt(planar2 + z * z), CMath.atan2(CMath.sqrt(planar2), z), CMath.atan2(y, x)];
}
