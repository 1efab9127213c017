// Misner (T, phi, y, z) -> Minkowski (t, x) = (T e^{phi/2} -+ e^{-phi/2}) (arXiv:1102.0907 eqs. 8, 9 inverted) -> polar (t, r, theta, phi)
(T, mphi, y, z) => {
    const grow = CMath.exp(mphi / 2), shrink = CMath.exp(-mphi / 2);
    const t = T * grow - shrink, x = T * grow + shrink;
    return [t, CMath.sqrt(x * x + y * y + z * z), CMath.atan2(CMath.sqrt(x * x + y * y), z), CMath.atan2(y, x)];
}
