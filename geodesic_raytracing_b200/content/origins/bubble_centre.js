// distance from (t, r, theta, phi) to the centre of the warp bubble at x = velocity * t
function distance_to_bubble(t, r, theta, phi) {
    $cfg.velocity.$default = 2;
    const planar = r * CMath.sin(theta);
    const dx = planar * CMath.cos(phi) - $cfg.velocity * t;
    return CMath.fast_length(dx, planar * CMath.sin(phi), r * CMath.cos(theta));
}

distance_to_bubble
