#!/bin/bash
# compute-sanitizer memcheck + racecheck over one small frame per code path (prepass, adaptive sampling, stripes, geodesic camera)
cat > /tmp/grt_sanitize.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import geodesic_raytracing_b200 as grt
c = grt.Content(grt.CONTENT_DIR)
ctx = grt.Context(0)
ctx.set_background(grt.synthetic_background(256, 128))
for name in ("kerr_schild", "schwarzschild_fast", "alcubierre"):
    m = c.load_by_name(name)
    ctx.set_metric(m)
    ctx.set_camera()
    ctx.set_features(adaptive_sampling=1, use_prepass=1 if name == "kerr_schild" else -1, redshift=1 if name == "alcubierre" else 0)
    img, st = ctx.render(96, 64)
    part, _ = ctx.render(96, 64, stripes=[(0, 16), (32, 48)])
    assert np.array_equal(part[:16], img[:16]) and np.array_equal(part[16:], img[32:48])
    ctx.set_tuning(2, 8)
    img2, _ = ctx.render(96, 64)
    assert np.array_equal(img, img2)
    ctx.set_tuning(3, 0)
    if name == "alcubierre":
        n = ctx.snapshot_geodesic(0.2, True)
        ctx.set_camera_on_geodesic(1.0)
        ctx.render(64, 48)
        ctx.clear_camera_geodesic()
    print(name, "ok", st.trial_steps)
PY
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/grt_sanitize.py 2>&1 | tail -8
  echo "exit $?"
done
