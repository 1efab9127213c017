#!/usr/bin/env python3
"""Generate tests/golden/clref_frames.npz: outputs of the REFERENCE'S OWN KERNELS (test infrastructure).

Run in a container where /root/reference is mounted.  /root/reference/cl.cl is compiled as C++ for the host
(oracle/ref_build.py, oracle/ref_shim/) against the functions generated for each bundled metric, and the
reference's kernel sequence (main.cpp:2356-2526) is run on small frames.  The fixtures let machines without
the reference check the oracle restatement (tests/test_reference_kernels.py) and the CUDA path
(tests/test_gpu_parity.py) against output of the reference's own device code."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import geodesic_raytracing_b200 as grt  # noqa: E402
import oracle  # noqa: E402

W, H = 128, 96
BG_W, BG_H = 512, 256
# name -> list of (tag, feature overrides, frame kwargs)
CASES = {
    "minkowski": [("default", {}, {})],
    "schwarzschild": [("default", {}, {}), ("redshift_boosted", {"redshift": 1}, {"basis_speed": (0.3, 0.1, -0.2)})],
    "kerr_schild": [("default", {}, {}), ("full_sampling", {"adaptive_sampling": 0}, {}), ("camera2", {}, {"camera_pos": (0, 3, -6, 1.5), "camera_quat": (0.1, -0.6, 0.2, 0.7)})],
    "alcubierre": [("redshift", {"redshift": 1}, {})],
    "wormhole (morris-thorne)": [("flip", {}, {"flip": True})],
    "kerr_boyer": [("default", {}, {})],
    "double_unequal_kerr": [("default", {}, {})],
    # the families of the reference catalogue with code paths of their own
    "kerr_newman_schild": [("default", {}, {})],
    "double_kerr": [("default", {}, {})],
    "godel_cylindrical": [("default", {}, {"camera_pos": (0.0, 1.0, -3.0, 0.5)})],
    "godel_cylindrical_forward": [("forward", {}, {"camera_pos": (0.0, 1.0, -3.0, 0.5)})],
    "misner_4d": [("chronal", {}, {"camera_pos": (-2.0, 0.5, -4.0, 0.3)})],
}


def main():
    content = grt.Content(grt.CONTENT_DIR)
    bg = grt.synthetic_background(BG_W, BG_H)
    out = {}
    for name, cases in CASES.items():
        m = content.load_by_name(name)
        ref = oracle.reference_for_metric(m)
        for tag, feat_kw, kw in cases:
            feat = ref.features(max_acceleration_change=m.info.max_acceleration_change, **feat_kw)
            r = ref.frame(W, H, feat, background=bg, background2=bg[::-1].copy(), **kw)
            key = f"{name}|{tag}"
            rd = r["rdata"]
            out[key + "|image"] = r["image"]
            for f in ("tex_coord", "z_shift", "terminated", "side"):
                out[key + "|" + f] = np.ascontiguousarray(rd[f])
            out[key + "|counts"] = np.array([r["trial_steps"], r["rays_traced"], r["extra_rays"]], np.int64)
            if r["termination"] is not None:
                out[key + "|termination"] = r["termination"]
            out[key + "|tetrad"] = r["tetrad"]
            print(key, "steps", r["trial_steps"], "finished", float((rd["terminated"] == 1).mean()))
    # the geodesic camera path (main.cpp:2675-2760, 2264-2296)
    m = content.load_by_name("alcubierre")
    ref = oracle.reference_for_metric(m)
    small = ref.frame(16, 16, ref.features(max_acceleration_change=m.info.max_acceleration_change, adaptive_sampling=0))
    feat = ref.features(max_acceleration_change=m.info.max_acceleration_change)
    pos, vel, ds, tr, boosted = ref.geodesic_snapshot(small["camera_generic"], small["tetrad"], (0, 0, 0.2), feat, max_len=4000)
    out["alcubierre|geodesic|pos"], out["alcubierre|geodesic|vel"], out["alcubierre|geodesic|ds"] = pos, vel, ds
    out["alcubierre|geodesic|transported"], out["alcubierre|geodesic|boosted"] = tr, boosted
    cam, tet, v = ref.geodesic_interpolate(pos, vel, ds, tr, 2.0, 1, (0, 0, 0.2))
    out["alcubierre|geodesic|interp_cam"], out["alcubierre|geodesic|interp_tetrad"], out["alcubierre|geodesic|interp_vel"] = cam, tet, v
    path = os.path.join(ROOT, "tests", "golden", "clref_frames.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
