#!/usr/bin/env python3
"""Generate tests/golden/reference_metric_vectors.json from the reference's own metric scripts (test infrastructure).

Run in a container where /root/reference is mounted. For each listed script the reference file is loaded
unchanged through the plugin boundary, the CPU oracle is built on the generated functions, and g, dg, the
reference-formulation acceleration (GEO_ACCEL) and the coordinate maps are evaluated at seeded points.
`qjs_g` / `qjs_dg` come from the REFERENCE'S OWN JavaScript engine: the vendored QuickJS compiled from
/root/reference/deps/quickjs, running number.js and the script with numeric dual numbers (oracle/ref_quickjs.cpp).
The bundled content (geodesic_raytracing_b200/content) is tested against these vectors on machines that do
not have the reference (tests/test_reference_scripts.py)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import geodesic_raytracing_b200 as grt  # noqa: E402
import oracle  # noqa: E402

REFERENCE = "/root/reference/scripts"
NAMES = ["minkowski", "schwarzschild_fast", "schwarzschild", "schwarzschild_skewed", "schwarzschild_ingoing_ef", "wormhole (morris-thorne)", "kerr_schild", "alcubierre", "double_unequal_kerr", "kerr_boyer", "schwarzschild_ingoing_ef_hawking", "configurable_wormhole",
         "krasnikov_tube", "misner_4d", "godel_cylindrical", "kerr_newman_schild", "double_kerr"]


def points(name, n=4):
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    pos = rng.uniform(0.6, 2.5, (n, 4))
    vel = rng.uniform(-1, 1, (n, 4))
    vel[:, 0] = -1
    return pos.astype(np.float32), vel.astype(np.float32)


def script_of(name):
    """js file of the metric whose JSON carries this friendly name (content_manager.cpp:263-278)."""
    for f in sorted(os.listdir(REFERENCE)):
        if f.endswith(".json"):
            try:
                if json.load(open(os.path.join(REFERENCE, f))).get("name") == name:
                    return os.path.join(REFERENCE, f[:-5] + ".js")
            except Exception:
                pass
    raise KeyError(name)


def main():
    content = grt.Content(REFERENCE)
    qjs = oracle.QuickJSReference()     # the reference's own JS engine (vendored QuickJS) on the reference's own scripts
    out = {}
    for name in NAMES:
        m = content.load_by_name(name)
        o = oracle.for_metric(m)
        js = script_of(name)
        pos, vel = points(name)
        rec = {"params": m.param_names, "defaults": m.param_defaults, "info": {k: getattr(m.info, k) for k in
               ["is_big", "is_constant_theta", "use_prepass", "singular", "adaptive_precision", "detect_singularities", "coordinate_system", "has_periodicity"]},
               "w_v": list(m.info.w_v), "points": []}
        for p, v in zip(pos, vel):
            if m.info.is_constant_theta:
                p = p.copy(); v = v.copy()
                p[2] = np.float32(np.pi / 2)
                v[2] = 0
            g, dg = o.metric_eval(p)
            a = o.accel(p, v)
            c = o.coords(p, v)
            qg, qdg, _, _ = qjs.metric(js, p.astype(np.float64))
            rec["points"].append({"pos": p.tolist(), "vel": v.tolist(), "g": g.tolist(), "dg": dg.tolist(), "qjs_g": qg.tolist(), "qjs_dg": qdg.tolist(), "accel_ref": a["ref"].tolist(),
                                  "to_polar": c["to_polar"].tolist(), "from_polar": c["from_polar"].tolist(), "distance": c["distance"], "period": c["period"].tolist()})
        out[name] = rec
        print(name, "ok")
    path = os.path.join(ROOT, "tests", "golden", "reference_metric_vectors.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
