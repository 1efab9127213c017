#!/usr/bin/env python3
"""NVRTC-compile every metric of a content folder for sm_100a (no GPU needed) and report registers / spills of k_integrate and the
SASS instructions on the path of one Verlet step (tools/sass_step_static.py) of the build that is used.
usage: compile_catalogue.py <content folder> [cache dir]"""
import os
import re
import subprocess
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import geodesic_raytracing_b200 as grt  # noqa: E402
import sass_step_static  # noqa: E402


def main():
    folder = sys.argv[1]
    cache = sys.argv[2] if len(sys.argv) > 2 else "/tmp/grt_catalogue_cache"
    os.makedirs(cache, exist_ok=True)
    os.environ["GRT_CACHE_DIR"] = cache
    content = grt.Content(folder)
    print(f"{'metric':36s} {'ops(accel)':>10s} {'ops(ref)':>9s} {'regs':>9s} {'stack':>6s} {'tier':>5s} {'instr/step':>10s} {'FP32':>5s} {'build s':>8s}")
    for name in content.metric_names():
        m = content.load_by_name(name)
        before = set(os.listdir(cache))
        t = time.time()
        m.compile(False, True)
        dt = time.time() - t
        added = set(os.listdir(cache)) - before
        new = [f for f in added if f.endswith(".cubin")]
        tier = [open(os.path.join(cache, f)).read().strip() for f in added if f.endswith(".tier")]
        found = []
        step = {}
        for f in new:   # the free compile and, if an occupancy tier was accepted, the tiered build that is actually used
            out = subprocess.run(["cuobjdump", "-res-usage", os.path.join(cache, f)], capture_output=True, text=True).stdout
            i = out.find("k_integrate")
            if i >= 0:
                found.append((int(re.search(r"REG:(\d+)", out[i:]).group(1)), int(re.search(r"STACK:(\d+)", out[i:]).group(1))))
                try:
                    import contextlib, io
                    with contextlib.redirect_stdout(io.StringIO()):
                        total, fp, _ = sass_step_static.main(os.path.join(cache, f))
                    step[found[-1][0]] = (total, fp)
                except Exception:      # noqa: BLE001  (a loop shape the path finder does not know)
                    pass
        found.sort(reverse=True)
        regs = "->".join(str(r) for r, _ in found) or "?"
        stack = str(found[-1][1]) if found else "?"
        info = m.info
        used = step.get(found[-1][0], ("?", "?")) if found else ("?", "?")
        print(f"{name:36s} {info.flops_accel:10d} {info.flops_accel_reference:9d} {regs:>9s} {stack:>6s} {(tier[0] if tier else '-'):>5s} {used[0]!s:>10s} {used[1]!s:>5s} {dt:8.1f}", flush=True)


if __name__ == "__main__":
    main()
