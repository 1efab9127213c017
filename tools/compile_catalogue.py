#!/usr/bin/env python3
"""NVRTC-compile every metric of a content folder for sm_100a (no GPU needed) and report registers / spills of k_integrate.
usage: compile_catalogue.py <content folder> [cache dir]"""
import os
import re
import subprocess
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import geodesic_raytracing_b200 as grt  # noqa: E402


def main():
    folder = sys.argv[1]
    cache = sys.argv[2] if len(sys.argv) > 2 else "/tmp/grt_catalogue_cache"
    os.makedirs(cache, exist_ok=True)
    os.environ["GRT_CACHE_DIR"] = cache
    content = grt.Content(folder)
    print(f"{'metric':36s} {'ops(accel)':>10s} {'ops(ref)':>9s} {'regs':>5s} {'stack':>6s} {'build s':>8s}")
    for name in content.metric_names():
        m = content.load_by_name(name)
        before = set(os.listdir(cache))
        t = time.time()
        m.compile(False, True)
        dt = time.time() - t
        new = [f for f in set(os.listdir(cache)) - before if f.endswith(".cubin")]
        regs = stack = "?"
        if new:
            out = subprocess.run(["cuobjdump", "-res-usage", os.path.join(cache, new[0])], capture_output=True, text=True).stdout
            i = out.find("k_integrate")
            if i >= 0:
                regs = re.search(r"REG:(\d+)", out[i:]).group(1)
                stack = re.search(r"STACK:(\d+)", out[i:]).group(1)
        info = m.info
        print(f"{name:36s} {info.flops_accel:10d} {info.flops_accel_reference:9d} {regs:>5s} {stack:>6s} {dt:8.1f}", flush=True)


if __name__ == "__main__":
    main()
