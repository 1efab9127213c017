#!/usr/bin/env python3
"""Where the warps of k_integrate wait: per-SASS-instruction stall samples of one launch of an `ncu --set full --import-source on`
report, summed by stall reason and listed for the instructions with the most samples.
usage: ncu_stall_sites.py <report.ncu-rep> [launch index, default 0 = the 1/16 prepass: a few lone warps] [top N]"""
import collections
import csv
import io
import subprocess
import sys


def main(path, which=0, top=40):
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for line in csv.reader(io.StringIO(src)):
        if line and line[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
        elif cur is not None and line:
            cur.append(line)
    blk = blocks[which]
    h = {n: i for i, n in enumerate(blk[0])}
    reasons = [n for n in blk[0] if n.startswith("stall_") and "Not Issued" not in n]
    tot = collections.Counter()
    rows = []
    for r in blk[1:]:
        n = int(r[h["# Samples"]] or 0)
        ex = int(r[h["Instructions Executed"]] or 0)
        per = {k: int(r[h[k]] or 0) for k in reasons}
        for k, v in per.items():
            tot[k] += v
        rows.append((n, ex, r[h["Source"]].strip(), per))
    all_samples = sum(n for n, *_ in rows)
    print(f"launch {which}: {all_samples} samples")
    print("by reason: " + "  ".join(f"{k[6:]} {v} ({100 * v / max(all_samples, 1):.1f}%)" for k, v in tot.most_common(10)))
    steps = collections.Counter(ex for _, ex, s, _ in rows if s.startswith(("FFMA", "FMUL"))).most_common(1)[0][0]
    print(f"warp-steps {steps}; samples per warp-step {all_samples / max(steps, 1):.3f}")
    for n, ex, s, per in sorted(rows, key=lambda t: -t[0])[:top]:
        why = "  ".join(f"{k[6:]} {v}" for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"{n:8d} {100 * n / max(all_samples, 1):5.1f}%  x{ex:<10d} {s[:60]:60s} {why}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
