#!/usr/bin/env python3
"""Per-step SASS instruction mix of k_integrate from an `ncu --set full --import-source on` report, and the DRAM bytes of
every profiled launch -> the JSON bench.py reads (profiles/<round>_ncu_traffic.json).
usage: ncu_step_mix.py <report.ncu-rep> <out.json> [launch index for the instruction mix, default 1 = first pass]"""
import collections
import csv
import io
import json
import subprocess
import sys


def pages(path, page, extra=()):
    out = subprocess.run(["ncu", "-i", path, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return out


def main(path, out_path, which=1):
    raw = list(csv.reader(io.StringIO(pages(path, "raw"))))
    hdr = raw[0]
    idx = {h: i for i, h in enumerate(hdr)}
    rows = raw[2:]
    unit = {h: raw[1][i] for h, i in idx.items()}

    def bytes_of(r, name):
        v, u = float(r[idx[name]]), unit[name]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

    dram = [bytes_of(r, "dram__bytes_read.sum") + bytes_of(r, "dram__bytes_write.sum") for r in rows]
    # source page: one block per launch, "Kernel Name" line first
    src = pages(path, "source", ("--print-source", "sass"))
    blocks, cur = [], None
    for line in csv.reader(io.StringIO(src)):
        if line and line[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
        elif cur is not None and line:
            cur.append(line)
    blk = blocks[which]
    h = {n: i for i, n in enumerate(blk[0])}
    execd = collections.Counter()
    counts = []
    for r in blk[1:]:
        op = r[h["Source"]].split()
        if not op:
            continue
        name = op[1] if op[0].startswith("@") else op[0]
        name = name.split(".")[0].rstrip(";")
        n = int(r[h["Instructions Executed"]])
        execd[name] += n
        if name in ("FFMA", "FMUL"):
            counts.append(n)
    steps = collections.Counter(counts).most_common(1)[0][0]        # executions of the step body (warp-steps)
    total = sum(execd.values())
    fp = execd["FFMA"] + execd["FMUL"] + execd["FADD"]
    res = {
        "kernel": "k_integrate",
        "source": f"{path} (ncu --set full, the integrate launches of one 4K kerr_schild frame: prepass, first pass, refinement)",
        "dram_bytes_per_launch": dram,
        "mean_dram_bytes_per_launch": sum(dram) / len(dram),
        "sass_instr_per_step": round(total / steps, 1),
        "sass_fp_instr_per_step": round(fp / steps, 1),
        "sass_mix_per_step": {k: round(v / steps, 1) for k, v in execd.most_common(14)},
        "warp_steps_in_launch": steps,
        "warp_instructions_per_frame": sum(float(r[idx["smsp__inst_executed.sum"]]) for r in rows),
        "pipe_fma_pct": float(rows[which][idx["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]]),
        "lanes_active_per_instruction": float(rows[which][idx["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
        "registers": int(float(rows[which][idx["launch__registers_per_thread"]])),
        "issue_active_pct": float(rows[which][idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
        "note_instr": "warp-level SASS instructions executed per warp-step in the first-pass launch (ncu source page: all executed instructions / executions of the step body); FP = FFMA + FMUL + FADD",
    }
    json.dump(res, open(out_path, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1)
