#!/usr/bin/env python3
"""Static look at one Verlet step of k_integrate in a compiled metric program (no GPU needed).

Follows the step's path through the SASS of the integration loop (from the loop head to the back edge; a forward
conditional branch is taken when the straight-line block behind it holds a store, an atomic, a vote or a call, i.e. the
refill / finish / slow paths are skipped) and prints
  * the number of instructions on that path, by opcode (what ncu's source page reports per warp-step, give or take
    the turn-parity branch), and
  * a dependency-chain estimate: the longest register-dependency path through the same instructions including the
    loop-carried state, with fixed latencies (4 cycles FP32/ALU, 5 across pipes ignored, MUFU 18, LDC 30) - what a lone
    warp's step cannot go below.
usage: sass_step_static.py <program.cubin> [-v]"""
import collections
import re
import subprocess
import sys

LAT = collections.defaultdict(lambda: 4, {"MUFU": 18, "LDC": 30, "LDCU": 12, "BRA": 0, "BSSY": 0, "BSYNC": 0, "NOP": 0, "YIELD": 0, "WARPSYNC": 0,
                                          "VOTE": 8, "S2R": 20})
BRANCH_BUBBLE = 0.0
ISSUE_INTERVAL = 2.0   # a lone warp's FP32 instruction occupies the 16-lane pipe of its scheduler for two cycles
taken = set()
RARE = ("STG", "ATOM", "ATOMG", "RED", "VOTE", "VOTEU", "CALL", "LDG", "SHFL", "NANOSLEEP", "EXIT")


def load(cubin):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", "k_integrate", cubin], capture_output=True, text=True, check=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def parse(text):
    pred = None
    t = text
    if t.startswith("@"):
        pred, t = t.split(None, 1)
    op, _, rest = t.partition(" ")
    return pred, op, rest


def regs(s):
    out = []
    for m in re.finditer(r"\b(UR|R|UP|P)(\d+)\b(\.64)?", s):
        kind, n, wide = m.group(1), int(m.group(2)), m.group(3)
        out.append(f"{kind}{n}")
        if wide:
            out.append(f"{kind}{n + 1}")
    return out


def operands(op, rest):
    """destination registers, source registers"""
    base = op.split(".")[0]
    parts = [p.strip() for p in rest.split(",")] if rest else []
    ndst = 1
    if base in ("FSETP", "ISETP", "PLOP3", "DSETP", "HSETP2"):
        ndst = 2
    if base in ("BRA", "BSSY", "BSYNC", "STG", "EXIT", "NOP", "YIELD", "WARPSYNC", "BREAK", "RED"):
        ndst = 0
    if base == "LOP3" and parts and parts[0].startswith("P"):
        ndst = 2
    if base in ("IADD3",) and len(parts) > 1 and parts[1] == "PT":
        pass
    dst, src = [], []
    for i, p in enumerate(parts):
        (dst if i < ndst else src).extend(regs(p))
    if base in ("CS2R",) or ".64" in op or ".128" in op:
        dst = dst + [f"R{int(d[1:]) + 1}" for d in dst if d.startswith("R")]
    return [d for d in dst if d not in ("RZ", "PT")], src


def step_path(ins):
    addr = {a: i for i, (a, _) in enumerate(ins)}
    # the loop's back edge: the last unconditional backward BRA whose target precedes most of the code
    back, span = None, 0
    for i, (a, t) in enumerate(ins):
        pred, op, rest = parse(t)
        if op == "EXIT":
            break                                           # what follows are the out-of-line stubs of divergent shuffles
        if op == "BRA" and pred is None:
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m and a - int(m.group(1), 16) > span:        # the longest backward jump is the integration loop
                back, span = (i, int(m.group(1), 16)), a - int(m.group(1), 16)
    assert back, "no loop back edge found"
    head, tail = addr[back[1]], back[0]
    # the step is the path from the loop head to the back edge that executes the most FP32 arithmetic without touching a store, an
    # atomic, a vote or a shuffle (refill, finish and slow paths): dynamic programming over the forward edges of the loop body
    NEG = -10 ** 9
    best = {tail: (0, None)}
    for i in range(tail - 1, head - 1, -1):
        a, t = ins[i]
        pred, op, rest = parse(t)
        base = op.split(".")[0]
        w = (1 if base in ("FFMA", "FMUL", "FADD") else 0) - (10 ** 6 if base in RARE else 0)
        succ = []
        if base == "BRA" and ".DIV" not in op:
            tgt = addr.get(int(re.search(r"0x([0-9a-f]+)", rest).group(1), 16))
            if tgt is not None and i < tgt <= tail:
                succ.append(tgt)
            if pred is not None or ".U" in op:
                succ.append(i + 1)
        elif base != "EXIT" or pred is not None:
            succ.append(i + 1)
        cands = [(best[j][0], j) for j in succ if j in best]
        best[i] = (w + max(cands)[0], max(cands)[1]) if cands else (NEG, None)
    path, i = [], head
    while i is not None:
        path.append(ins[i])
        i = best[i][1]
    return path


def main(cubin, verbose=False):
    ins = load(cubin)
    path = step_path(ins)
    mix = collections.Counter()
    for a, t in path:
        _, op, _ = parse(t)
        mix[op.split(".")[0]] += 1
    total = sum(mix.values())
    fp = mix["FFMA"] + mix["FMUL"] + mix["FADD"]
    print(f"{cubin}")
    print(f"step path: {total} instructions, {fp} FP32 (FFMA {mix['FFMA']} FMUL {mix['FMUL']} FADD {mix['FADD']}), MUFU {mix['MUFU']}")
    print("  " + "  ".join(f"{k} {v}" for k, v in mix.most_common()))
    # three consecutive steps (loop-carried state), report the last step's increment of
    #  (a) the pure dataflow chain and (b) an in-order issue model: a warp issues in program order, one instruction per
    #  cycle at best, and waits until the sources of the next instruction are ready - the time of a lone warp's step
    ready = collections.defaultdict(float)      # dataflow
    ready_io = collections.defaultdict(float)   # in-order
    t_issue, finish, finish_io, trace = 0.0, [], [], []
    for rep in range(3):
        trace = []
        for a, t in path:
            pred, op, rest = parse(t)
            base = op.split(".")[0]
            dst, src = operands(op, rest)
            if pred:
                src = src + regs(pred)
            start = max([ready[s_] for s_ in src] + [0.0])
            for d in dst:
                ready[d] = start + LAT[base]
            want = max([ready_io[s_] for s_ in src] + [0.0])
            stall = max(0.0, want - (t_issue + ISSUE_INTERVAL))
            t_issue = max(t_issue + ISSUE_INTERVAL, want)
            if base in ("BRA",) and (pred is None or True):
                t_issue += BRANCH_BUBBLE if (a, t) in taken else 0.0
            for d in dst:
                ready_io[d] = t_issue + LAT[base]
            trace.append((t_issue, stall, a, t))
        finish.append(max(ready.values()))
        finish_io.append(t_issue)
    per_step = finish[2] - finish[1]
    per_step_io = finish_io[2] - finish_io[1]
    print(f"dataflow chain: {per_step:.0f} cycles per step; in-order issue model: {per_step_io:.0f} cycles per step ({per_step_io / max(total, 1):.2f} per instruction)")
    if verbose:
        t0 = finish_io[1]
        for ti, stall, a, t in trace:
            print(f"   {ti - t0:7.0f} {('+%d' % stall) if stall else '':>5}  {a:05x}  {t}")
    per_step = per_step_io
    return total, fp, per_step


if __name__ == "__main__":
    main(sys.argv[1], "-v" in sys.argv[2:])
