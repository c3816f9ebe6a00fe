#!/usr/bin/env python
"""tools/sass_loops.py — static size of every SASS loop of a kernel, with the source lines it comes from.

    python tools/sass_loops.py <file.cubin|file.o> [function-substring]

Development aid (CPU only): compile one instantiation with -lineinfo, then look at how many instructions the
inner loops of the estimator / DFT really cost and which source lines they belong to."""
import re
import subprocess
import sys
from collections import Counter


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    dis = subprocess.run(["nvdisasm", "-g", "-c", path], capture_output=True, text=True).stdout.splitlines()
    func, cur = None, None
    ins = []      # (func, idx, label or None, text, (file, line))
    for l in dis:
        m = re.match(r"^(\S+):\s*$", l)
        if m and not m.group(1).startswith(".L"):
            func = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"^\s*(\.L_x_\d+):", l)
        if m:
            ins.append((func, m.group(1), None, cur))
            continue
        m = re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
        if m and func:
            ins.append((func, None, m.group(1).strip(), cur))
    labels = {}
    for i, (f, lab, t, c) in enumerate(ins):
        if lab:
            labels[(f, lab)] = i
    for i, (f, lab, t, c) in enumerate(ins):
        if not t or want not in (f or ""):
            continue
        m = re.search(r"BRA\s+`\((\.L_x_\d+)\)", t)
        if not m:
            continue
        tgt = labels.get((f, m.group(1)))
        if tgt is None or tgt >= i:
            continue
        body = [x for x in ins[tgt:i + 1] if x[2]]
        ops = Counter(re.sub(r"^@!?U?P\d+\s+", "", x[2]).split()[0].split(".")[0] for x in body)
        src = Counter(x[3] for x in body if x[3])
        fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
        top = ", ".join(f"{a}:{b}x{n}" for (a, b), n in src.most_common(4))
        print(f"{f[-60:]:60s} loop {len(body):4d} instr  fp64 {fp64:3d}  movs {ops['IMAD'] + ops['MOV']:3d}  lds {ops['LDS']:2d}  [{top}]")


if __name__ == "__main__":
    main()
