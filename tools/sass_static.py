#!/usr/bin/env python
"""Static SASS statistics of one kernel (no GPU needed): code size, opcode histogram, and the loops found from
backward branches (address range, instruction count, opcode mix).  A proxy to read before spending GPU time;
executed counts come from ncu (tools/sass_by_line.py).

usage: sass_static.py <file.o|.so|.cubin> <kernel-name-substring> [--loops N]
"""
import collections
import re
import subprocess
import sys

obj, kname = sys.argv[1:3]
n_loops = int(sys.argv[sys.argv.index("--loops") + 1]) if "--loops" in sys.argv else 4
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if kname not in name:
        continue
    ops = collections.Counter()
    for _, t in ins:
        f = t.split()
        op = f[1] if f[0].startswith("@") else f[0]
        ops[op.split(".")[0]] += 1
    print(f"{name}\n  {len(ins)} instructions, {16 * len(ins)} bytes")
    print("  " + " ".join(f"{o}:{c}" for o, c in ops.most_common(24)))
    loops = []
    for a, t in ins:
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a:
                loops.append((a - tgt, tgt, a))
    for span, lo, hi in sorted(loops, reverse=True)[:n_loops]:
        body = [t for a, t in ins if lo <= a <= hi]
        bo = collections.Counter()
        for t in body:
            f = t.split()
            op = f[1] if f[0].startswith("@") else f[0]
            bo[op.split(".")[0]] += 1
        print(f"  loop {lo:#x}..{hi:#x}: {len(body)} instructions  " + " ".join(f"{o}:{c}" for o, c in bo.most_common(14)))
