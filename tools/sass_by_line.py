#!/usr/bin/env python
"""Join an ncu source-page CSV (per-SASS-instruction executed counts) with nvdisasm -g line info.

usage: sass_by_line.py <file.ncu-rep> <cubin> <mangled-kernel-substring> [races]
Prints executed warp-instructions per source line (and per inlined call-site chain root), divided by `races`.
"""
import csv, io, re, subprocess, sys, collections

rep, cubin, kname = sys.argv[1:4]
races = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
base = int(rows[0]["Address"], 16)
counts = {int(r["Address"], 16) - base: (int(r["Instructions Executed"]), int(r["# Samples"] or 0), r["Source"].strip()) for r in rows}

dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
in_k = False
cur = None
by_line = collections.Counter(); samp_line = collections.Counter(); by_op = collections.Counter()
per_line_ops = collections.defaultdict(collections.Counter)
total = 0
for l in dis:
    if l.startswith(".text."):
        in_k = kname in l
        continue
    if not in_k:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f = m.group(1).split("/")[-1]
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        root = (inl[-1][0].split("/")[-1], int(inl[-1][1])) if inl else (f, int(m.group(2)))
        cur = (root, (f, int(m.group(2))))
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur:
        off = int(m.group(1), 16)
        if off in counts:
            c, s, src = counts[off]
            op = m.group(2).split()[0] if not m.group(2).startswith("@") else m.group(2).split()[1]
            op = op.split(".")[0]
            by_line[cur[0]] += c; samp_line[cur[0]] += s; by_op[op] += c; total += c
            per_line_ops[cur[0]][op] += c
print(f"total executed warp-instr: {total}  per race: {total / races:.1f}")
print("-- by root source line (file:line  instr/race  share  stall-samples)")
tot_s = sum(samp_line.values()) or 1
for k, c in sorted(by_line.items(), key=lambda kv: kv[0]):
    if c / races < 0.5: continue
    ops = " ".join(f"{o}:{v / races:.0f}" for o, v in per_line_ops[k].most_common(6))
    print(f"{k[0]}:{k[1]:4d}  {c / races:9.1f}  {100.0 * c / total:5.1f}%  samp {100.0 * samp_line[k] / tot_s:5.1f}%   {ops}")
print("-- by opcode")
for o, c in by_op.most_common(30):
    print(f"{o:10s} {c / races:9.1f} {100.0 * c / total:5.1f}%")
