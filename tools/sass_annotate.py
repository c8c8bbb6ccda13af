#!/usr/bin/env python
"""Print a kernel's SASS with the source line (and the root line of the inlining chain) in front of each instruction.
usage: sass_annotate.py <cubin> <mangled-kernel-substring>"""
import re, subprocess, sys
cubin, kname = sys.argv[1:3]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
in_k = False; cur = ""
for l in dis:
    if l.startswith(".text."):
        in_k = kname in l; continue
    if not in_k: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f = m.group(1).split("/")[-1].replace("native_kernel.cu", "K").replace("native_math.cuh", "M").replace("sm_30_intrinsics.hpp", "i30").replace("sm_80_rt.hpp", "rt80")
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        root = inl[-1][1] if inl else ""
        cur = f"{f}:{m.group(2)}" + (f"<{root}" if root else "")
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: print(f"{cur:16s} {m.group(1)}  {m.group(2)}")
    elif l.startswith(".L_"): print(l)
