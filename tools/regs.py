#!/usr/bin/env python
"""Summarise ptxas -v logs: kernel, registers, stack, spills, smem."""
import glob, re, subprocess, sys, os
d = sys.argv[1] if len(sys.argv) > 1 else "."
for f in sorted(glob.glob(os.path.join(d, "*.ptxas.log"))):
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers.*?(?:(\d+) bytes smem)?\n", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void sbce::", "")
        print("%-14s %-40s regs %3s stack %4s spill %s/%s smem %s" % (os.path.basename(f)[:-10], name[:40], m.group(5), m.group(2), m.group(3), m.group(4), m.group(6)))
