#!/usr/bin/env python
"""SASS evidence for profiles/: per-kernel instruction counts AND the listing of each hot kernel's inner loop.

    python tools/sass_evidence.py [libsbce.so] > profiles/rNN_sass_evidence.md

Counts: DMMA = FP64 tensor-pipe MMA (DMMA.8x8x4; mma.sync.m16n8k8.f64 is issued as four of them),
UBLKCP = TMA bulk copy (cp.async.bulk), SYNCS = mbarrier operations, LDGSTS = cp.async, BAR = CTA barriers.
Excerpts: for the kernels listed in HOT the basic block with the most DMMA instructions (the steady-state
inner loop: blocks ending in a backward branch are preferred), instruction text only, so that the operand feeding (LDS / LDGSTS / UBLKCP) next to the MMAs is
visible, not just counted.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    ROOT, "semi-blind-channel-estimation-for-mimo-ris-communication-system-using-em-algo_b200", "libsbce.so")
MNEMS = ["DMMA", "UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DADD", "LDS", "SHFL", "BAR"]
HOT = ["k_gram_tma4", "k_chol_solve", "k_heff_qr_mma<4, 4>", "k_enum<4, 4, false"]


def demangle(n):
    out = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    out = re.sub(r"\(.*", "", out)
    return out.replace("void sbce::", "").replace("sbce::", "")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle(m.group(1))
            kernels[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and cur is not None:
            kernels[cur].append((int(m.group(1), 16), m.group(2).strip()))
    print("# SASS evidence (cuobjdump -sass %s, sm_100a)\n" % os.path.basename(LIB))
    print("Static instruction counts per kernel: `DMMA` = FP64 tensor-pipe MMA (`DMMA.8x8x4`), `UBLKCP` = TMA bulk copy")
    print("(`cp.async.bulk`), `SYNCS` = mbarrier operations, `LDGSTS` = `cp.async`, `BAR` = CTA barriers.  No `UTC*MMA` /")
    print("TMEM instruction can appear: the path is complex FP64 and `tcgen05.mma` has no f64 kind (DESIGN.md section 4).\n")
    print("| kernel | instr | " + " | ".join(MNEMS) + " |")
    print("|---|---:|" + "---:|" * len(MNEMS))
    for name, ins in sorted(kernels.items()):
        if not any(k in name for k in ("k_gram", "k_chol", "k_heff", "k_enum", "k_pm", "k_ls", "k_gen")):
            continue
        c = [sum(1 for _, t in ins if re.search(r"(^|\s)%s(\.|\s|$)" % mn, t)) for mn in MNEMS]
        print("| `%s` | %d | %s |" % (name, len(ins), " | ".join(str(x) for x in c)))
    print()
    for hot in HOT:
        for name, ins in kernels.items():
            if not name.startswith(hot):
                continue
            # basic blocks = runs between branch targets / branches; pick the one with most DMMA
            targets = set()
            for _, t in ins:
                m = re.search(r"BRA(?:\.\w+)*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", t)
                if m:
                    targets.add(int(m.group(1), 16))
            blocks, curb = [], []
            for addr, t in ins:
                if addr in targets and curb:
                    blocks.append(curb)
                    curb = []
                curb.append((addr, t))
                if re.search(r"(^|\s)(BRA|EXIT|RET)", t):
                    blocks.append(curb)
                    curb = []
            if curb:
                blocks.append(curb)
            def is_loop(bl):   # ends in a backward branch
                m = re.search(r"BRA(?:\.\w+)*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", bl[-1][1])
                return bool(m) and int(m.group(1), 16) <= bl[-1][0]

            best = max(blocks, key=lambda bl: sum(1 for _, t in bl if "DMMA" in t or "UBLKCP" in t) * (1.0 if is_loop(bl) else 0.3))
            nd = sum(1 for _, t in best if "DMMA" in t)
            if nd == 0 and "enum" not in name:
                continue
            if "enum" in name:
                best = max(blocks, key=lambda bl: sum(1 for _, t in bl if "DFMA" in t or "DADD" in t))
            print("## `%s`: hottest basic block (%d instructions, %d DMMA)\n" % (name, len(best), nd))
            print("```")
            for addr, t in best[:160]:
                if t.startswith("NOP"):
                    continue
                print("/*%05x*/ %s" % (addr, t))
            if len(best) > 160:
                print("... (%d more)" % (len(best) - 160))
            print("```\n")


if __name__ == "__main__":
    main()
