#!/usr/bin/env python
"""Source-level stall attribution from an `ncu --set full --import-source on` report, as a small tracked table.

    python tools/source_stalls.py gpurun_out/r02z_full.ncu-rep > profiles/r02z_source_stalls.md

For each hot kernel: the stall-reason mix of the warp samples and the source lines that hold the most samples
(share of the kernel's samples, share of its executed instructions, dominant stall reasons).  The report itself
(tens of MB) stays in gpurun_out/.
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "semi-blind-channel-estimation-for-mimo-ris-communication-system-using-em-algo_b200", "csrc")
KERNELS = [("k_chol_solve", "chol.cu"), ("k_gram_tma4", "mstep.cu"), ("k_enum", "estep.cu"), ("k_heff_qr_mma", "estep.cu")]
NAMES = ["stall_long_sb", "stall_wait", "stall_math", "stall_barrier", "stall_short_sb", "stall_sleep", "stall_selected",
         "stall_not_selected", "stall_branch_resolving", "stall_no_inst", "stall_dispatch", "stall_mio", "stall_lg"]


def toi(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep = sys.argv[1]
    print("# Source-level stall attribution (`ncu --page source` of `%s`)\n" % os.path.basename(rep))
    print("Warp-state samples per source line of the final kernels (B200, default bench workload, 1184 trials per launch).")
    print("`samples` = share of the kernel's warp samples, `instr` = share of its executed warp instructions; the reasons")
    print("listed are those holding more than 20 % of the line's samples.  Lines of inlined helpers are attributed to the")
    print("helper's own file (`common.cuh`, `tensor.cuh`).\n")
    for kern, main_file in KERNELS:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern,
                              "--print-source=cuda,sass"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        heads = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
        if not heads:
            continue
        data = []
        for k, hi in enumerate(heads):
            f = rows[hi - 2][1].split("/")[-1]
            hdr = rows[hi]
            S, IE = hdr.index("# Samples"), hdr.index("Instructions Executed")
            idx = {n: hdr.index(n) for n in NAMES if n in hdr}
            end = heads[k + 1] - 2 if k + 1 < len(heads) else len(rows)
            for r in rows[hi + 1:end]:
                if r and r[0].strip().isdigit() and len(r) > max(idx.values()):
                    n = toi(r[S])
                    if n:
                        data.append((f, int(r[0]), n, {kk: toi(r[i]) for kk, i in idx.items()}, toi(r[IE])))
        tot = sum(d[2] for d in data)
        totie = max(1, sum(d[4] for d in data))
        agg = collections.Counter()
        for _, _, _, st, _ in data:
            agg.update(st)
        fn = rows[heads[0] - 1][1]
        print("## `%s`\n" % fn.split("(")[0].replace("void sbce::", "").replace("sbce::", ""))
        print("stall mix: " + ", ".join("%s %.1f %%" % (k.replace("stall_", ""), 100.0 * v / tot)
                                        for k, v in agg.most_common() if v / tot > 0.01) + "\n")
        try:
            src = open(os.path.join(CSRC, main_file)).read().splitlines()
        except OSError:
            src = []
        print("| file:line | samples | instr | source | dominant stalls |")
        print("|---|---:|---:|---|---|")
        for f, ln, n, st, ie in sorted(data, key=lambda x: -x[2])[:18]:
            text = src[ln - 1].strip()[:80].replace("|", "\\|") if f == main_file and ln <= len(src) else ""
            dom = ", ".join("%s %d %%" % (k.replace("stall_", ""), round(100.0 * v / n)) for k, v in st.items() if v > 0.2 * n)
            print("| `%s:%d` | %.1f %% | %.1f %% | `%s` | %s |" % (f, ln, 100.0 * n / tot, 100.0 * ie / totie, text, dom))
        print()


if __name__ == "__main__":
    main()
