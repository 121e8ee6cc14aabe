#!/usr/bin/env python
"""Turn the scratch outputs of tools/gpu_final.sh (gpurun_out/<tag>_*) into the tracked evidence under profiles/:
bench lines, the ncu launch list, the `--set full` summary of the hot kernels and profiles/ncu_traffic.json
(per-launch DRAM bytes that bench.py reports as roofline.traffic).

    python tools/collect_profiles.py r02i
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]

for f in sorted(os.listdir(G)):
    if f.startswith(tag + "_bench") and f.endswith(".json") or f == tag + "_reference_arm.json" or f == tag + "_tests.log":
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
        print("copied", f)
launch = os.path.join(G, tag + "_launches.csv")
if os.path.exists(launch):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "launches", launch,
                    os.path.join(P, tag + "_launches.md"), "python bench.py --steps 1 --warmup 3 --kernels-only"],
                   stdout=subprocess.DEVNULL)
    print("wrote", tag + "_launches.md")
rep = os.path.join(G, tag + "_full.ncu-rep")
if os.path.exists(rep):
    out = os.path.join(P, tag + "_top_kernels_full.csv")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "full", rep, out], stdout=subprocess.DEVNULL)
    rows = list(csv.reader(open(out)))
    names = rows[1][2:]
    rd = [float(x) for x in next(r for r in rows if r[0] == "dram__bytes_read.sum")[2:]]
    wr = [float(x) for x in next(r for r in rows if r[0] == "dram__bytes_write.sum")[2:]]
    unit = next(r for r in rows if r[0] == "dram__bytes_read.sum")[1]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
    key = {"k_chol": "chol", "k_gram": "gram", "k_enum": "enum", "k_heff": "heff_qr"}
    traffic = {}
    for n, a, b in zip(names, rd, wr):
        for k, v in key.items():
            if k in n:
                traffic[v] = (a + b) * scale
    tj = os.path.join(P, "ncu_traffic.json")
    t = json.load(open(tj))
    bench = json.loads(open(os.path.join(G, tag + "_bench.json")).read().strip().splitlines()[-1])
    t["config_2"] = dict(source="profiles/%s_top_kernels_full.csv" % tag,
                         trials_per_launch=bench["config"]["trials_per_step_per_gpu"], dram_bytes_per_launch=traffic)
    json.dump(t, open(tj, "w"), indent=1)
    print("updated ncu_traffic.json", traffic)
