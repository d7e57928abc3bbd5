#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture (.ncu-rep read here, no GPU needed): DRAM bytes (read + write) per
launch of the value-pass kernel and per CG iteration of the persistent kernel, for bench.py's `roofline.traffic`.
usage: python tools/ncu_traffic.py <file.ncu-rep> <workload> <n_gpus> <cg iterations in the captured launch> [source label]"""
import csv
import json
import os
import subprocess
import sys

rep, workload, n_gpus, its = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
label = sys.argv[5] if len(sys.argv) > 5 else os.path.basename(rep)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
ki, ri, wi, ti = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
units = rows[1]


def to_bytes(v, u):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


recs = {}
for r in rows[2:]:
    name = r[ki]
    b = to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi])
    if "cg_persistent" in name:
        recs["cg_persistent_kernel"] = dict(workload=workload, n_gpus=n_gpus, kernel="cg_persistent_kernel", bytes_per_launch=b,
                                            iterations_in_launch=its, bytes_per_iteration=b / its, source=label)
    elif "assemble" in name:
        recs[name.split("<")[0].split("(")[0]] = dict(workload=workload, n_gpus=n_gpus, kernel="assemble", bytes_per_launch=b, source=label,
                                                      kernel_name=name[:80])
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
try:
    tab = json.load(open(path))
except Exception:
    tab = []
tab = [t for t in tab if not (t["workload"] == workload and t["n_gpus"] == n_gpus)] + list(recs.values())
json.dump(tab, open(path, "w"), indent=1)
print(json.dumps(list(recs.values()), indent=1))
