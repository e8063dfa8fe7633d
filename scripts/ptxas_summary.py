"""Print registers / spill / stack of selected kernels from an `nvcc -Xptxas -v` log: python scripts/ptxas_summary.py ab_libs/x.ptxas.log [regex]"""
import re, sys, subprocess
log = open(sys.argv[1]).read().splitlines()
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"k_trace_rays|k_extend|k_connect")
cur = None; info = {}
for l in log:
    m = re.search(r"Compiling entry function '([^']+)'", l)
    if m: cur = m.group(1); info[cur] = []
    elif cur and ("registers" in l or "spill" in l): info[cur].append(l.replace("ptxas info    : ", "").strip())
names = list(info)
dem = subprocess.run(["c++filt", "--"] + (names or ["_none"]), capture_output=True, text=True).stdout.splitlines()
for n, d in zip(names, dem):
    if pat.search(d):
        print(d.split("(")[0][:70].ljust(70), " | ".join(info[n]))
