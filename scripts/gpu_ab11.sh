#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/ab11_pytest.log
bb() {
  echo "== build_bench lib=$1 scene=$2"
  FOUNDATION_PT_LIB=$PWD/ab_libs/$1.so timeout 300 python scripts/build_bench.py $2 2>&1 | tail -2
}
{
bb prev terrain
bb box32 terrain
bb box32b terrain
bb prev terrain
bb box32 terrain
bb box32b terrain
bb prev spheres
bb box32b spheres
bb prev instanced
bb box32b instanced
} 2>&1 | tee gpurun_out/ab11.log
FOUNDATION_PT_LIB=$PWD/ab_libs/box32b.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ab11_launches.csv python scripts/probe.py --scene terrain --rays 1024 --reps 1 > /dev/null 2>&1
python - <<'PY'
import csv,re,collections
lines=[l for l in open('gpurun_out/ab11_launches.csv') if l.startswith('"')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    k=re.sub(r"[(<].*","",row["Kernel Name"]).replace("void ",""); v=float(row["Metric Value"].replace(",",""))
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v/1e3
for k,a in sorted(agg.items(),key=lambda x:-x[1][1])[:16]: print(f"{k:24s} n={a[0]:4d} {a[1]:9.1f} us")
PY
