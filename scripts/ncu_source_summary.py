"""Groups an `ncu --page source --csv` dump of one kernel into runs of SASS instructions with (nearly) the same execution count and prints, per
run: instruction count, executions, share of issued warp instructions, live lanes, share of stall samples, first instruction (markdown table)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ix["Address"]].startswith("0x"):
        continue
    f = lambda k: float(r[ix[k]].replace(",", "") or 0)
    data.append((r[ix["Source"]].strip(), f("Instructions Executed"), f("Thread Instructions Executed"), f("# Samples")))
tot_i = sum(d[1] for d in data); tot_s = sum(d[3] for d in data)
print(f"Totals: {tot_i / 1e9:.2f} G warp instructions, {int(tot_s)} stall samples, {len(data)} SASS instructions\n")
print("| SASS lines | instr. | executions | share of issued instr. | live lanes | share of stall samples | first instruction |")
print("|---|---|---|---|---|---|---|")
i = 0
while i < len(data):
    j = i
    while j + 1 < len(data) and data[i][1] > 0 and abs(data[j + 1][1] - data[i][1]) <= 0.002 * data[i][1]:
        j += 1
    ins = sum(d[1] for d in data[i:j + 1]); thr = sum(d[2] for d in data[i:j + 1]); smp = sum(d[3] for d in data[i:j + 1])
    if ins / tot_i >= 0.003 or smp / tot_s >= 0.003:
        print(f"| {i + 1}-{j + 1} | {j - i + 1} | {data[i][1] / 1e6:.1f} M | {100 * ins / tot_i:.1f} % | {thr / ins if ins else 0:.1f} | {100 * smp / tot_s:.1f} % | `{data[i][0]}` |")
    i = j + 1
ops = {}
for d in data:
    op = d[0].split()[0] if not d[0].startswith("@") else d[0].split()[1]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + d[1]
print("\nIssued warp instructions by opcode: " + ", ".join(f"{k} {100 * v / tot_i:.1f} %" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
