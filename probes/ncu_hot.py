"""Hot spots of an `ncu --page source --csv` export: python probes/ncu_hot.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
sec = int(sys.argv[3]) if len(sys.argv) > 3 else len(starts) - 1     # which kernel section (default: the last)
h = rows[starts[sec]]
end = starts[sec + 1] - 1 if sec + 1 < len(starts) else len(rows)
data = [r for r in rows[starts[sec] + 1:end] if len(r) == len(h)]
si = h.index("Warp Stall Sampling (All Samples)")
val = lambda r: int(r[si] or 0)
tot = sum(val(r) for r in data)
print("total samples", tot, "instructions", len(data))
idx = sorted(range(len(data)), key=lambda i: -val(data[i]))[:top]
for i in sorted(idx):
    print(f"{i:6d} {data[i][1].strip()[:100]:100s} {val(data[i]):7d} {100.0 * val(data[i]) / tot:5.1f}%")
print("per 250-instruction region: start samples% DMMA LDG STG LDS")
for b in range(0, len(data), 250):
    seg = data[b:b + 250]
    print(f"{b:6d} {100.0 * sum(val(r) for r in seg) / tot:5.1f}% ", sum('DMMA' in r[1] for r in seg), sum('LDG' in r[1] for r in seg),
          sum('STG' in r[1] for r in seg), sum('LDS' in r[1] for r in seg))
