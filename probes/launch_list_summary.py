"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
`bench.py --steps 1 --warmup 1 --skip parity,e2e,other --no-cpu` (3 passes of the path: warm-up, timed step, phase-timing
step): per-kernel launch counts and times, the update kernel's share of a step and its DRAM bytes per launch.

    python probes/launch_list_summary.py profiles/r02_launches_bench_8Mx512.csv [rows cols] > profiles/r02_update_traffic.json
"""
import collections, csv, json, re, sys

path = sys.argv[1]
m = int(sys.argv[2]) if len(sys.argv) > 2 else 8_000_000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 512
PASSES = 3
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ix = {h: i for i, h in enumerate(rows[hi])}
SCALE = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ms = collections.defaultdict(float); cnt = collections.Counter(); dram = collections.defaultdict(float)
for r in rows[hi + 1:]:
    if len(r) < len(ix):
        continue
    k = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[ix["Kernel Name"]]))
    k = re.sub(r"<.*", "", k)
    v = float(r[ix["Metric Value"]].replace(",", ""))
    name, unit = r[ix["Metric Name"]], r[ix["Metric Unit"]]
    if name == "gpu__time_duration.sum":
        ms[k] += v * SCALE.get(unit, 1e-6); cnt[k] += 1
    elif name.startswith("dram__bytes"):
        dram[k] += v * BYTES.get(unit, 1.0)
upd = "pl::caqr_update2_kernel"
total = sum(v for k, v in ms.items() if k.startswith("pl::"))    # the path's own kernels; torch kernels in the list generate the input and run the checks, outside the timed region
npad = -(-n // 32) * 32
K = npad // 32
# algorithmic bytes of one pass of the update launches: C read + written once (16 B per entry), V read once (8 B)
alg = 0.0
for p in range(K):
    rows_p = m - 32 * p
    alg += 16.0 * rows_p * (npad - 32 * (p + 1)) + 8.0 * rows_p * 32      # factorisation: trailing columns
    alg += 16.0 * rows_p * (npad - 32 * p) + 8.0 * rows_p * 32            # form-Q: columns p..K-1
out = {
    "dram_bytes_per_launch": dram[upd] / max(cnt[upd], 1),
    "algorithmic_bytes_per_launch": alg / (cnt[upd] / PASSES) if cnt[upd] else None,
    "launches_per_step": cnt[upd] // PASSES,
    "share_of_step_ncu": ms[upd] / total if total else None,
    "ncu_ms_per_step_update": ms[upd] / PASSES,
    "ncu_ms_per_step_all": total / PASSES,
    "kernels": {k: {"launches_per_step": cnt[k] / PASSES, "ms_per_step": ms[k] / PASSES} for k in sorted(ms, key=lambda k: -ms[k]) if k.startswith("pl::")},
    "note": "dram__bytes_read.sum + dram__bytes_write.sum averaged over the caqr_update2_kernel launches (3 passes of the path: warm-up, "
            "timed step, phase-timing step) of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
            "--clock-control none python bench.py --steps 1 --warmup 1 --skip parity,e2e,other --no-cpu` (" + path + "); algorithmic = "
            "16 rows cols + 8 rows 32 bytes per launch (C read and written once, V read once); the excess is V tiles re-read from "
            "DRAM by the column-chunk CTAs of a strip",
}
out["ratio"] = out["dram_bytes_per_launch"] / out["algorithmic_bytes_per_launch"] if out["algorithmic_bytes_per_launch"] else None
print(json.dumps(out, indent=1))
