"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`).

    python tools/ncu_launch_summary.py gpurun_out/r02_launches_bench_pos.csv profiles/r02_launches_bench_pos_summary.txt "<command line>"
"""
import csv
import sys
from collections import defaultdict

src, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    tot[r[ik]] += v
    cnt[r[ik]] += 1
total = sum(tot.values())
with open(out, "w") as f:
    f.write(f"# {note}\n# ncu --metrics gpu__time_duration.sum --clock-control none: kernels are serialised and cold-cache, compare SHARES\n")
    f.write(f"{'kernel':72s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write(f"{k[:70]:72s} {cnt[k]:8d} {v:12.1f} {v / cnt[k]:10.1f} {100 * v / total:6.1f}%\n")
print(open(out).read())
