"""profiles/traffic.json from an `ncu --set full` capture of the dominant kernel (roofline `traffic` of bench.py).

On the GPU box (one gpurun call; the plain run first, as the profiling recipe asks):

    python bench.py --no-cpu --no-other --steps 2 --warmup 3 > gpurun_out/plain.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:decode_small_fwd -s 3 -c 1 \
        -o gpurun_out/r02_fwd python bench.py --no-cpu --no-other --steps 2 --warmup 3

Here (no GPU needed):

    python tools/measure_traffic.py gpurun_out/r02_fwd.ncu-rep pos_K45_V20k profiles/r02_decode_small_fwd.txt

reads dram__bytes_read.sum + dram__bytes_write.sum of the captured launch, writes them per launch into
profiles/traffic.json under the workload's name together with where they came from, and writes the text summary
(tools/ncu_summary.py) next to it."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, workload = sys.argv[1], sys.argv[2]
summary = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(name):
    return hdr.index(name)


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


recs = []
for r in data:
    rd = to_bytes(r[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")])
    wr = to_bytes(r[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")])
    recs.append({"kernel": r[col("Kernel Name")], "dram_read": rd, "dram_write": wr,
                 "time_ms": float(r[col("gpu__time_duration.sum")].replace(",", "")) *
                            {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[col("gpu__time_duration.sum")], 1e-6)})
assert recs, "no launches in the report"
top = max(recs, key=lambda x: x["time_ms"])
path = os.path.join(ROOT, "profiles", "traffic.json")
try:
    tj = json.load(open(path))
except Exception:
    tj = {}
tj[workload] = int(top["dram_read"] + top["dram_write"])
tj["_source"] = (f"ncu --set full --clock-control none, {os.path.basename(rep)}: {top['kernel'][:60]} dram__bytes_read.sum "
                 f"{top['dram_read'] / 1e9:.3f} GB + dram__bytes_write.sum {top['dram_write'] / 1e9:.3f} GB per launch "
                 f"(tools/measure_traffic.py)")
tj.pop("_comment", None)
json.dump(tj, open(path, "w"), indent=1)
print(json.dumps(tj, indent=1))
if summary:
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, summary], check=True,
                   stdout=subprocess.DEVNULL)
