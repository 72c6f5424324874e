"""ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum) of scripts/prof_point_ops.py + its
--order-out file -> entries of profiles/ncu_traffic.json (DRAM bytes per launch per op; ops with several launches are summed)."""
import csv
import json
import sys

csv_path, order_path, out_path, capture = sys.argv[1:5]
order = json.load(open(order_path))
rows = {}
with open(csv_path) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    kid = int(r["ID"])
    d = rows.setdefault(kid, {"kernel": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1)
    d[r["Metric Name"]] = v * scale
ids = sorted(rows)
assert len(ids) == len(order["order"]), (len(ids), len(order["order"]))
table = json.load(open(out_path))
acc = {}
for kid, name in zip(ids, order["order"]):
    a = acc.setdefault(name, {"batch": order["batch"], "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "gpu_time_ms": 0.0, "kernels": [], "capture": capture})
    a["dram_read_bytes"] += rows[kid].get("dram__bytes_read.sum", 0.0)
    a["dram_write_bytes"] += rows[kid].get("dram__bytes_write.sum", 0.0)
    a["gpu_time_ms"] += rows[kid].get("gpu__time_duration.sum", 0.0)
    a["kernels"].append(rows[kid]["kernel"].split("(")[0])
table.update(acc)
json.dump(table, open(out_path, "w"), indent=1)
for k, v in acc.items():
    print("%-45s read %8.1f MB  write %8.1f MB  %7.3f ms" % (k, v["dram_read_bytes"] / 1e6, v["dram_write_bytes"] / 1e6, v["gpu_time_ms"]))
