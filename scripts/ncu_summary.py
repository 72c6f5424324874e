"""Summaries for profiles/: (1) `ncu --set full --csv --page raw` rows -> per kernel variant: launches, time, tensor-pipe
active, issue active, DRAM bytes, L2 hit rate, top stall reasons;  (2) a `--metrics gpu__time_duration.sum` launch list ->
share of the step per kernel family.   usage: ncu_summary.py full <raw.csv> | launches <launches.csv> [skip]"""
import collections
import csv
import re
import sys


def read(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    return rows[0], rows[1:]


def short(name):
    name = re.sub(r"\(.*", "", name).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    return name


def full(path):
    hdr, rows = read(path)
    col = {h: i for i, h in enumerate(hdr)}
    data = rows[1:]

    def g(r, h):
        try:
            return float(r[col[h]].replace(",", ""))
        except (KeyError, ValueError):
            return float("nan")

    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    agg = collections.OrderedDict()
    for r in data:
        key = (short(r[col["Kernel Name"]]), r[col["Block Size"]], r[col["Grid Size"]])
        t = g(r, "gpu__time_duration.sum")
        a = agg.setdefault(key, collections.defaultdict(float))
        a["n"] += 1
        a["t"] += t
        for name, h in (("tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), ("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        ("l2hit", "lts__t_sector_hit_rate.pct"), ("warps", "sm__warps_active.avg.pct_of_peak_sustained_active")):
            a[name] += g(r, h) * t
        a["dr"] += g(r, "dram__bytes_read.sum")
        a["dw"] += g(r, "dram__bytes_write.sum")
        a["regs"] = g(r, "launch__registers_per_thread")
        for h in stalls:
            a[h] += g(r, h) * t
    tot = sum(a["t"] for a in agg.values())
    print("%-28s %-12s %-12s %3s %9s %6s %7s %6s %6s %5s %9s %9s  top stalls (warps stalled per issue-active cycle)" %
          ("kernel", "block", "grid", "n", "time us", "share", "tensor%", "issue%", "L2hit%", "regs", "dramR MB", "dramW MB"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        top = sorted(((a[h] / a["t"], h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stalls), reverse=True)[:4]
        print("%-28s %-12s %-12s %3d %9.1f %5.1f%% %7.1f %6.1f %6.1f %5d %9.1f %9.1f  %s" %
              (k[0][:28], k[1], k[2], a["n"], a["t"] / 1e3, 100 * a["t"] / tot, a["tensor"] / a["t"], a["issue"] / a["t"], a["l2hit"] / a["t"],
               a["regs"], a["dr"] / 1e6, a["dw"] / 1e6, ", ".join("%s %.2f" % (n, v) for v, n in top)))


def launches(path, skip=0):
    hdr, rows = read(path)
    col = {h: i for i, h in enumerate(hdr)}
    fam = collections.defaultdict(lambda: [0, 0.0])
    seen = 0
    for r in rows[1:]:
        if r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        seen += 1
        if seen <= skip:
            continue
        name = short(r[col["Kernel Name"]])
        name = re.sub(r"<.*", "", name)
        v = float(r[col["Metric Value"]].replace(",", ""))
        scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(r[col["Metric Unit"]], 1e-3)
        fam[name][0] += 1
        fam[name][1] += v * scale
    tot = sum(v[1] for v in fam.values())
    print("launches %d (skipped %d), kernel time %.2f ms" % (seen - skip, skip, tot / 1e3))
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print("%-40s n=%4d %9.1f us %5.1f%%" % (k[:40], v[0], v[1], 100 * v[1] / tot))


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2])
    else:
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
