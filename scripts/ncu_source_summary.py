"""`ncu -i report.ncu-rep --page source --csv` (SASS view of one kernel captured with --set full --import-source on) ->
text summary: warp-stall sampling by reason, executed warp instructions by opcode (optionally per unit of work), hottest
instructions.   usage: ncu_source_summary.py <source.csv> [units] [unit-name]"""
import collections
import csv
import sys

path = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
uname = sys.argv[3] if len(sys.argv) > 3 else "launch"
rows = list(csv.reader(open(path)))
kernel = rows[0][1] if rows and len(rows[0]) > 1 else "?"
hdr = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) >= len(hdr) and r[0].startswith("0x")]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
I = lambda r, h: int(r[ix[h]] or 0)
tot_s = sum(I(r, "# Samples") for r in data)
tot_e = sum(I(r, "Instructions Executed") for r in data)
print("kernel: %s" % kernel[:120])
print("SASS instructions %d, warp instructions executed %d (%.1f per %s), stall samples %d" % (len(data), tot_e, tot_e / units, uname, tot_s))
print("stall reasons: " + "  ".join("%s %.1f%%" % (s[6:], 100.0 * sum(I(r, s) for r in data) / max(tot_s, 1)) for s in sorted(stalls, key=lambda s: -sum(I(r, s) for r in data))[:10]))
c, cs = collections.Counter(), collections.Counter()
for r in data:
    src = r[ix["Source"]].split()
    op = src[1] if src[0].startswith("@") else src[0]
    c[op] += I(r, "Instructions Executed")
    cs[op] += I(r, "# Samples")
print("%-34s %14s %8s %9s" % ("opcode", "per " + uname, "share", "samples"))
for op, v in c.most_common(40):
    print("%-34s %14.1f %7.1f%% %8.1f%%" % (op, v / units, 100.0 * v / tot_e, 100.0 * cs[op] / max(tot_s, 1)))
print("hottest instructions (stall samples):")
for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:15]:
    top = sorted(((s[6:], I(r, s)) for s in stalls), key=lambda kv: -kv[1])[:2]
    print("  %6d  %-60s %s" % (I(r, "# Samples"), r[ix["Source"]].strip()[:60], top))
