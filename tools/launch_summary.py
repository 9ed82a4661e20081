"""Per-kernel summary of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file LIST.csv ...):
    python tools/launch_summary.py LIST.csv ["header line" ...] > profiles/NAME_summary.txt"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"^void ", "", r[ki]); name = re.sub(r"^lv::", "", name); name = re.split(r"[<(]", name)[0]
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[ui], 1e-6)
    e = tot.setdefault(name, [0, 0.0]); e[0] += 1; e[1] += v
s = sum(e[1] for e in tot.values()) or 1.0
for h in sys.argv[2:]:
    print("# " + h)
print("%-28s %8s %14s %8s" % ("kernel", "launches", "total_ms", "share"))
for k, (n, ms) in tot.items():
    print("%-28s %8d %14.3f %7.1f%%" % (k, n, ms, 100.0 * ms / s))
