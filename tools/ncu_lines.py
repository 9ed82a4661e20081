"""Per-source-line instruction counts from an ncu report (ncu -i REP --page source --print-source cuda,sass --csv):
    python tools/ncu_lines.py REP KERNEL_REGEX [top]
prints, per file, the lines with the most executed warp instructions (and their share, samples, avg active threads)."""
import csv, subprocess, sys, collections

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, per = None, None, collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if r[0] in ("Function Name", "Kernel Name") or hdr is None:
        continue
    if r[0].strip().isdigit():
        try:
            ie = hdr.index("Instructions Executed"); te = hdr.index("Thread Instructions Executed"); sm = hdr.index("# Samples")
            key = (fname, int(r[0]))
            e = per.setdefault(key, [0, 0, 0, r[1]])
            e[0] += int(r[ie] or 0); e[1] += int(r[te] or 0); e[2] += int(r[sm] or 0)
        except (ValueError, IndexError):
            pass
tot = sum(e[0] for e in per.values()) or 1
tots = sum(e[2] for e in per.values()) or 1
print("total warp instructions %d, samples %d" % (tot, tots))
byfile = collections.Counter()
for (f, l), e in per.items():
    byfile[f] += e[0]
for f, n in byfile.most_common():
    print("  %-20s %5.1f %%" % (f, 100.0 * n / tot))
for (f, l), e in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-16s %5d  inst %5.1f%%  samples %5.1f%%  lanes %4.1f  %s" % (f, l, 100.0 * e[0] / tot, 100.0 * e[2] / tots, e[1] / max(e[0], 1), e[3][:110]))
