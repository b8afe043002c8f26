"""Per source line (in line order) warp-instructions per unit of work from an
`ncu --page source --csv --print-source cuda,sass` dump.
usage: ncu_lines_by_line.py <csv> <units> [min_per_unit] [file-substring]"""
import csv
import sys

path, units = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
only = sys.argv[4] if len(sys.argv) > 4 else ""
rows = list(csv.reader(open(path)))
hdr, fname, out = None, "", []
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and r and r[0].isdigit():
        try:
            out.append((fname, int(r[0]), int(r[hdr.index("Instructions Executed")]),
                        int(r[hdr.index("# Samples")]), r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(o[2] for o in out)
ts = sum(o[3] for o in out)
print("total per unit %.1f" % (tot / units))
byfile = {}
for f, l, i, s, src in out:
    byfile[f] = byfile.get(f, 0) + i
print({k: round(v / units, 1) for k, v in byfile.items()})
for f, l, i, s, src in sorted(out, key=lambda o: (o[0], o[1])):
    if only in f and i / units >= thr:
        print(f"{f}:{l} {i / units:6.1f} samp {100 * s / max(ts, 1):4.1f}%  {src}")
