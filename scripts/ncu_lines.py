"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump: warp
instructions executed and stall samples per CUDA source line, top N."""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = None
out = []
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and r and r[0].isdigit():
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        try:
            out.append((int(r[i_inst]), int(r[i_samp]), fname, int(r[0]), r[1].strip()[:90]))
        except ValueError:
            pass
tot = sum(o[0] for o in out)
tots = sum(o[1] for o in out)
print(f"total warp-instructions {tot}, samples {tots}")
for inst, samp, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{inst:10d} {100*inst/tot:5.1f}%  samp {100*samp/max(tots,1):5.1f}%  {f}:{ln}  {src}")
