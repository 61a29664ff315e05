"""Per-source-line instruction and stall-sample shares of one kernel from an `ncu --set full --import-source on` capture:
   ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > src.csv ; python profiles/ncu_lines.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lines = []
fpath, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] in ("Function Name",) or hdr is None:
        continue
    if r[0] != "" and r[0].isdigit():
        # the source text may itself contain quotes and commas (inline PTX): index the counters from the right
        k = len(r) - len(hdr)
        g = lambda name: r[hdr.index(name) + k]
        try:
            lines.append((fpath, int(r[0]), " ".join(r[1:2 + k]).strip(), int(g("Instructions Executed")), int(g("# Samples")),
                          float(g("Avg. Threads Executed") or 0)))
        except ValueError:
            pass
tot = sum(l[3] for l in lines)
samp = sum(l[4] for l in lines)
print(f"total warp instructions {tot}, samples {samp}")
for f, n, txt, ins, sm, thr in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100 * ins / tot:5.1f}% inst {100 * sm / max(samp, 1):5.1f}% stall  thr {thr:4.1f}  {f}:{n}  {txt[:110]}")
