"""Per-kernel totals and shares from an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file ...).
usage: python profiles/launch_shares.py <launches.csv> "<command that was profiled>" """
import csv
import re
import statistics
import sys

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"')) if len(r) > 14 and r[0] != "ID"]
launch = [(re.sub(r"\(.*", "", r[4]).replace("nbvgpu::", "").replace("void ", ""), int(r[8].strip("()").split(",")[0]), float(r[14].replace(",", "")) / 1e3) for r in rows]
tot = {}
for name, grid, us in launch:
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += us
all_us = sum(t[1] for t in tot.values())
print(f"launch list: ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 {cmd}")
print(f"{len(launch)} launches captured, {all_us / 1e3:.2f} ms of device time (cold-cache, serialised: shares, not absolutes)")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:90]:45s} n={n:4d}  total {us / 1e3:9.3f} ms  avg {us / n:9.1f} us  share {100 * us / all_us:5.1f}%")
# full-size planning steps: k_gain launches with the full grid (the largest seen), with the kernels that follow each one
big = max((g for n, g, _ in launch if n.startswith("k_gain")), default=0)
steps, cur = [], None
for name, grid, us in launch:
    if name.startswith("k_gain") and grid == big:
        cur = {"k_gain": us}
        steps.append(cur)
    elif cur is not None and name.split("<")[0] in ("k_best_yaw", "k_best_yaw_score", "k_score_best", "k_check_segments", "k_check_segments_warp"):
        cur[name.split("<")[0]] = cur.get(name.split("<")[0], 0.0) + us
if steps:
    keys = sorted({k for s in steps for k in s}, key=lambda k: -statistics.median(s.get(k, 0.0) for s in steps))
    med = {k: statistics.median(s.get(k, 0.0) for s in steps) for k in keys}
    print(f"\nwithin one full-size planning step ({len(steps)} steps in the list, medians, cold-cache under ncu):")
    print("  " + ", ".join(f"{k} {med[k]:.1f} us" for k in keys) + f" -> k_gain share {100 * med['k_gain'] / sum(med.values()):.1f}% of launched device time in the step")
