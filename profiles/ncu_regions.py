"""Per-region shares of k_gain from the source page of an `ncu --set full --import-source on` capture: the kernel's phases are found by
the marker comments of gain_kernels.cuh, so the table follows the file as it is at the time of the capture.
   python profiles/ncu_regions.py src.csv nbv_exploration_planner_b200/csrc/gain_kernels.cuh"""
import csv
import sys

src = open(sys.argv[2]).read().split("\n")


def line_of(marker, start=0):
    for i in range(start, len(src)):
        if marker in src[i]:
            return i + 1
    raise SystemExit(f"marker not found: {marker}")


k_gain = line_of("__global__ void __launch_bounds__(THREADS, 1024 / THREADS) k_gain")
marks = [("frames (build_yaw_plane / build_yaw_rays)", line_of("__device__ __noinline__ void build_yaw_plane") - 8),
         ("classifier frame (to_camera)", line_of("__device__ __forceinline__ void to_camera") - 1),
         ("DDA advance (all loops)", line_of("__device__ __forceinline__ void dda_advance") - 8),
         ("(other kernels)", line_of("enum { kDense = 0")),
         ("prologue: set-up, slot-grid probes, frames", k_gain),
         ("dense volume fill", line_of("// (c) dense status volume", k_gain)),
         ("after the fill", line_of("const int n_rays = s_umax * yc;", k_gain) - 1),
         ("ray set-up (incl. the look-up lambdas the loops call)", line_of("// ---- ray end point", k_gain)),
         ("general stretch", line_of("while (__any_sync(0xffffffffu, run))", k_gain)),
         ("top/bottom stretch", line_of("// ---- only the nearer of top/bottom is open", k_gain)),
         ("window stretch", line_of("// ---- window: every plane is sure", k_gain)),
         ("ray epilogue + CTA epilogue", line_of("const unsigned walked = S1 >> 16;", k_gain)),
         ("(rest of the file)", line_of("// Node scoring (rrt.cpp:220-246)", k_gain))]
rows = list(csv.reader(open(sys.argv[1])))
hdr = fpath = None
acc = {m[0]: [0, 0, 0] for m in marks}
acc["(exact planes, helpers above line %d)" % marks[0][1]] = [0, 0, 0]
acc["(other files: intrinsics)"] = [0, 0, 0]
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or fpath is None or not r[0].isdigit():
        continue
    k = len(r) - len(hdr)
    g = lambda n: r[hdr.index(n) + k]
    try:
        n, ins, smp, thr = int(r[0]), int(g("Instructions Executed")), int(g("# Samples")), int(g("Thread Instructions Executed"))
    except ValueError:
        continue
    if fpath != "gain_kernels.cuh":
        key = "(other files: intrinsics)"
    else:
        key = "(exact planes, helpers above line %d)" % marks[0][1]
        for name, lo in marks:
            if n >= lo:
                key = name
    a = acc[key]
    a[0] += ins
    a[1] += smp
    a[2] += thr
tot = sum(a[0] for a in acc.values()) or 1
samp = sum(a[1] for a in acc.values()) or 1
print(f"\nper region of k_gain (warp-instructions {tot}, stall samples {samp}):")
for k, (i, s, t) in acc.items():
    if i:
        print(f"  {100 * i / tot:5.1f}% inst {100 * s / samp:5.1f}% stall-samples  {t / i:5.1f} lanes/inst  {k}")
