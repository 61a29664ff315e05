"""Merges one `ncu --set full` capture of the ray-sweep kernel into profiles/ncu_traffic.json, keyed by the BASELINE
configuration it was taken on (bench.py reads roofline.traffic from there; a configuration without a capture reports null):
    python profiles/summarize_ncu.py gpurun_out/r2_cfg3_prof.ncu-rep cfg3 "2000 of 10000 candidates per launch"
dram bytes are per LAUNCH of the capture; `candidates_in_capture` lets bench.py scale them to its own launch size."""
import csv
import json
import os
import subprocess
import sys

rep = sys.argv[1]
config = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, zip(units, vals)))


def to_bytes(k):
    u, v = d[k]
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


out = {
    "source": os.path.basename(rep),
    "kernel": d["Kernel Name"][1] if "Kernel Name" in d else "k_gain",
    "k_gain_dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
    "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
    "gpu_time_duration_us_under_ncu": float(d["gpu__time_duration.sum"][1].replace(",", "")) * {"us": 1, "ms": 1e3, "ns": 1e-3}[d["gpu__time_duration.sum"][0]],
    "grid_size": d["launch__grid_size"][1], "registers_per_thread": d["launch__registers_per_thread"][1],
    "issue_active_pct": d["smsp__issue_active.avg.pct_of_peak_sustained_active"][1],
    "alu_pipe_pct": d["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"][1],
    "fma_pipe_pct": d["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"][1],
    "fp64_pipe_pct": d["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"][1],
    "warps_active_pct": d["sm__warps_active.avg.pct_of_peak_sustained_active"][1],
    "l1_hit_pct": d["l1tex__t_sector_hit_rate.pct"][1], "l2_hit_pct": d["lts__t_sector_hit_rate.pct"][1],
    "dram_throughput_pct": d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][1],
    "lts_throughput_pct": d["lts__throughput.avg.pct_of_peak_sustained_elapsed"][1],
    "warp_instructions": d["smsp__inst_executed.sum"][1],
    "threads_per_instruction": d["smsp__thread_inst_executed_per_inst_executed.ratio"][1],
}
out["note"] = note
out["block_size"] = d["launch__block_size"][1]
out["shared_mem_per_block_dynamic_kb"] = d["launch__shared_mem_per_block_dynamic"][1]
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
try:
    allrec = json.load(open(path))
    if "k_gain_dram_bytes_per_launch" in allrec:   # round-1 layout: a single record (cfg2)
        allrec = {"cfg2_round1": allrec}
except Exception:
    allrec = {}
allrec[config] = out
json.dump(allrec, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
