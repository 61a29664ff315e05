"""Print the figures of one or more bench.py JSON lines side by side: python scripts/show_bench.py gpurun_out/x.json ..."""
import json
import sys


def show(tag, r):
    c = r.get("config", {})
    ro = r.get("roofline") or {}
    print(f"{tag:28s} value {r['value']:>10.0f} e2e {r['e2e']['value']:>10.0f} ms/step {r['ms_per_step']:8.3f} launches {r.get('gpu_launches')}"
          f" roof {ro.get('frac') and round(ro['frac'], 4)} kern_ms {ro.get('avg_launch_ms') and round(ro['avg_launch_ms'], 4)}"
          f" step-kern_us {ro.get('step_minus_kernel_us') and round(ro['step_minus_kernel_us'], 1)} replays {ro.get('graph_replays')}"
          f" repl {(c.get('map_replication') or {}).get('ms')} upd_ms {c.get('map_update_ms_per_step')} lat_us {r.get('single_candidate_latency_us')}")


for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    show(f.split("/")[-1], d)
    for k, v in d.get("extra", {}).items():
        show("   " + k, v)
    if "cpu_baseline" in d:
        print("   cpu_baseline", round(d["cpu_baseline"]["value"], 1), d["cpu_baseline"]["sample"])
