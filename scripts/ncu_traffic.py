#!/usr/bin/env python3
"""Per-launch counters of the rollout kernel from an `ncu --set full` report -> profiles/rollout_traffic.json
(bench.py reads it for roofline.traffic and the issue-slot line).

    python scripts/ncu_traffic.py gpurun_out/rollout_X.ncu-rep EPISODES [out.json]
"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, episodes = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "rollout_traffic.json")
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
hdr, units, vals = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name)
    v, u = float(vals[i].replace(",", "")), units[i]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return v * scale
d = {
    "source": os.path.basename(rep), "kernel": vals[hdr.index("Kernel Name")], "episodes_per_launch": episodes,
    "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
    "dram_bytes_per_launch": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
    "warp_instructions_per_launch": get("smsp__inst_executed.sum"),
    "warp_instructions_per_episode": get("smsp__inst_executed.sum") / episodes,
    "duration_s_under_ncu": get("gpu__time_duration.sum"),
    "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": get("launch__registers_per_thread"),
    "l2_hit_pct": get("lts__t_sector_hit_rate.pct"), "l1_hit_pct": get("l1tex__t_sector_hit_rate.pct"),
    "fp64_pipe_pct": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    # the L1 / shared-memory data pipe (shared-memory and global wavefronts): the second resource the kernel runs close to
    "l1_data_pipe_pct": get("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
    "shared_wavefronts": get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "shared_bank_conflicts": get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    "resident_warps_per_sm": get("sm__warps_active.avg.per_cycle_active"),
}
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d))
