#!/usr/bin/env python3
"""Summarise an .ncu-rep of the episode kernel: headline metrics + warp-instruction share per source line.

    python scripts/ncu_summary.py gpurun_out/rollout_X.ncu-rep [--lib eirgrid_b200/libeirgrid_b200.so] [--top 30]

SASS samples are mapped to episode.cu lines through the cubin's line table (nvdisasm -g), which needs -lineinfo.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads",
        "launch__shared_mem_per_block_dynamic", "sm__maximum_warps_per_active_cycle_pct", "launch__occupancy_limit_warps",
        # throughput of the units the path actually leans on, as % of peak
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum.per_second", "l1tex__t_bytes.sum.per_second",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "derived__memory_l1_wavefronts_shared_excessive",
        "sm__warps_active.avg.per_cycle_active", "achieved_occupancy", "sm__cycles_active.avg"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    lib = os.path.join(ROOT, "eirgrid_b200", "libeirgrid_b200.so")
    top = 30
    if "--lib" in sys.argv:
        lib = sys.argv[sys.argv.index("--lib") + 1]
    if "--top" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1])
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for h, u, v in zip(hdr, units, vals):
        if h in WANT or any(h.startswith(w) for w in ("smsp__average_warps_issue_stalled", "smsp__average_warp_latency")) \
                or ("warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")):
            print("  %-75s %-12s %s" % (h, u, v))
    # stall reasons (pct of samples)
    stalls = [(h, float(v)) for h, v in zip(hdr, vals) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and v]
    tot = sum(v for _, v in stalls) or 1
    print("stall reasons (pc samples):")
    for h, v in sorted(stalls, key=lambda x: -x[1])[:10]:
        print("  %-45s %5.1f %%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot))
    # source attribution
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    shdr = src[1]
    ci, cs, ct = shdr.index("Instructions Executed"), shdr.index("# Samples"), shdr.index("Thread Instructions Executed")
    kname = src[0][1]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith("episode")][0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    # pick the function whose mangled name matches the kernel's template arguments
    m_args = re.search(r"eg_episode_kernel<\(bool\)(\d), \(int\)(\d), \(int\)(\d)>|eg_episode_kernel<(\d), (\d), (\d)>", kname)
    a_ = [g for g in m_args.groups() if g is not None] if m_args else ["0", "0", "1"]
    want_replay, want_wide = a_[0] == "1", a_[1] == "1"
    fn_ok, line, m = False, None, {}
    for l in sass.split("\n"):
        s = l.strip()
        if s.startswith(".text."):
            fn_ok = ("eg_episode_kernelILb%dELi%sELi%sEE" % (want_replay, a_[1], a_[2])) in s
        mm = re.match(r'//## File "(.*)", line (\d+)', s)
        if mm:
            line = int(mm.group(2)) if mm.group(1).endswith("episode.cu") else -1
            continue
        mm = re.match(r"/\*([0-9a-f]{4,})\*/", s)
        if mm and fn_ok:
            m[int(mm.group(1), 16)] = line
    base = int(src[2][0], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for r in src[2:]:
        a = agg[m.get(int(r[0], 16) - base)]
        a[0] += int(r[ci] or 0)
        a[1] += int(r[cs] or 0)
        a[2] += int(r[ct] or 0)
    tot_i = sum(a[0] for a in agg.values()) or 1
    tot_s = sum(a[1] for a in agg.values()) or 1
    text = open(os.path.join(ROOT, "eirgrid_b200", "csrc", "episode.cu")).read().split("\n")
    print("source lines by warp instructions executed (total %d, samples %d):" % (tot_i, tot_s))
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        t = text[ln - 1].strip()[:95] if ln and ln > 0 else str(ln)
        print("  %5s %5.1f%% inst %5.1f%% samples  lanes %4.1f | %s" % (ln, 100 * a[0] / tot_i, 100 * a[1] / tot_s, a[2] / max(a[0], 1), t))


if __name__ == "__main__":
    main()
