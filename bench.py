#!/usr/bin/env python3
"""bench.py — episodes/s of the batched 2025-2050 rollout + scoring + weight-update step on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--episodes E] [--impl reference]

A step = one training batch: E episodes per GPU (default 65,536 in flight, BASELINE configs[2]) sampled from the
current action-weight table, simulated 2025-2050, scored, and summarised into the update statistics; for N > 1 the
statistics table is summed with one NCCL allreduce (weak scaling: E per GPU). Prints ONE JSON line on rank 0.

  value     episodes/s with everything resident in HBM (kernels + collective only), CUDA-event timed per step,
            L2 flushed between steps, max over ranks
  e2e       episodes/s through BatchTrainer.step(): weights uploaded from the host every step, statistics and the
            batch winner read back, host-side update applied (wall clock between device synchronisations)
  roofline  rollout kernel: algorithmic HBM bytes / measured kernel time vs MEASURED_PEAKS.json (the path is not
            HBM-bound, DESIGN.md §roofline says what binds instead)
  cpu_baseline  the CPU oracle (a port of the reference's Rust loop) on this box's host cores, bounded sample

`--impl reference` times that CPU port alone (rank 0 only) and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
ASSETS = os.path.join(ROOT, "tests", "golden", "ireland_map")
METRIC = "full 2025-2050 episodes simulated and scored per second"
UNIT = "episodes/s"
RESULT_BYTES, TRAJ_BYTES, POLICY_BYTES = 64, 1088, 38784


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the timed regions run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.stop = threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.02)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=3)

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_reference_rate(seconds_target, threads, literal=True, mode_fast=False):
    """episodes/s of the CPU oracle (port of the reference loop) on a bounded sample."""
    import oracle_lib as O
    w = O.World.ireland(fast=True)
    ow = O.Weights()
    mode = O.FAST if mode_fast else O.FAITHFUL
    n = max(threads * 2, 8)
    t0 = time.perf_counter()
    w.rollout(ow, n, seed=20250101, mode=mode, literal_scan=literal and not mode_fast, threads=threads, want_sites=False, want_yearly=False)
    dt = time.perf_counter() - t0
    rate = n / dt
    n2 = int(max(n, min(rate * seconds_target, 4_000_000)))
    t0 = time.perf_counter()
    w.rollout(ow, n2, seed=20250101, first_episode=n, mode=mode, literal_scan=literal and not mode_fast, threads=threads, want_sites=False,
              want_yearly=False)
    dt = time.perf_counter() - t0
    return n2 / dt, n2, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import oracle_lib as O
    threads = os.cpu_count() or 1
    w = O.World.ireland(fast=True)
    ow = O.Weights()
    # size one step so that the whole run (warm-up + K steps) is about a minute of CPU work, at most 2 s per step
    probe = max(threads, 4)
    t0 = time.perf_counter()
    w.rollout(ow, probe, seed=20250101, mode=O.FAITHFUL, literal_scan=True, threads=threads, want_sites=False, want_yearly=False)
    step_seconds = min(2.0, 60.0 / max(args.steps + args.warmup, 1))
    per_step = max(probe, int(probe / (time.perf_counter() - t0) * step_seconds))
    per_step = max(2 * threads, (per_step + threads - 1) // threads * threads)  # whole rounds of the thread pool
    first = probe
    for _ in range(args.warmup):
        w.rollout(ow, per_step, seed=20250101, first_episode=first, mode=O.FAITHFUL, literal_scan=True, threads=threads, want_sites=False, want_yearly=False)
        first += per_step
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res, traj, _, _ = w.rollout(ow, per_step, seed=20250101, first_episode=first, mode=O.FAITHFUL, literal_scan=True, threads=threads,
                                    want_sites=False, want_yearly=False)
        ow.update(res, traj)  # the write-lock section, sequential like the reference
        first += per_step
    dt = time.perf_counter() - t0
    value = args.steps * per_step / dt
    sample = "%d steps x %d episodes, literal 100x100 placement scan + sequential weight update, %d threads" % (args.steps, per_step, threads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "policy rollout + scoring + batch weight-update statistics, %d episodes in flight per GPU "
                                   "(BASELINE configs[2] batch shape), Irish map: 130 settlements, 59 existing plants, 2601 candidate sites" % args.episodes,
                       "map": "ireland", "episodes_per_step": per_step,
                       "reference_arm": "CPU oracle port of the reference loop (the Rust reference cannot be built here): literal 100x100 placement "
                                        "scan, per-evaluation opinion sums, sequential per-episode weight update; each step is a bounded sample of the workload"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--episodes", type=int, default=65536, help="episodes in flight per GPU per step")
    ap.add_argument("--impl", default="eirgrid_b200")
    ap.add_argument("--seed", type=int, default=20250101)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", choices=["ireland", "scaled10"], default="ireland",
                    help="ireland: shipped map, 65,536 episodes in flight per GPU (BASELINE configs[2] batch shape, the headline); "
                         "scaled10: synthetic 10x scaled grid of configs[3] (eirgrid_b200/synthetic.py)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    import torch
    import torch.distributed as dist
    from eirgrid_b200 import trainer as T

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    map_arrays = None
    if args.workload == "scaled10":
        from eirgrid_b200 import synthetic
        map_arrays = synthetic.scaled_map(ASSETS, factor=10)
    tr = T.BatchTrainer(args.episodes, seed=args.seed, device=local, asset_dir=ASSETS, map_arrays=map_arrays, distributed=world > 1)
    n_total = args.episodes * world
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=tr.device)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    W = max(args.warmup, 3)
    K = args.steps
    tr.upload_weights()  # value is measured on the initial action-weight table (first batch of a training run)
    for k in range(W):
        tr.launch_rollout(first_episode=(k * world + rank) * args.episodes)
        tr.launch_stats()
        tr.reduce_stats()
        tr.warm_exchange()  # first-use costs of the small torch ops / collectives of step(), weights untouched
    barrier()

    with ClockSampler(local) as clocks:
        # ---- value: device-resident step (rollout + statistics kernels [+ allreduce]), CUDA events per step -------
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
        launches0 = tr.ctx.kernel_launches()
        barrier()
        for k in range(K):
            with torch.cuda.stream(tr.stream):
                flush.zero_()
            ev[k][0].record(tr.stream)
            tr.launch_rollout(first_episode=((W + k) * world + rank) * args.episodes)
            ev[k][1].record(tr.stream)
            tr.launch_stats()
            tr.reduce_stats()
            ev[k][2].record(tr.stream)
        barrier()
        launches = tr.ctx.kernel_launches() - launches0
        step_ms = sum(ev[k][0].elapsed_time(ev[k][2]) for k in range(K))
        rollout_ms = sum(ev[k][0].elapsed_time(ev[k][1]) for k in range(K)) / K
        t = torch.tensor([step_ms], dtype=torch.float64, device=tr.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms = float(t.item())
        value = n_total * K / (step_ms / 1e3)

        # ---- e2e: the public training step, host weights in, statistics out, host update applied ----------------
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            tr.step()
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=tr.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        e2e = n_total * K / e2e_s

        # ---- e2e with every episode's result and action record copied to the host (SimulationResult per episode)
        res_h = torch.empty(args.episodes * RESULT_BYTES, dtype=torch.uint8).pin_memory()
        traj_h = torch.empty(args.episodes * TRAJ_BYTES, dtype=torch.uint8).pin_memory()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            tr.upload_weights()
            tr.launch_rollout()
            with torch.cuda.stream(tr.stream):
                res_h.copy_(tr.d_results, non_blocking=True)
                traj_h.copy_(tr.d_traj, non_blocking=True)
            tr.stream.synchronize()
        barrier()
        full_s = time.perf_counter() - t0

        # ---- the device-resident step again on the table the e2e steps have trained (what explains e2e vs value) ---------
        KT = min(K, 20)
        tr.upload_weights()
        evt = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(KT)]
        barrier()
        for k in range(KT):
            with torch.cuda.stream(tr.stream):
                flush.zero_()
            evt[k][0].record(tr.stream)
            tr.launch_rollout(first_episode=((W + K + k) * world + rank + 1000) * args.episodes)
            tr.launch_stats()
            tr.reduce_stats()
            evt[k][1].record(tr.stream)
        barrier()
        t = torch.tensor([sum(evt[k][0].elapsed_time(evt[k][1]) for k in range(KT))], dtype=torch.float64, device=tr.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        value_trained = n_total * KT / (float(t.item()) / 1e3)
    clock_summary = clocks.summary()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    info = tr.ctx.map_info()
    ns = info["grid_n"] ** 2
    # algorithmic HBM bytes of one rollout launch: every episode writes its result and action record
    # once; the static tables (walk lists 16 B + order 2 B per (class, year, site), per-site factors, plant terms, small
    # tables, stamp pattern, policy snapshot) are read once and then live in L2
    static_bytes = 7 * 26 * ns * (16 + 2) + ns * 16 + 26 * 15 * 3 * 26 * 16 + 40000 + 3200 + POLICY_BYTES
    algo_bytes = args.episodes * (RESULT_BYTES + TRAJ_BYTES) + static_bytes
    achieved = algo_bytes / (rollout_ms / 1e3) / 1e9
    traffic, issue = None, None
    tp = os.path.join(ROOT, "profiles", "rollout_traffic.json")
    if os.path.exists(tp):
        try:
            prof = json.load(open(tp))
            if args.workload != "ireland":
                raise KeyError("profile is for the shipped map")
            if prof.get("episodes_per_launch") == args.episodes:
                traffic = prof.get("dram_bytes_per_launch")
            # what binds instead of HBM: warp-instruction issue. Instructions per episode are a property of the code and
            # the workload (counted by ncu, profiles/), the rate is measured live here.
            sms, sched = torch.cuda.get_device_properties(local).multi_processor_count, 4
            ipe = prof["warp_instructions_per_episode"]
            peak_issue = sms * sched * (clock_summary.get("sm_mhz") or 1965.0) * 1e6
            issue = {"warp_instructions_per_episode": ipe, "achieved_ginst_s": ipe * args.episodes / (rollout_ms / 1e3) / 1e9,
                     "peak_ginst_s": peak_issue / 1e9, "frac": ipe * args.episodes / (rollout_ms / 1e3) / peak_issue,
                     "source": "profiles/rollout_traffic.json (ncu smsp__inst_executed.sum) x live kernel time"}
            # FP64 ceiling without FMA (the path is compiled --fmad=false), measured live; the kernel's share of it from ncu
            import ctypes
            peak64 = ctypes.c_double()
            if T._lib.lib().eg_microbench_fp64(local, ctypes.byref(peak64)) == 0:
                issue["fp64_peak_tflops_no_fma"] = peak64.value
                issue["fp64_pipe_active_pct"] = prof.get("fp64_pipe_pct")
        except Exception:
            traffic, issue = None, None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": step_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "policy rollout + scoring + batch weight-update statistics, %d episodes in flight per GPU "
                               "(BASELINE configs[2] batch shape), %s" % (args.episodes, "Irish map: 130 settlements, 59 existing plants, 2601 candidate sites"
                                                                          if args.workload == "ireland" else
                                                                          "synthetic 10x scaled grid (configs[3]): %d settlements, %d existing plants, %d candidate sites"
                                                                          % (info["n_settlements"], info["n_existing"], ns)),
                   "map": args.workload,
                   "episodes_per_gpu": args.episodes, "l2": "256 MiB buffer written between timed steps (L2 flush)",
                   "timing": "CUDA events on the launching stream per step, max over ranks",
                   "rollout_kernel_ms": rollout_ms,
                   "value_on_trained_table": value_trained},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": POLICY_BYTES, "d2h_bytes_per_step": tr.d2h_bytes_per_step,
                "what": "BatchTrainer.step(): weights H2D from the host, rollout + statistics + winner-record kernels, statistics and "
                        "winner record D2H (pinned), host-side weight update applied. The weights learn during these steps (stagnation-mode "
                        "tables sample ~20 % more actions and plants per episode), so the rollout itself is slower here than in `value`, "
                        "which is timed on the initial table",
                "with_all_results_to_host": {"value": args.episodes * world * K / full_s, "unit": UNIT,
                                             "d2h_bytes_per_step": args.episodes * (RESULT_BYTES + TRAJ_BYTES)}},
        "gpu_launches": int(launches),
        "clocks": clock_summary,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "eg_episode_kernel<REPLAY=false, WIDE, MODE=1> (rollout, lean training instantiation)",
                     "algorithmic_bytes_per_launch": algo_bytes, "issue": issue,
                     "note": "the path moves ~1.2 KB per episode and is bound by warp-instruction issue/latency, not HBM: "
                             "see roofline.issue, DESIGN.md §4.1 and profiles/"},
    }
    if world == 1 and not args.no_cpu_baseline and args.workload == "ireland":
        threads = os.cpu_count() or 1
        rate, n_cpu, dt = cpu_reference_rate(args.cpu_seconds, threads, literal=True)
        rate_fast, n_fast, dt_fast = cpu_reference_rate(3.0, threads, mode_fast=True)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d episodes in %.1f s, oracle in reference-cost mode (literal 100x100 placement scan, per-evaluation opinion sums), %d threads"
                                          % (n_cpu, dt, threads),
                                "fast_mode": {"value": rate_fast, "sample": "%d episodes in %.1f s with the exact table restructurings" % (n_fast, dt_fast)}}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
