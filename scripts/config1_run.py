#!/usr/bin/env python3
"""BASELINE configs[0]: the reference's default run — 1,000 iterations on the shipped map, 2025-2050, last 10 % replay the best
strategy — once on the CPU (oracle in reference-cost mode, all host threads, 16 episodes per snapshot like 16 rayon workers,
sequential per-episode update) and once on one B200 (same rule through `python -m eirgrid_b200 --update-mode sequential`
semantics: batch 16, eg_update). Same master seed: the GPU rollouts are bit-identical to the oracle's, so both runs take the
same decisions and end with the same weights; the point is the wall time.

    python scripts/config1_run.py [iterations=1000]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from eirgrid_b200 import _abi, _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
SEED, CHUNK = 20250101, 16
threads = os.cpu_count() or 1
final_full = N * 10 // 100

def run(rollout, weights, update):
    done, t0 = 0, time.perf_counter()
    while done < N:
        n = min(CHUNK, N - done)
        replay = done >= N - final_full and bool(weights.best()[0])
        cfg = _abi.RunCfg(replay_best=int(replay))
        res, traj = rollout(weights, n, done, cfg)
        update(weights, res, traj, replay)
        done += n
    return time.perf_counter() - t0

world = O.World.ireland(fast=True)
ow = O.Weights()
t_cpu = run(lambda w, n, first, cfg: world.rollout(w, n, seed=SEED, first_episode=first, cfg=cfg, mode=O.FAITHFUL, literal_scan=True,
                                                   threads=threads, want_sites=False, want_yearly=False)[:2],
            ow, lambda w, r, t, rp: w.update(r, t, replay=rp))
ctx = _lib.Context(0)
ctx.map_load_dir(os.path.join(ROOT, "tests", "golden", "ireland_map"))
gw = _lib.Weights()
t_gpu = run(lambda w, n, first, cfg: ctx.rollout(w, n, seed=SEED, first_episode=first, cfg=cfg)[:2], gw,
            lambda w, r, t, rp: w.update(r, t, replay_best=rp, rng_seed=0))
same = bytes(ow.table()) == bytes(gw.table())
t = gw.table()
out = {"iterations": N, "cpu": {"wall_s": t_cpu, "episodes_per_s": N / t_cpu, "threads": threads, "what": "oracle, reference-cost mode"},
       "gpu": {"wall_s": t_gpu, "episodes_per_s": N / t_gpu, "what": "1 B200, 16 episodes per launch (launch + copy latency bound)"},
       "identical_final_weights": same, "best_metrics": list(t.best_metrics), "iterations_without_improvement": int(t.iterations_without_improvement)}
print(json.dumps(out))
