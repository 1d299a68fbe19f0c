#!/usr/bin/env python3
"""Distribution-level comparison of the weight update rules (north star, correctness point 2).

For several master seeds, train N iterations from fresh weights with
  seq16     reference rule: episodes sampled 16 at a time from one snapshot (16 stale rayon workers), every episode's record
            applied in order with the reference's per-episode arithmetic (eg_update) — `--update-mode sequential`
  batchB    batch-synchronous rule (device statistics + eg_update_apply_stats) with B episodes per snapshot
and report the final best score, the best run's metrics and the Spearman rank correlation of the learned weight tables
against seq16 of the same seed. Rollouts are the CUDA kernel in every arm (bit-identical to the CPU oracle).

    python scripts/learning_distribution.py [N=8192] [seeds=6] [out.json]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eirgrid_b200 import _lib, trainer as T  # noqa: E402

ASSETS = os.path.join(ROOT, "tests", "golden", "ireland_map")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
SEEDS = int(sys.argv[2]) if len(sys.argv) > 2 else 6
OUT = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "learning_distribution.json")


def score(m):
    import math
    net, opinion, cost = m[0], m[1], m[2]
    if net > 0:
        return 1.0 - min(net / 1e6, 1.0)
    nc = max(cost / 5e10, 1.0)
    cs = 1.0 - min(math.log(nc) / math.log(100.0), 1.0)
    cw = 0.8 if nc > 8 else 0.5
    return 1.0 + (cs * cw + opinion * (1.0 - cw))


def ranks(a):
    r = np.empty(a.size)
    r[np.argsort(a.ravel(), kind="stable")] = np.arange(a.size)
    return r


def spearman(a, b):
    ra, rb = ranks(a), ranks(b)
    return float(np.corrcoef(ra, rb)[0, 1])


def run_sequential(seed, chunk):
    ctx = _lib.Context(0)
    ctx.map_load_dir(ASSETS)
    w = _lib.Weights()
    done = 0
    while done < N:
        res, traj, _, _ = ctx.rollout(w, chunk, seed=seed, first_episode=done)
        w.update(res, traj, rng_seed=seed)
        done += chunk
    ctx.close()
    return w


def run_batch(seed, batch):
    tr = T.BatchTrainer(batch, seed=seed, device=0, asset_dir=ASSETS)
    for _ in range(N // batch):
        tr.step()
    w = tr.weights
    tr.close()
    return w


def summarize(w):
    t = w.table()
    arr = np.ctypeslib.as_array(t.weights).copy()
    m = list(t.best_metrics)
    return {"best_score": score(m), "net_emissions": m[0], "opinion": m[1], "cost": m[2], "iwi": int(t.iterations_without_improvement)}, arr


def main():
    arms = [("seq16", lambda s: run_sequential(s, 16)), ("batch16", lambda s: run_batch(s, 16)), ("batch256", lambda s: run_batch(s, 256)),
            ("batch4096", lambda s: run_batch(s, 4096))]
    out = {"iterations": N, "seeds": SEEDS, "arms": {}}
    tables = {}
    for name, fn in arms:
        rows = []
        for s in range(1, SEEDS + 1):
            summ, arr = summarize(fn(1000 + s))
            tables[(name, s)] = arr
            rows.append(summ)
        out["arms"][name] = rows
    for name, _ in arms:
        rows = out["arms"][name]
        bs = np.array([r["best_score"] for r in rows])
        rho = [spearman(tables[(name, s)], tables[("seq16", s)]) for s in range(1, SEEDS + 1)]
        # rank correlation between two seq16 runs of different seeds = the noise floor of that statistic
        out["arms"][name + "_summary"] = {"best_score_mean": float(bs.mean()), "best_score_std": float(bs.std()), "best_score_min": float(bs.min()),
                                          "best_score_max": float(bs.max()), "net_zero_runs": int(sum(r["net_emissions"] <= 0 for r in rows)),
                                          "spearman_vs_seq16_mean": float(np.mean(rho))}
    floor = [spearman(tables[("seq16", s)], tables[("seq16", s % SEEDS + 1)]) for s in range(1, SEEDS + 1)]
    out["spearman_seq16_between_seeds_mean"] = float(np.mean(floor))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump(out, open(OUT, "w"), indent=1)
    for k, v in out["arms"].items():
        if k.endswith("_summary"):
            print(k, json.dumps(v))
    print("spearman floor (seq16 across seeds)", out["spearman_seq16_between_seeds_mean"])


if __name__ == "__main__":
    main()
