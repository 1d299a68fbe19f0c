#!/usr/bin/env python3
"""BASELINE configs[2] as stated: 100,000-iteration training runs on one B200, 10 master seeds, episodes in flight per
weights snapshot in {16, 256, 4096, 65536}, under

  inorderB   the reference's rule — every episode's record applied in episode order (learning.rs:131-373, strategy.rs:19-258),
             on the GPU (eg_train_batch_inorder) — with B episodes sampled from one snapshot. inorder16 is the reference's own
             shape (16 rayon workers reading a stale snapshot, multi_simulation.rs:425-508).
  batchB     the batch-synchronous rule (device statistics + eg_update_apply_stats), B episodes per snapshot.

Reports per arm: final best score (mean, sd over seeds), the best run's metrics, Spearman rank correlation of the learned
26 x 61 weight table with inorder16 of the same seed, and — the yardstick for that — of inorder16 with inorder16 of the
other seeds. Rollouts are the CUDA kernel in every arm.

    python scripts/config3_distribution.py [N=100000] [seeds=10] [out.json]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eirgrid_b200 import trainer as T  # noqa: E402

ASSETS = os.path.join(ROOT, "tests", "golden", "ireland_map")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
SEEDS = int(sys.argv[2]) if len(sys.argv) > 2 else 10
OUT = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "config3_distribution.json")
BATCHES = (16, 256, 4096, 65536)


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
    return float(np.corrcoef(ranks(a), ranks(b))[0, 1])


def run(rule, batch, seed):
    tr = T.BatchTrainer(batch, seed=seed, device=0, asset_dir=ASSETS)
    done, improvements = 0, 0
    t0 = time.perf_counter()
    while done < N:
        n = min(batch, N - done)
        if rule == "inorder":
            st = tr.step_inorder(n, rng_seed=seed)
        else:
            tr.set_batch(n)
            st = tr.step()
        improvements += int(st.n_improvements)
        done += n
    dt = time.perf_counter() - t0
    t = tr.weights.table()
    m = list(t.best_metrics)
    out = {"best_score": score(m), "net_emissions": m[0], "opinion": m[1], "cost": m[2], "reliability": m[3],
           "iwi": int(t.iterations_without_improvement), "iterations": int(t.iteration_count), "improvements": improvements,
           "seconds": dt}
    arr = np.ctypeslib.as_array(t.weights).copy()
    tr.close()
    return out, arr


def main():
    seeds = [3001 + i for i in range(SEEDS)]
    arms = [("inorder", b) for b in BATCHES] + [("batch", b) for b in BATCHES]
    res = {"n_iterations": N, "seeds": seeds, "arms": {}}
    tables = {}
    for rule, b in arms:
        name = "%s%d" % (rule, b)
        rows = []
        for s in seeds:
            r, arr = run(rule, b, s)
            rows.append(r)
            tables[(name, s)] = arr
        sc = np.array([r["best_score"] for r in rows])
        res["arms"][name] = {"runs": rows, "best_score_mean": float(sc.mean()), "best_score_sd": float(sc.std(ddof=1)) if len(sc) > 1 else 0.0,
                             "best_score_min": float(sc.min()), "best_score_max": float(sc.max()),
                             "net_zero_and_reliable": int(sum(r["net_emissions"] <= 0 and r["reliability"] == 1.0 for r in rows)),
                             "seconds_mean": float(np.mean([r["seconds"] for r in rows])),
                             "iterations_per_s": float(N / np.mean([r["seconds"] for r in rows]))}
        print(name, "score %.4f +- %.4f  [%0.4f, %.4f]  %.2f s/run" % (sc.mean(), res["arms"][name]["best_score_sd"], sc.min(), sc.max(),
                                                                      res["arms"][name]["seconds_mean"]), flush=True)
    ref = "inorder16"
    for name in res["arms"]:
        res["arms"][name]["spearman_vs_inorder16_same_seed"] = float(np.mean([spearman(tables[(name, s)], tables[(ref, s)]) for s in seeds]))
    cross = [spearman(tables[(ref, a)], tables[(ref, b)]) for i, a in enumerate(seeds) for b in seeds[i + 1:]]
    res["spearman_inorder16_between_seeds"] = {"mean": float(np.mean(cross)), "min": float(np.min(cross)), "max": float(np.max(cross))} if cross else None
    # mean table per arm (the seed noise averages out): rank agreement of the arm's MEAN table with inorder16's mean table
    mean_ref = np.mean([tables[(ref, s)] for s in seeds], axis=0)
    for name in res["arms"]:
        res["arms"][name]["spearman_of_seed_mean_table_vs_inorder16"] = spearman(np.mean([tables[(name, s)] for s in seeds], axis=0), mean_ref)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump(res, open(OUT, "w"), indent=1)
    ref_mean, ref_sd = res["arms"][ref]["best_score_mean"], res["arms"][ref]["best_score_sd"]
    lines = ["| arm | final best score (mean ± sd, min..max) | Δ vs inorder16 (in sd of inorder16) | net-zero & reliable | Spearman vs inorder16 (same seed / seed-mean table) | iterations/s |",
             "|---|---|---|---|---|---|"]
    for name, a in res["arms"].items():
        lines.append("| %s | %.4f ± %.4f (%.4f..%.4f) | %+.4f (%+.1f) | %d/%d | %.2f / %.2f | %.0f |" % (
            name, a["best_score_mean"], a["best_score_sd"], a["best_score_min"], a["best_score_max"], a["best_score_mean"] - ref_mean,
            (a["best_score_mean"] - ref_mean) / max(ref_sd, 1e-12), a["net_zero_and_reliable"], len(seeds),
            a["spearman_vs_inorder16_same_seed"], a["spearman_of_seed_mean_table_vs_inorder16"], a["iterations_per_s"]))
    if res["spearman_inorder16_between_seeds"]:
        c = res["spearman_inorder16_between_seeds"]
        lines.append("")
        lines.append("inorder16 against inorder16 of another seed: Spearman %.2f (%.2f..%.2f) — the yardstick for the weight-table column." % (c["mean"], c["min"], c["max"]))
    open(os.path.splitext(OUT)[0] + ".md", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
