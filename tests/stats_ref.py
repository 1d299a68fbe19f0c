"""numpy restatement of the batch-synchronous update statistics, used to check
eirgrid_b200/csrc/stats.cu and to drive the CPU (gloo) test of the multi-rank combine step."""
import math

import numpy as np

from eirgrid_b200 import _abi

HEADER = 8
YEAR_STRIDE = 3 * _abi.N_ACTIONS + _abi.N_DEFICIT_KEYS
FIXED = 16777216.0
DEFICIT_KEY_TYPE = [8, 7, 12, 11, 9, 0, 1, 4, 10, 5, 2, 3, 13, 14]


def default_score(net, opinion, cost):
    if net > 0.0:
        return 1.0 - min(net / 1000000.0, 1.0)
    normalized = max(cost / 50000000000.0, 1.0)
    cost_score = 1.0 - min(math.log(normalized) / math.log(100.0), 1.0)
    cw = 0.8 if normalized > 8.0 else 0.5
    return 1.0 + (cost_score * cw + opinion * (1.0 - cw))


def deficit_key(code):
    if code == 60:
        return 14
    if code < 45 and code % 3 == 0 and code // 3 in DEFICIT_KEY_TYPE:
        return DEFICIT_KEY_TYPE.index(code // 3)
    return -1


def contrast_consts(table, best_score):
    iwi = table.iterations_without_improvement
    stag = 1.0 + 0.2 * (iwi / 10.0) ** 1.8
    alr = table.learning_rate * (1.0 + 0.1 * iwi)
    return dict(has_best=bool(table.has_best), force=iwi > 800, best_score=best_score,
                threshold=0.1 * max(math.exp(-iwi / 500.0), 0.00001 / 0.1), stagnation=stag, alr=alr)


def batch_stats(results, trajs, consts, best, best_deficit):
    """best / best_deficit: per-year lists of action codes of the snapshot's best strategy."""
    stats = np.zeros(_abi.STATS_WORDS, np.int64)
    scores = np.zeros(len(results))
    for e in range(len(results)):
        r, t = results[e], trajs[e]
        score = default_score(float(r["net_emissions"]), float(r["public_opinion"]), float(r["total_cost"]))
        scores[e] = score
        stats[0] += 1
        passed = False
        log_pen = log_mild = 0
        if consts["has_best"]:
            det = (consts["best_score"] - score) / consts["best_score"] if consts["best_score"] > 0 else 0.0
            passed = det > consts["threshold"] or consts["force"]
            if passed:
                if det < 0:
                    log_pen = log_mild = -(1 << 40)
                else:
                    combined = det ** 0.3 * consts["stagnation"]
                    log_pen = int(np.rint(math.log(1.0 / (1.0 + consts["alr"] * 1.5 * combined)) * FIXED))
                    log_mild = int(np.rint(math.log(1.0 / (1.0 + consts["alr"] * combined * 0.5)) * FIXED))
                stats[1] += 1
        rows = _abi.traj_rows(t)
        for y in range(26):
            base = HEADER + y * YEAR_STRIDE
            nd = len(rows[y][0])
            run = [int(a) for a in rows[y][0]] + [int(a) for a in rows[y][1]]
            cur = run + run[:nd]
            cb = list(best[y]) + list(best_deficit[y])
            for i, a in enumerate(cur):
                if i < len(run):
                    stats[base + 2 * 61 + a] += 1
                else:
                    k = deficit_key(a)
                    if k >= 0:
                        stats[base + 3 * 61 + k] += 1
                if not passed:
                    continue
                if a not in cb:
                    stats[base + a] += log_pen
                elif i < len(cb) and cb[i] != a:
                    stats[base + 61 + a] += log_mild
    return stats, scores
