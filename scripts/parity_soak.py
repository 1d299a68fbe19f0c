#!/usr/bin/env python3
"""Large-sample parity soak: CUDA rollouts (through the C ABI) against the CPU oracle, episode by episode.

Not part of the test suite (minutes of CPU time on the GPU box). For a series of weight tables — initial, trained by
sequential updates, and with the stagnation counter moved into each sampling regime (iwi 0 / 150 / 600 / 900 / 1300 /
3500) — it rolls out `n` episodes on the GPU and in the oracle's fast mode with the same (seed, episode id) streams and
compares action records, chosen sites, yearly metrics and final metrics byte for byte (score within 1e-12).
Writes one JSON line per table and a summary: usage `python scripts/parity_soak.py [n_per_table] [out.json]`.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import oracle_lib as O  # noqa: E402
from eirgrid_b200 import _abi, _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "parity_soak.json")
world = O.World.ireland(fast=True)
ctx = _lib.Context(0)
ctx.map_load(*[os.path.join(O.MAP_DIR, f) for f in ("settlements.json", "ireland_generators.csv", "coastline_points.json")])

ow, gw = O.Weights(), _lib.Weights()
rows, total, bad_total = [], 0, 0


def compare(tag, seed, cfg=None):
    global total, bad_total
    t0 = time.perf_counter()
    res, traj, sites, yearly = ctx.rollout(gw, n, seed=seed, cfg=cfg, want_sites=True, want_yearly=True)
    t1 = time.perf_counter()
    eres, etraj, esites, eyearly = world.rollout(ow, n, seed=seed, cfg=cfg)
    t2 = time.perf_counter()
    bad = np.zeros(n, bool)
    bad |= np.frombuffer(traj.tobytes(), np.uint8).reshape(n, -1).__ne__(np.frombuffer(etraj.tobytes(), np.uint8).reshape(n, -1)).any(1)
    bad |= np.frombuffer(sites.tobytes(), np.uint8).reshape(n, -1).__ne__(np.frombuffer(esites.tobytes(), np.uint8).reshape(n, -1)).any(1)
    for f in ("net_emissions", "public_opinion", "total_cost", "power_reliability", "n_generators", "n_offsets", "n_deficit_actions",
              "n_additional_actions", "flags"):
        bad |= res[f] != eres[f]
    bad |= np.abs(res["score"] - eres["score"]) > 1e-12 * np.abs(eres["score"])
    for f in yearly["y"].dtype.names:
        if f != "reserved":
            bad |= (yearly["y"][f] != eyearly["y"][f]).any(1)
    row = {"table": tag, "seed": seed, "episodes": n, "mismatching_episodes": int(bad.sum()), "first_bad": int(np.flatnonzero(bad)[0]) if bad.any() else None,
           "flagged": int((res["flags"] != 0).sum()), "mean_plants": float(res["n_generators"].mean()),
           "mean_actions": float((res["n_deficit_actions"].astype(float) + res["n_additional_actions"]).mean()),
           "gpu_s": t1 - t0, "oracle_s": t2 - t1}
    print(json.dumps(row), flush=True)
    rows.append(row)
    total += n
    bad_total += int(bad.sum())
    return eres, etraj


eres, etraj = compare("initial", 101)
# sequential reference-rule updates with the first 4096 episodes: a best strategy and a non-zero stagnation counter
ow.update(eres[:4096], etraj[:4096])
gw.update(eres[:4096], etraj[:4096])
assert bytes(ow.table()) == bytes(gw.table())
eres, etraj = compare("after 4096 sequential updates", 102)
ow.update(eres[:4096], etraj[:4096])
gw.update(eres[:4096], etraj[:4096])
for iwi in (0, 150, 600, 900, 1300, 3500):
    t = ow.table()
    t.iterations_without_improvement = iwi
    ow.set_table(t)
    gw.set_table(t)
    compare("trained, iwi=%d" % iwi, 200 + iwi)
compare("trained, iwi=3500, energy sales off", 300, cfg=_abi.RunCfg(enable_energy_sales=0))
compare("trained, iwi=3500, cost_only", 301, cfg=_abi.RunCfg(cost_only=1))
summary = {"episodes_compared": total, "mismatching_episodes": bad_total, "tables": rows}
os.makedirs(os.path.dirname(out_path), exist_ok=True)
json.dump(summary, open(out_path, "w"), indent=1)
print("TOTAL %d episodes, %d mismatching" % (total, bad_total))
sys.exit(1 if bad_total else 0)
