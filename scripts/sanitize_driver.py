#!/usr/bin/env python3
"""Small workload touching every kernel (for compute-sanitizer): site tables, rollout (initial + stagnating weights),
replay, statistics + winner record, location analysis, and a wide-map rollout."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from eirgrid_b200 import _abi, _lib, synthetic
ASSETS = os.path.join(ROOT, "tests", "golden", "ireland_map")
ctx = _lib.Context(0)
ctx.map_load_dir(ASSETS)
w = _lib.Weights()
res, traj, sites, yearly = ctx.rollout(w, 96, seed=3, want_sites=True, want_yearly=True)
w.update(res, traj)
t = w.table(); t.iterations_without_improvement = 900; w.set_table(t)
res, traj, _, _ = ctx.rollout(w, 96, seed=4)
rres, rsites, _ = ctx.replay(traj)
assert np.array_equal(rres["total_cost"], res["total_cost"])
stats = np.zeros(_abi.STATS_WORDS, np.int64); rec = np.zeros(16 + 64 + 1088, np.uint8)
cfg = _abi.RunCfg()
_lib.check(_lib.lib().eg_train_batch_begin(ctx.h, w.h, __import__("ctypes").byref(cfg), 5, 0, 96))
_lib.check(_lib.lib().eg_train_batch_end(ctx.h, _abi.ptr(stats), _abi.ptr(rec)))
ctx.location_analysis(True, first_point=0, n_points=200)
ctx.close()
ctx = _lib.Context(0)
sx, sy, spop, ex, ey, et, ec, cx, cy = synthetic.load_ireland_arrays(ASSETS)
ctx.map_set(sx[:40], sy[:40], spop[:40], ex[:12], ey[:12], et[:12], ec[:12], cx, cy, 101, 500.0)
res, traj, _, _ = ctx.rollout(_lib.Weights(), 24, seed=6)
ctx.close()
print("sanitize driver ok", int(res["n_generators"].sum()))
