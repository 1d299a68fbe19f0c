"""Wall time of the in-order update: eg_update (host) vs eg_update_device (GPU) on the same 65,536 records, per regime."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eirgrid_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ctx = _lib.Context(0)
ctx.map_load_dir(os.path.join(ROOT, "tests", "golden", "ireland_map"))
dev = torch.device("cuda", 0)
for iwi in (0, 600, 1000, 1500):
    w = _lib.Weights()
    res, traj, _, _ = ctx.rollout(w, 256, seed=1)
    w.update(res, traj)
    t = w.table()
    t.iterations_without_improvement = iwi
    w.set_table(t)
    res, traj, _, _ = ctx.rollout(w, n, seed=2, first_episode=1000)
    d_res = torch.from_numpy(res.view(np.uint8).reshape(-1).copy()).to(dev)
    d_traj = torch.from_numpy(traj.view(np.uint8).reshape(-1).copy()).to(dev)
    torch.cuda.synchronize()
    hw = w.clone()
    t0 = time.perf_counter()
    hst = hw.update(res, traj, rng_seed=3)
    t_host = time.perf_counter() - t0
    times = []
    for rep in range(4):
        dw = w.clone()
        t0 = time.perf_counter()
        dst = ctx.update_device(dw, n, d_res, d_traj, rng_seed=3)
        times.append(time.perf_counter() - t0)
    same = bytes(hw.table()) == bytes(dw.table())
    print("iwi0=%4d n=%d: host %.1f ms (%.2f us/episode), device %.3f ms (%.1f ns/episode, first call %.3f ms), "
          "improvements %d, applied %d, identical table: %s" % (
              iwi, n, t_host * 1e3, t_host / n * 1e6, min(times[1:]) * 1e3, min(times[1:]) / n * 1e9, times[0] * 1e3,
              dst.n_improvements, dst.n_contrast_applied, same), flush=True)
ctx.close()
