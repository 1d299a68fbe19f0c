#!/usr/bin/env python3
"""Placement-walk statistics from a library built with -DEG_WALK_STATS (EIRGRID_LIB_NAME=libeg_walkstats.so)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from eirgrid_b200 import _lib
ctx = _lib.Context(0)
ctx.map_load_dir(os.path.join(ROOT, "tests", "golden", "ireland_map"))
res, traj, _, _ = ctx.rollout(_lib.Weights(), 16384, seed=20250101)
raw = res.view(np.uint8).reshape(len(res), -1)[:, 56:64].copy().view(np.uint64).ravel()
steps, evals, cands, pairs = raw & 0xFFFF, (raw >> 16) & 0xFFFF, (raw >> 32) & 0xFFFF, ((raw >> 48) & 0xFFFF) * 16
g = res["n_generators"].astype(float)
print("plants/episode %.1f | walk steps/placement %.2f | evaluation steps/placement %.2f | candidate lanes/evaluation step %.2f | plant iterations/episode %.0f"
      % (g.mean(), (steps / g).mean(), (evals / g).mean(), (cands / np.maximum(evals, 1)).mean(), pairs.mean()))
