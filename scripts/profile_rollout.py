#!/usr/bin/env python3
"""Minimal driver for ncu: a few rollout launches of N episodes with the initial weights (no torch needed)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eirgrid_b200 import _abi, _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = _lib.Context(0)
ctx.map_load_dir(os.path.join(ROOT, "tests", "golden", "ireland_map"))
w = _lib.Weights()
for r in range(reps):
    res, traj, _, _ = ctx.rollout(w, n, seed=20250101, first_episode=r * n)
print("ok", float(res["score"].mean()), int(res["n_generators"].sum()))
