#!/usr/bin/env python3
"""Minimal driver for ncu on the synthetic 10x map (wide kernel instantiation)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eirgrid_b200 import _lib, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = _lib.Context(0)
ctx.map_set(*synthetic.scaled_map(os.path.join(ROOT, "tests", "golden", "ireland_map"), factor=10))
w = _lib.Weights()
for r in range(2):
    res, traj, _, _ = ctx.rollout(w, n, seed=20250101, first_episode=r * n)
print("ok", float(res["score"].mean()), int(res["n_generators"].sum()), int(res["flags"].max()))
