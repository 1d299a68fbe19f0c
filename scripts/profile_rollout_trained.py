#!/usr/bin/env python3
"""ncu driver: a few training steps, then rollouts from the trained (stagnation-regime) table."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eirgrid_b200 import trainer as T
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
tr = T.BatchTrainer(n, seed=20250101, device=0, asset_dir=os.path.join(ROOT, "tests", "golden", "ireland_map"))
for _ in range(5):
    tr.step()
tr.upload_weights()
for _ in range(2):
    tr.launch_rollout()
tr.stream.synchronize()
res, traj = tr.fetch_results()
print("ok", float(res["n_generators"].mean()), float((res["n_deficit_actions"].astype(float) + res["n_additional_actions"]).mean()))
