#!/usr/bin/env python3
"""Driver for ncu on the table the end-to-end steps train: K BatchTrainer steps (batch update rule, stagnation counter far
above 500), then rollout launches on that table. `ncu -k regex:eg_episode_kernel -s K -c 1` captures the first of them."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eirgrid_b200 import trainer as T  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tr = T.BatchTrainer(n, seed=20250101, device=0, asset_dir=os.path.join(ROOT, "tests", "golden", "ireland_map"))
for _ in range(k):
    st = tr.step()
tr.upload_weights()
for r in range(2):
    tr.launch_rollout()
tr.stream.synchronize()
res, traj = tr.fetch_results()
print("ok iwi", st.iterations_without_improvement, "mean plants", float(res["n_generators"].mean()),
      "actions", float((res["n_deficit_actions"].astype(float) + res["n_additional_actions"]).mean()))
