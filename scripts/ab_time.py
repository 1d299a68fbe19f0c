#!/usr/bin/env python3
"""Time the rollout kernel of the library selected by EIRGRID_LIB_NAME (CUDA events, device-resident)."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from eirgrid_b200 import _abi, _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
ctx = _lib.Context(0, stream.cuda_stream)
ctx.map_load_dir(os.path.join(ROOT, "tests", "golden", "ireland_map"))
w = _lib.Weights()
ctx.weights_upload(w)
d_res = torch.empty(n * 64, dtype=torch.uint8, device=dev)
d_traj = torch.empty(n * 1088, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
times = []
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx.rollout_device(n, 20250101, 0, d_res, d_traj)
    e1.record(stream)
    stream.synchronize()
    times.append(e0.elapsed_time(e1))
h = hashlib.sha1(d_res.cpu().numpy().tobytes() + d_traj.cpu().numpy().tobytes()).hexdigest()[:12]
times.sort()
print("%-20s n=%d  min %.3f ms  median %.3f ms  -> %.2f M episodes/s  sha %s" % (
    os.environ.get("EIRGRID_LIB_NAME", "default"), n, times[0], times[len(times) // 2], n / times[len(times) // 2] / 1e3, h))

# the same on the table the end-to-end steps train (3 batch-rule steps: stagnation counter far above 500)
from eirgrid_b200 import trainer as T  # noqa: E402
ctx.close()
tr = T.BatchTrainer(n, seed=20250101, device=0, asset_dir=os.path.join(ROOT, "tests", "golden", "ireland_map"))
for _ in range(3):
    tr.step()
tr.upload_weights()
times = []
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tr.stream)
    tr.launch_rollout(first_episode=10_000_000)
    e1.record(tr.stream)
    tr.stream.synchronize()
    times.append(e0.elapsed_time(e1))
res, traj = tr.fetch_results()
h = hashlib.sha1(res.tobytes() + traj.tobytes()).hexdigest()[:12]
times.sort()
print("%-20s trained table     min %.3f ms  median %.3f ms  -> %.2f M episodes/s  sha %s" % (
    os.environ.get("EIRGRID_LIB_NAME", "default"), times[0], times[len(times) // 2], n / times[len(times) // 2] / 1e3, h))
