"""torchrun check of BatchTrainer.step_inorder_sharded: N ranks roll out shards of each batch, all-gather the records and apply the
reference's in-order update replicated — the weights must equal, bit for bit, those of ONE GPU running the same batches alone.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/inorder_sharded_check.py
"""
import hashlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eirgrid_b200 import trainer as T  # noqa: E402

ASSETS = os.path.join(ROOT, "tests", "golden", "ireland_map")
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
totals = [4096, 4097, 1001, 8192]  # even and ragged splits
tr = T.BatchTrainer(max(totals), seed=31, device=local, asset_dir=ASSETS, distributed=True)
for t in totals:
    st = tr.step_inorder_sharded(t, rng_seed=5)
sha = hashlib.sha256(bytes(tr.weights.table())).hexdigest()[:16]
shas = [None] * world
dist.all_gather_object(shas, sha)
assert len(set(shas)) == 1, shas
if rank == 0:
    solo = T.BatchTrainer(max(totals), seed=31, device=local, asset_dir=ASSETS, distributed=False)
    for t in totals:
        solo.step_inorder(t, rng_seed=5)
    solo_sha = hashlib.sha256(bytes(solo.weights.table())).hexdigest()[:16]
    assert solo_sha == sha, (solo_sha, sha)
    b1, b2 = tr.weights.best(), solo.weights.best()
    assert all((x == y).all() for x, y in zip(b1[1] + b1[2], b2[1] + b2[2]))
    print("inorder sharded ok: %d ranks, weights sha %s == single GPU, iterations %d, best score %.6f" % (
        world, sha, tr.weights.table().iteration_count, st.best_score), flush=True)
dist.barrier()
dist.destroy_process_group()
