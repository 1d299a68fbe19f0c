"""world_size-2 gloo test of the multi-rank exchange step on the CPU: episode-id sharding, int64 statistics
allreduce, batch-winner selection and the identical host-side apply. The device-side statistics are replaced by
their numpy restatement (tests/stats_ref.py), everything after them is the product's own host code."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _best_lists(w):
    has, b, d = w.best()
    return [x.tolist() for x in b], [x.tolist() for x in d]


def _setup_weights():
    """A snapshot with a best strategy so that the contrast statistics are non-trivial."""
    import oracle_lib as O
    from eirgrid_b200 import _lib
    world = O.World.ireland(fast=True)
    ow, gw = O.Weights(), _lib.Weights()
    res, traj, _, _ = world.rollout(ow, 24, seed=5, threads=2)
    ow.update(res, traj)
    gw.update(res, traj)
    return world, ow, gw


def _one_step(gw, ow, world, first, n_per_rank, rank, world_size, use_dist):
    import stats_ref
    from eirgrid_b200 import trainer as T
    lo = first + rank * n_per_rank
    res, traj, _, _ = world.rollout(ow, n_per_rank, seed=77, first_episode=lo, threads=2)
    t = gw.table()
    consts = stats_ref.contrast_consts(t, stats_ref.default_score(*list(t.best_metrics)[:3]))
    best, best_def = _best_lists(gw)
    stats, scores = stats_ref.batch_stats(res, traj, consts, best, best_def)
    k = int(np.lexsort((np.arange(n_per_rank), -scores))[0])
    rec = T.pack_record(scores[k], lo + k, res[k:k + 1], traj[k:k + 1])
    st = torch.from_numpy(stats)
    all_rec = torch.zeros(world_size * T.REC_BYTES, dtype=torch.uint8)
    if use_dist:
        dist.all_reduce(st, op=dist.ReduceOp.SUM)
        dist.all_gather_into_tensor(all_rec, torch.from_numpy(rec))
    else:
        all_rec[:] = torch.from_numpy(rec)
    return T.combine_and_apply(gw, st.numpy(), all_rec.numpy(), n_per_rank * world_size, first)


def _worker(rank, world_size, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    world, ow, gw = _setup_weights()
    first = 1000
    for step in range(2):
        st = _one_step(gw, ow, world, first, 32, rank, world_size, True)
        ow.set_table(gw.table())  # sampling weights for the next batch follow the product's update
        first += 64
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.frombuffer(bytes(gw.table()), np.uint8))
    np.save(os.path.join(out_dir, "iwi%d.npy" % rank), np.array([st.iterations_without_improvement, st.n_improvements, st.batch_best_episode]))
    dist.destroy_process_group()


def test_two_ranks_match_one_rank(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, r1), "ranks diverged"
    # single process over the same 64-episode batches
    world, ow, gw = _setup_weights()
    first = 1000
    for step in range(2):
        import stats_ref
        from eirgrid_b200 import trainer as T
        res, traj, _, _ = world.rollout(ow, 64, seed=77, first_episode=first, threads=2)
        t = gw.table()
        consts = stats_ref.contrast_consts(t, stats_ref.default_score(*list(t.best_metrics)[:3]))
        best, best_def = _best_lists(gw)
        stats, scores = stats_ref.batch_stats(res, traj, consts, best, best_def)
        k = int(np.lexsort((np.arange(64), -scores))[0])
        rec = T.pack_record(scores[k], first + k, res[k:k + 1], traj[k:k + 1])
        T.combine_and_apply(gw, stats, rec, 64, first)
        ow.set_table(gw.table())
        first += 64
    single = np.frombuffer(bytes(gw.table()), np.uint8)
    assert np.array_equal(single, r0), "2-rank result differs from the 1-rank result"
