"""GPU tests of the in-order update on the device (csrc/update.cu, eg_update_device): the reference's per-episode rule
(core/multi_simulation.rs:494-508, weights/learning.rs:131-373, weights/strategy.rs:19-258) applied by the GPU must leave
the weights object in the SAME state, bit for bit, as the host form eg_update on the same records — in every stagnation
regime, for replay records (quirk Q10), across the 16,384-episode passes, and for records no rollout would produce."""
import os

import numpy as np
import pytest

import oracle_lib as O
from eirgrid_b200 import _abi, _lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "tests", "golden", "ireland_map")


def _to_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).to(torch.device("cuda", 0))


def _state(w):
    t = w.table()
    has, b, d = w.best()
    return (bytes(t), has, [x.tolist() for x in b], [x.tolist() for x in d])


def _assert_same(host_w, dev_w, host_st, dev_st, what):
    hs, ds = _state(host_w), _state(dev_w)
    th, td = host_w.table(), dev_w.table()
    for name, a, b in zip(("weights", "deficit_weights", "count_weights"), th.arrays(), td.arrays()):
        assert np.array_equal(a, b), "%s: %s differ in %d entries, max rel %.3g" % (
            what, name, int((a != b).sum()), float(np.abs(a / b - 1).max()))
    assert hs == ds, what
    for f in ("n_episodes", "n_improvements", "n_contrast_applied", "iterations_without_improvement", "best_score",
              "batch_best_score", "batch_best_episode", "n_flagged"):
        assert getattr(host_st, f) == getattr(dev_st, f), (what, f, getattr(host_st, f), getattr(dev_st, f))


def _both(gpu_ctx, w, res, traj, replay=False, rng_seed=0):
    """host update on a clone, device update on another clone; returns (host weights, device weights, stats)"""
    import torch
    hw, dw = w.clone(), w.clone()
    hst = hw.update(res, traj, replay_best=replay, rng_seed=rng_seed)
    d_res, d_traj = _to_dev(res), _to_dev(traj)
    torch.cuda.synchronize()
    dst = gpu_ctx.update_device(dw, len(res), d_res, d_traj, replay_best=replay, rng_seed=rng_seed)
    return hw, dw, hst, dst


def _weights_with(gpu_ctx, iwi, seed=3, trained_batches=2, oracle=False):
    """a weights object with a best strategy, a few learning steps behind it and the given stagnation counter
    (oracle=True: also the CPU oracle's weights object taken through the same records)"""
    w, ow = _lib.Weights(), O.Weights()
    for i in range(trained_batches):
        res, traj, _, _ = gpu_ctx.rollout(w, 64, seed=seed, first_episode=64 * i)
        w.update(res, traj, rng_seed=seed)
        ow.update(res, traj, rng_seed=seed)
    for x in (w, ow):
        t = x.table()
        t.iterations_without_improvement = iwi
        x.set_table(t)
    return (w, ow) if oracle else w


@pytest.mark.parametrize("iwi", [0, 90, 450, 790, 1190, 1500])
def test_device_update_equals_host_update_in_every_regime(gpu_ctx, iwi):
    """thresholded contrast (< 800), forced contrast (> 800), the randomisation stream (> 1200), and the crossings between
    them inside one batch (790 -> 800+, 1190 -> 1200+)"""
    w = _weights_with(gpu_ctx, iwi)
    res, traj, _, _ = gpu_ctx.rollout(w, 700, seed=100 + iwi, first_episode=10_000)
    hw, dw, hst, dst = _both(gpu_ctx, w, res, traj, rng_seed=77)
    _assert_same(hw, dw, hst, dst, "iwi %d" % iwi)
    assert hst.n_contrast_applied > 0
    # a second batch on top of the first: the state the first call left on the host object is what the second starts from
    res2, traj2, _, _ = gpu_ctx.rollout(hw, 300, seed=200 + iwi, first_episode=20_000)
    hst2 = hw.update(res2, traj2, rng_seed=78)
    import torch
    d_res, d_traj = _to_dev(res2), _to_dev(traj2)
    torch.cuda.synchronize()
    dst2 = gpu_ctx.update_device(dw, 300, d_res, d_traj, rng_seed=78)
    _assert_same(hw, dw, hst2, dst2, "iwi %d, second batch" % iwi)


def test_device_update_from_fresh_weights_and_history(gpu_ctx):
    """no best strategy yet: the first episode always becomes the best (strategy.rs:58), contrast starts with the second"""
    w = _lib.Weights()
    res, traj, _, _ = gpu_ctx.rollout(w, 2000, seed=5)
    hw, dw, hst, dst = _both(gpu_ctx, w, res, traj)
    _assert_same(hw, dw, hst, dst, "fresh")
    assert hst.n_improvements >= 2
    # improvement history: same iterations, scores and metrics (the timestamps are wall-clock)
    import json
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        hw.save_to_file(os.path.join(d, "h.json"))
        dw.save_to_file(os.path.join(d, "d.json"))
        jh, jd = json.load(open(os.path.join(d, "h.json"))), json.load(open(os.path.join(d, "d.json")))
    strip = lambda hist: [{k: v for k, v in r.items() if k != "timestamp"} for r in hist]
    assert strip(jh["improvement_history"]) == strip(jd["improvement_history"]) and len(jh["improvement_history"]) == hst.n_improvements
    assert jh["best_weights"] == jd["best_weights"] and jh["best_metrics"] == jd["best_metrics"]
    for n in (1, 2, 31):  # tiny batches
        hw, dw, hst, dst = _both(gpu_ctx, w, res[:n], traj[:n])
        _assert_same(hw, dw, hst, dst, "fresh, n=%d" % n)


def test_device_update_replay_records_quirk_q10(gpu_ctx):
    """replay iterations record every additional action twice and the first four deficit actions twice in the deficit
    list; the best strategy an improving replay episode leaves behind has those doubled lists"""
    w = _weights_with(gpu_ctx, 20)
    cfg = _abi.RunCfg(replay_best=1)
    res, traj, _, _ = gpu_ctx.rollout(w, 500, seed=9, first_episode=50_000, cfg=cfg)
    hw, dw, hst, dst = _both(gpu_ctx, w, res, traj, replay=True, rng_seed=4)
    _assert_same(hw, dw, hst, dst, "replay")
    # a second replay generation on top (lists of the best strategy grow, 18 -> 42 -> ...)
    res2, traj2, _, _ = gpu_ctx.rollout(hw, 500, seed=10, first_episode=60_000, cfg=cfg)
    hw2, dw2, hst2, dst2 = _both(gpu_ctx, hw, res2, traj2, replay=True, rng_seed=4)
    _assert_same(hw2, dw2, hst2, dst2, "replay, second generation")


def test_device_update_65536_records_across_passes_and_against_the_oracle(gpu_ctx, oracle_world):
    """BASELINE configs[2] batch shape: 65,536 records = four passes of the device pipeline; and the CPU oracle (glibc
    exp/pow where the product uses the correctly rounded csrc/eg_math.hpp) ends with the same discrete state and the same
    table to 1e-9"""
    for iwi in (0, 1000):
        w, ow = _weights_with(gpu_ctx, iwi, seed=12, oracle=True)
        n = 65536
        res, traj, _, _ = gpu_ctx.rollout(w, n, seed=31 + iwi, first_episode=1 << 20)
        hw, dw, hst, dst = _both(gpu_ctx, w, res, traj, rng_seed=11)
        _assert_same(hw, dw, hst, dst, "65536 records, iwi %d" % iwi)
        assert hst.n_flagged == 0
        ost = ow.update(res, traj, rng_seed=11)
        assert (ost.n_improvements, ost.iterations_without_improvement) == (dst.n_improvements, dst.iterations_without_improvement)
        to, td = ow.table(), dw.table()
        for a, b in zip(to.arrays()[:2], td.arrays()[:2]):
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=0)
        bo, bd = ow.best(), dw.best()
        assert all(np.array_equal(x, y) for x, y in zip(bo[1] + bo[2], bd[1] + bd[2]))
    # odd sizes around the pass length
    w = _weights_with(gpu_ctx, 300, seed=13)
    res, traj, _, _ = gpu_ctx.rollout(w, 16384 * 2 + 1, seed=77, first_episode=1 << 21)
    for n in (16383, 16384, 16385, 16384 * 2 + 1):
        hw, dw, hst, dst = _both(gpu_ctx, w, res[:n], traj[:n], rng_seed=3)
        _assert_same(hw, dw, hst, dst, "n=%d" % n)


def test_device_update_on_records_no_rollout_produces(gpu_ctx):
    """synthetic records: scores that improve again and again (hundreds of best-strategy changes in one batch), the same
    action hundreds of times in one year (more multiplications than the byte-sized count holds), action codes outside the
    key set, counts that run past the record's capacity, flagged episodes"""
    rng = np.random.default_rng(2024)
    n = 3000
    res = np.zeros(n, _abi.RESULT_DTYPE)
    res["net_emissions"] = np.where(rng.random(n) < 0.3, rng.uniform(1, 2e6, n), -rng.uniform(0, 1e5, n))
    res["public_opinion"] = rng.uniform(0.2, 1.0, n) * np.linspace(0.5, 1.0, n)  # slowly rising: many improvements
    res["total_cost"] = 10 ** rng.uniform(9, 13, n)
    res["power_reliability"] = (rng.random(n) < 0.9).astype(np.float64)
    res["flags"] = (rng.random(n) < 0.01).astype(np.uint32) * 4
    traj = np.zeros(n, _abi.TRAJ_DTYPE)
    for e in range(n):
        kind = rng.integers(0, 6)
        nd = np.zeros(26, np.int64)
        na = np.zeros(26, np.int64)
        if kind == 0:      # one year holds nearly everything, one action repeated
            y = rng.integers(0, 26)
            nd[y], na[y] = rng.integers(0, 400), rng.integers(0, 500)
        elif kind == 1:    # counts past the capacity
            nd[:] = rng.integers(0, 60, 26)
            na[:] = rng.integers(0, 60, 26)
        else:
            nd[0] = rng.integers(0, 14)
            nd[1:] = rng.integers(0, 2, 25)
            na[:] = rng.integers(0, 6, 26)
        traj["n_deficit"][e] = nd
        traj["n_additional"][e] = na
        acts = rng.integers(0, 61, _abi.TRAJ_CAPACITY)
        if kind == 0:
            acts[:] = rng.choice([24, 36, 60, 1, 45])
            acts[rng.integers(0, _abi.TRAJ_CAPACITY, 30)] = rng.integers(0, 61, 30)
        if kind == 2:
            acts[rng.integers(0, _abi.TRAJ_CAPACITY, 40)] = rng.integers(61, 256, 40)  # not action codes
        if kind == 3:      # deficit rows made of deficit keys only, like real records
            acts[:] = rng.choice([24, 21, 36, 33, 27, 0, 3, 12, 30, 15, 6, 9, 39, 42, 60], _abi.TRAJ_CAPACITY)
        traj["actions"][e] = acts
    for iwi, replay in ((0, False), (850, False), (1300, True), (40, True)):
        w = _weights_with(gpu_ctx, iwi, seed=21)
        hw, dw, hst, dst = _both(gpu_ctx, w, res, traj, replay=replay, rng_seed=99)
        _assert_same(hw, dw, hst, dst, "synthetic records, iwi %d, replay %s" % (iwi, replay))
        assert hst.n_improvements > 5 and hst.n_flagged > 0
    # and from fresh weights (no best strategy): the first record becomes the best whatever its score
    hw, dw, hst, dst = _both(gpu_ctx, _lib.Weights(), res, traj, rng_seed=5)
    _assert_same(hw, dw, hst, dst, "synthetic records, fresh weights")


def test_rule_arithmetic_host_and_device_agree_bit_for_bit():
    """the premise of the above: csrc/eg_math.hpp and csrc/update_rule.hpp give the same bits on the host and on the device"""
    rng = np.random.default_rng(7)
    n = 200_000
    cases = [
        (0, -rng.uniform(0, 40, n), None), (1, 10 ** rng.uniform(-3, 6, n), None), (2, rng.uniform(0, 1.2, n), 0.3),
        (2, rng.integers(0, 200000, n) / 10.0, 1.8), (3, rng.uniform(-1e5, 2e6, n), 10 ** rng.uniform(9, 13, n)),
        (4, rng.uniform(0.5, 2.2, n), rng.integers(0, 3000, n)), (5, rng.uniform(0.5, 2.2, n), rng.integers(0, 3000, n)),
        (6, rng.uniform(0.5, 2.2, n), rng.integers(0, 3000, n)), (7, np.zeros(n), rng.integers(0, 100000, n)),
        (8, np.zeros(n), rng.integers(0, 100000, n)), (9, rng.integers(0, 1 << 32, n), rng.integers(0, 1 << 32, n)),
    ]
    for fn, x, y in cases:
        h = _lib.rule_math(fn, x, y, device=-1)
        d = _lib.rule_math(fn, x, y, device=0)
        assert h.tobytes() == d.tobytes(), "fn %d: %d of %d values differ" % (fn, int((h.view(np.uint64) != d.view(np.uint64)).sum()), n)


def test_train_batch_inorder_equals_rollout_plus_host_update(gpu_ctx):
    """eg_train_batch_inorder = snapshot upload + rollout + in-order device update: same weights as the two host-buffer calls"""
    w1, w2 = _lib.Weights(), _lib.Weights()
    for step in range(3):
        n = 4096
        res, traj, _, _ = gpu_ctx.rollout(w1, n, seed=8, first_episode=step * n)
        s1 = w1.update(res, traj, rng_seed=step)
        s2 = gpu_ctx.train_batch_inorder(w2, n, seed=8, first_episode=step * n, rng_seed=step)
        _assert_same(w1, w2, s1, s2, "step %d" % step)


def test_inorder_rule_sharded_over_two_gpus_equals_one_gpu():
    """BatchTrainer.step_inorder_sharded under torchrun with 2 ranks (rollouts sharded, records all-gathered over NCCL, in-order
    update replicated): same weights as one GPU alone, bit for bit. Needs two GPUs; skipped on a single-GPU box."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "scripts", "inorder_sharded_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "inorder sharded ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
