"""Parity and size-independent properties at BASELINE.json's own sizes (GPU).

configs[1]: 10,000 recorded action sequences (oracle-generated: the Rust reference cannot run here) replayed on the GPU,
            integer state bit-exact, floats bit-exact too (north_star bar: <= 1e-5), score within 1e-12.
configs[2]: 65,536 episodes in flight: determinism, invariance to how the batch is cut into launches/shards (what makes
            the multi-GPU split exact), replay round trip, and a 4,096-episode slice against the oracle.
configs[4]: location analysis sharded by site equals the unsharded analysis.
"""
import hashlib

import numpy as np
import pytest

import oracle_lib as O
from eirgrid_b200 import _abi, _lib

pytestmark = pytest.mark.gpu
SCORE_RTOL = 1e-12
FLOATS = ("net_emissions", "public_opinion", "total_cost", "power_reliability")
INTS = ("n_generators", "n_offsets", "n_deficit_actions", "n_additional_actions", "flags")


def test_config1_replay_of_10000_recorded_trajectories(gpu_ctx, oracle_world):
    n = 10000
    eres, etraj, esites, eyearly = oracle_world.rollout(O.Weights(), n, seed=1, first_episode=0)
    res, sites, yearly = gpu_ctx.replay(etraj)
    # integer state: chosen site of every new plant, plant/offset counts, population, active plants, reliability
    assert sites.tobytes() == esites.tobytes()
    for f in INTS:
        assert np.array_equal(res[f], eres[f]), f
    assert np.array_equal(yearly["y"]["total_population"], eyearly["y"]["total_population"])
    assert np.array_equal(yearly["y"]["active_generators"], eyearly["y"]["active_generators"])
    assert np.array_equal(res["power_reliability"], eres["power_reliability"])
    # floats: stated tolerance of the north star is 1e-5 relative; the path is in fact bit-exact in fp64
    for f in yearly["y"].dtype.names:
        if f in ("reserved", "total_population", "active_generators"):
            continue
        g, e = yearly["y"][f], eyearly["y"][f]
        np.testing.assert_allclose(g, e, rtol=1e-5, atol=0, err_msg=f)
        assert np.array_equal(g, e), f
    for f in FLOATS:
        assert np.array_equal(res[f], eres[f]), f
    np.testing.assert_allclose(res["score"], eres["score"], rtol=SCORE_RTOL, atol=0)
    assert (res["flags"] == 0).all()


def test_config2_full_batch_properties(gpu_ctx, oracle_world):
    n = 65536
    w = _lib.Weights()
    res, traj, sites, _ = gpu_ctx.rollout(w, n, seed=20250101, first_episode=0, want_sites=True)
    assert (res["flags"] == 0).all()
    h = hashlib.sha1(res.tobytes() + traj.tobytes() + sites.tobytes()).hexdigest()
    # 1. deterministic: same seed and ids -> same bytes (persistent warps claim episodes in a different order each launch)
    res2, traj2, sites2, _ = gpu_ctx.rollout(w, n, seed=20250101, first_episode=0, want_sites=True)
    assert hashlib.sha1(res2.tobytes() + traj2.tobytes() + sites2.tobytes()).hexdigest() == h
    # 2. cutting the batch into 8 shards (one per GPU of a box) gives the same episodes: ids, not positions, define them
    for k in (0, 3, 7):
        rs, ts, ss, _ = gpu_ctx.rollout(w, n // 8, seed=20250101, first_episode=k * (n // 8), want_sites=True)
        sl = slice(k * (n // 8), (k + 1) * (n // 8))
        assert rs.tobytes() == res[sl].tobytes() and ts.tobytes() == traj[sl].tobytes() and ss.tobytes() == sites[sl].tobytes()
    # 3. record -> replay round trip reproduces sites and metrics of every episode
    rres, rsites, _ = gpu_ctx.replay(traj)
    assert rsites.tobytes() == sites.tobytes()
    for f in FLOATS + INTS:
        assert np.array_equal(rres[f], res[f]), f
    # 4. a slice from the middle of the batch against the oracle
    first, m = 30000, 4096
    eres, etraj, esites, _ = oracle_world.rollout(O.Weights(), m, seed=20250101, first_episode=first, want_yearly=False)
    sl = slice(first, first + m)
    assert traj[sl].tobytes() == etraj.tobytes() and sites[sl].tobytes() == esites.tobytes()
    for f in FLOATS + INTS:
        assert np.array_equal(res[sl][f], eres[f]), f
    # 5. sanity of the sampled population: every episode ends reliable, plants and actions in the expected range
    assert (res["power_reliability"] == 1.0).all()
    assert 20 < res["n_generators"].mean() < 35 and res["n_generators"].max() <= 560


def test_config4_location_analysis_sharded_by_site(gpu_ctx):
    full = gpu_ctx.location_analysis(True)
    n = full.shape[0]
    parts, world = [], 8
    for r in range(world):  # rank r of 8 analyses the sites [r*n/8, (r+1)*n/8)
        lo, hi = r * n // world, (r + 1) * n // world
        parts.append(gpu_ctx.location_analysis(True, first_point=lo, n_points=hi - lo))
    assert np.array_equal(np.concatenate(parts), full)


def test_config0_reference_rule_training_run_equals_cpu_oracle(gpu_ctx, oracle_world):
    """The reference's own loop shape end to end: 16 episodes per snapshot, per-episode update in order, last 10 % of the
    iterations replaying the best strategy — on the GPU and on the CPU oracle from the same seed: identical weights."""
    n_iter, chunk, seed = 480, 16, 20250101
    final_full = n_iter * 10 // 100
    ow, gw = O.Weights(), _lib.Weights()
    done = 0
    while done < n_iter:
        replay = done >= n_iter - final_full and gw.has_best_actions()
        assert replay == (done >= n_iter - final_full and bool(ow.best()[0]))
        cfg = _abi.RunCfg(replay_best=int(replay))
        eres, etraj, _, _ = oracle_world.rollout(ow, chunk, seed=seed, first_episode=done, cfg=cfg, want_sites=False, want_yearly=False)
        res, traj, _, _ = gpu_ctx.rollout(gw, chunk, seed=seed, first_episode=done, cfg=cfg)
        assert traj.tobytes() == etraj.tobytes()
        ow.update(eres, etraj, replay=replay)
        gw.update(res, traj, replay_best=replay)
        done += chunk
    assert bytes(ow.table()) == bytes(gw.table())
    ho, bo, do_ = ow.best()
    hg, bg, dg = gw.best()
    assert ho and hg and all(np.array_equal(x, y) for x, y in zip(bo + do_, bg + dg))


def test_config4_location_analysis_all_26_years(gpu_ctx, oracle_world):
    """All sites x 15 types x 26 years: year y uses the settlements' populations of that year (round-half-away growth of 1 %).
    Checked against the oracle's analysis of a world whose settlements carry those populations."""
    from eirgrid_b200 import synthetic
    import os
    assets = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ireland_map")
    sx, sy, spop, ex, ey, et, ec, cx, cy = synthetic.load_ireland_arrays(assets)
    base = gpu_ctx.location_analysis(True)
    assert np.array_equal(gpu_ctx.location_analysis(True, year_index=0), base)
    pop = spop.astype(np.float64)
    changed = 0
    for y in range(1, 26):
        pop = np.floor(pop * 1.01 + 0.5)  # f64::round of positive values
        if y in (1, 12, 25):
            w = O.World.from_arrays(sx, sy, pop.astype(np.uint32), ex, ey, et, ec, cx, cy, 51, 1000.0, fast=False)
            got = gpu_ctx.location_analysis(True, year_index=y)
            assert np.array_equal(got, w.location_analysis(True)), y
            changed += int((got != base).sum())
    assert changed > 0, "population growth must move at least one urban / nearby-population decision by 2050"
