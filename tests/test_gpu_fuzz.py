"""Randomised parity (GPU vs oracle): weight tables, run options and recorded trajectories the training path rarely or never
produces — saturated/random weights, every stagnation regime, cost-only scoring, energy sales off, the heuristic count
sampler, 37 random actions in every year (plant list and offset list overflow, maximum-length placement loops)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from eirgrid_b200 import _abi, _lib

pytestmark = pytest.mark.gpu
FLOATS = ("net_emissions", "public_opinion", "total_cost", "power_reliability")
INTS = ("n_generators", "n_offsets", "n_deficit_actions", "n_additional_actions", "flags")


def _assert_equal(res, eres, yearly=None, eyearly=None):
    for f in FLOATS + INTS:
        assert np.array_equal(res[f], eres[f]), f
    np.testing.assert_allclose(res["score"], eres["score"], rtol=1e-12, atol=0)
    if yearly is not None:
        for f in yearly["y"].dtype.names:
            if f != "reserved":
                assert np.array_equal(yearly["y"][f], eyearly["y"][f]), f


@pytest.mark.parametrize("case", range(6))
def test_random_weight_tables_and_options(gpu_ctx, oracle_world, case):
    rs = np.random.RandomState(100 + case)
    ow, gw = O.Weights(), _lib.Weights()
    # give both a best strategy first (the update functions are pinned against each other in test_abi.py)
    eres, etraj, _, _ = oracle_world.rollout(ow, 24, seed=200 + case)
    ow.update(eres, etraj)
    gw.update(eres, etraj)
    t = ow.table()
    w, dw, cw = t.arrays()
    if case % 3 == 0:    # log-uniform between the clamps
        w = np.exp(rs.uniform(np.log(1e-4), np.log(0.999), w.shape))
        dw = np.exp(rs.uniform(np.log(1e-4), np.log(0.999), dw.shape))
    elif case % 3 == 1:  # saturated: most entries at a clamp
        w = rs.choice([1e-4, 0.999, 0.5], size=w.shape, p=[0.6, 0.3, 0.1])
        dw = rs.choice([1e-4, 0.999], size=dw.shape)
    else:                # many equal weights: ties in the sorted sampler
        w = np.round(rs.uniform(0.01, 0.2, w.shape), 2)
        dw = np.round(rs.uniform(0.01, 0.2, dw.shape), 2)
    cw = rs.uniform(0.0, 1.0, cw.shape)
    np.ctypeslib.as_array(t.weights)[:] = w
    np.ctypeslib.as_array(t.deficit_weights)[:] = dw
    np.ctypeslib.as_array(t.count_weights)[:] = cw
    t.iterations_without_improvement = [0, 150, 600, 1300, 5000, 70000][case]
    t.exploration_rate = [0.2, 0.05, 0.5, 0.2, 0.9, 0.01][case]
    t.has_count_weights = case % 2
    ow.set_table(t)
    gw.set_table(t)
    cfg = _abi.RunCfg(cost_only=case in (1, 4), enable_energy_sales=case != 2)
    n = 192
    eres, etraj, esites, eyearly = oracle_world.rollout(ow, n, seed=300 + case, first_episode=10 ** 9 * case, cfg=cfg)
    res, traj, sites, yearly = gpu_ctx.rollout(gw, n, seed=300 + case, first_episode=10 ** 9 * case, cfg=cfg, want_sites=True, want_yearly=True)
    assert traj.tobytes() == etraj.tobytes()
    assert sites.tobytes() == esites.tobytes()
    _assert_equal(res, eres, yearly, eyearly)


def test_random_full_trajectories_hit_every_capacity(gpu_ctx, oracle_world):
    rs = np.random.RandomState(7)
    n = 48
    t = np.zeros(n, _abi.TRAJ_DTYPE)
    for e in range(n):
        rows = []
        for y in range(26):
            nd = rs.randint(0, 8) if y == 0 else rs.randint(0, 3)
            na = rs.randint(0, 38 - nd) if e % 3 else 37 - nd           # a third of the episodes fill the whole record (26 x 37 = 962 of 984 slots)
            acts = rs.randint(0, 61, nd + na)
            if e % 4 == 1:
                acts = rs.randint(0, 45, nd + na)                          # plants only: > 560 plants -> EG_FLAG_GEN_OVERFLOW
            if e % 4 == 2:
                acts = rs.randint(45, 57, nd + na)                         # offsets only: > 520 offsets -> EG_FLAG_OFFSET_OVERFLOW
            if e == 5 and y == 3:                                          # one very long year (the record has no per-year limit)
                acts = np.concatenate([acts, rs.randint(45, 61, 984 - 26 * 37)])
                na = len(acts) - nd
            rows.append((acts[:nd], acts[nd:]))
        t[e] = _abi.pack_traj(rows)
    res, sites, yearly = gpu_ctx.replay(t)
    eres, etraj, esites, eyearly = oracle_world.replay(t)
    # The reference's lists are unbounded (and so are the oracle's); the device keeps at most EG_MAX_NEW_GENERATORS plants
    # and EG_MAX_OFFSETS offsets per episode — more than any sampled episode can build (<= 20 actions a year) — and flags
    # an episode that exceeds them. Unflagged episodes must agree exactly; flagged ones agree up to the year of the overflow.
    over = (res["flags"] & 3) != 0
    assert over.any() and (~over).any(), "both the overflow paths and the exact path must be exercised"
    ok = ~over
    assert sites[ok].tobytes() == esites[ok].tobytes()
    _assert_equal(res[ok], eres[ok], yearly[ok], eyearly[ok])
    assert (res["flags"][ok] == eres["flags"][ok]).all()
    for e in np.nonzero(over)[0]:
        if res["flags"][e] & 1:
            assert res["n_generators"][e] == 560 and eres["n_generators"][e] > 560
        if res["flags"][e] & 2:
            assert res["n_offsets"][e] == 520 and eres["n_offsets"][e] > 520
        # years that ended before the cap was reached are identical
        full_years = [y for y in range(26) if (eyearly["y"]["active_generators"][e, y] - eyearly["y"]["active_generators"][e, 0] < 400)]
        rs_, ers_ = _abi.traj_rows(t[e], sites["site"][e]), _abi.traj_rows(t[e], esites["site"][e])
        for y in full_years[:5]:
            assert rs_[y][0].tobytes() == ers_[y][0].tobytes() and rs_[y][1].tobytes() == ers_[y][1].tobytes()
    assert (res["flags"] & 1).any() and (res["flags"] & 2).any(), "the overflow paths were not exercised"


@pytest.mark.timeout(300)
def test_deficit_handler_ends_on_a_full_plant_list():
    """A map whose 2025 demand needs far more capacity than 560 plants can supply: the deficit loop must leave with
    EG_FLAG_GEN_OVERFLOW set instead of spinning on a plant list that no longer grows — in rollout and in replay mode."""
    from eirgrid_b200 import synthetic
    assets = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ireland_map")
    sx, sy, spop, ex, ey, et, ec, cx, cy = synthetic.load_ireland_arrays(assets)
    spop = np.full_like(spop, 40_000_000)  # 130 settlements x 40 M people x 1 kW = 5.2 TW
    ctx = _lib.Context(0)
    ctx.map_set(sx, sy, spop, ex, ey, et, ec, cx, cy, 51, 1000.0)
    res, traj, _, _ = ctx.rollout(_lib.Weights(), 64, seed=3)
    assert (res["flags"] & _abi.FLAG_GEN_OVERFLOW).all() and (res["n_generators"] == 560).all()
    assert (res["power_reliability"] == 0.0).all()
    rres, _, _ = ctx.replay(traj[:8])
    assert (rres["flags"] & _abi.FLAG_GEN_OVERFLOW).all()
    ctx.close()
