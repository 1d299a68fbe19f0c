"""Known-answer tests derived by hand from the reference's Rust (tests/kat_cases.py shows the arithmetic step by step, with the
Rust line of every step). The CPU oracle, the host library and the CUDA kernels are each held to the same numbers, so a shared
misreading of learning.rs / metrics_calculation.rs / map_handler.rs in oracle/ and csrc/ (written by the same hand) cannot
hide behind their mutual agreement.

Tolerance: products of a handful of doubles; exp / pow go through the platform libm here and in the oracle, through the
correctly rounded csrc/eg_math.hpp in the product -> 4 ulp (rtol 1e-15) covers a last-bit difference in one factor.
"""
import numpy as np
import pytest

import kat_cases as K
import oracle_lib as O
from eirgrid_b200 import _abi, _lib

RTOL = 1e-15 * 4


def _check_tables(table, w, dw, what):
    wa, dwa, _ = table.arrays()
    init_w, init_dw, _ = _lib.Weights().table().arrays()
    touched = np.zeros_like(wa, bool)
    for (y, a), v in w.items():
        np.testing.assert_allclose(wa[y, a], v, rtol=RTOL, atol=0, err_msg="%s: weight of year %d action %d" % (what, 2025 + y, a))
        touched[y, a] = True
    assert np.array_equal(wa[~touched], init_w[~touched]), what + ": an entry outside the two toy years' actions moved"
    touched_d = np.zeros_like(dwa, bool)
    for (y, k), v in dw.items():
        np.testing.assert_allclose(dwa[y, k], v, rtol=RTOL, atol=0, err_msg="%s: deficit weight of year %d key %d" % (what, 2025 + y, k))
        touched_d[y, k] = True
    assert np.array_equal(dwa[~touched_d], init_dw[~touched_d]), what + ": a deficit entry outside the toy case moved"
    assert table.iterations_without_improvement == 1 and table.iteration_count == 2 and table.has_best


def test_contrast_step_hand_values_oracle_and_host():
    recs, ress, w, dw = K.contrast_case()
    # the hand values themselves: boosted twice from 0.02 then mildly penalised, etc. — spot values that are easy to check by eye
    assert w[(0, K.BATTERY)] == min(min(0.07 * 1.4, 0.999) * 1.4, 0.999)
    assert w[(1, K.FOREST)] < 0.02 and w[(0, K.GAS_CC)] < w[(0, K.OFFSHORE)] * 0.06 / 0.08 * 1.0000001  # penalised twice vs once
    ow, gw = O.Weights(), _lib.Weights()
    ow.update(ress, recs)
    gw.update(ress, recs)
    _check_tables(ow.table(), w, dw, "oracle")
    _check_tables(gw.table(), w, dw, "host library")
    for weights in (ow, gw):
        has, best, best_def = weights.best()
        assert has and best[0].tolist() == [K.GAS_PEAKER, K.BATTERY, K.ONSHORE] and best_def[0].tolist() == [K.GAS_PEAKER, K.BATTERY]
        assert best[1].tolist() == [K.UTILITY_SOLAR, K.DO_NOTHING] and best_def[1].tolist() == []


def test_score_metrics_hand_values():
    """scoring.rs:5-45: emitting run; net-zero below and above 8 x MAX_ACCEPTABLE_COST (cost weight 0.5 -> 0.8)"""
    assert K.score_metrics(200000.0, 0.9, 1.0e10) == 0.8
    assert K.score_metrics(-1.0, 0.8, 2.0e10) == 1.9
    # cost 5e11 = 10 x 5e10: ln 10 / ln 100 = 0.5 -> cost score 0.5, weight 0.8: 1 + (0.5 * 0.8 + 0.7 * 0.2)
    hand = 1.0 + ((1.0 - 0.5) * 0.8 + 0.7 * (1.0 - 0.8))
    np.testing.assert_allclose(K.score_metrics(-5.0, 0.7, 5.0e11), hand, rtol=RTOL)
    # the product's score (host form of csrc/update_rule.hpp; fn 3 fixes the opinion at 0.5)
    got = _lib.rule_math(3, np.array([-5.0, 200000.0, -1.0, 2.0e6]), np.array([5.0e11, 1.0e10, 2.0e10, 1.0]))
    want = [1.0 + ((1.0 - 0.5) * 0.8 + 0.5 * (1.0 - 0.8)), 0.8, 1.0 + (1.0 * 0.5 + 0.5 * 0.5), 0.0]
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=0)
    # and the update's decisions rest on it: an improving episode is one whose score is strictly greater
    ow = O.Weights()
    r = np.concatenate([K.result(-5.0, 0.7, 5.0e11), K.result(-5.0, 0.7, 5.0e11), K.result(-5.0, 0.7000001, 5.0e11)])
    st = ow.update(r, np.zeros(3, _abi.TRAJ_DTYPE))
    assert st.n_improvements == 2 and st.iterations_without_improvement == 0
    np.testing.assert_allclose(st.best_score, K.score_metrics(-5.0, 0.7000001, 5.0e11), rtol=RTOL)


def _check_toy_year(res, sites, yearly, what):
    site, m = K.toy_2025_metrics()
    assert int(sites["site"][0][0]) == site, "%s: placement of the first plant (hand: site %d)" % (what, site)
    y0 = yearly["y"][0][0]
    for f, v in m.items():
        if isinstance(v, int):
            assert int(y0[f]) == v, (what, f)
        else:
            np.testing.assert_allclose(y0[f], v, rtol=RTOL, atol=0, err_msg="%s: 2025 %s" % (what, f))
    assert int(res["n_generators"][0]) >= 1 and int(res["n_offsets"][0]) == 0  # later years run into deficits of their own


def test_one_plant_world_2025_hand_values_oracle():
    """settlement demand, generator output, CO2, cost, public opinion, energy sales of 2025 and the placement search, by hand"""
    site, m = K.toy_2025_metrics()
    assert site == 10 * 51 + 10 or site == 30 * 51 + 20  # the search's maximum sits on one of the two settlements
    assert m["total_power_usage"] == pytest.approx(30.6, rel=1e-12) and m["total_power_generation"] == 49.5
    for fast in (True, False):
        world = O.World.from_arrays(*K.toy_map_arrays(), fast=fast)
        res, _, sites, yearly = world.replay(K.toy_record(), mode=O.FAST if fast else O.FAITHFUL)
        _check_toy_year(res, sites, yearly, "oracle fast=%s" % fast)


def _check_offset_years(res, sites, yearly, what):
    site, years = K.toy_offsets_metrics()
    assert int(sites["site"][0][0]) == site
    for yi, m in enumerate(years):
        row = yearly["y"][0][yi]
        for f, v in m.items():
            if isinstance(v, int):
                assert int(row[f]) == v, (what, 2025 + yi, f)
            else:
                np.testing.assert_allclose(row[f], v, rtol=RTOL, atol=1e-9 if v == 0.0 else 0, err_msg="%s: %d %s" % (what, 2025 + yi, f))
    assert int(res["n_offsets"][0]) == 3


def test_offsets_and_second_year_hand_values_oracle():
    """CO2 and offset accounting (maturity curve, capture efficiency), offset cost with multipliers, carbon-credit revenue,
    re-priced capital cost of year 2 minus year 1, population growth and per-capita demand of 2026, by hand"""
    _, years = K.toy_offsets_metrics()
    assert years[0]["total_carbon_offset"] == 127500.0 and years[0]["yearly_carbon_credit_revenue"] == 126000.0 * 75.0
    assert years[1]["total_population"] == 20200 + 10100 and years[1]["total_carbon_offset"] > years[0]["total_carbon_offset"]
    for fast in (True, False):
        world = O.World.from_arrays(*K.toy_map_arrays(), fast=fast)
        res, _, sites, yearly = world.replay(K.toy_offsets_record(), mode=O.FAST if fast else O.FAITHFUL)
        _check_offset_years(res, sites, yearly, "oracle fast=%s" % fast)


@pytest.mark.gpu
def test_offsets_and_second_year_hand_values_kernel():
    ctx = _lib.Context(0)
    ctx.map_set(*K.toy_map_arrays())
    res, sites, yearly = ctx.replay(K.toy_offsets_record())
    _check_offset_years(res, sites, yearly, "CUDA kernel")
    ctx.close()


@pytest.mark.gpu
def test_one_plant_world_2025_hand_values_kernel():
    ctx = _lib.Context(0)
    ctx.map_set(*K.toy_map_arrays())
    res, sites, yearly = ctx.replay(K.toy_record())
    _check_toy_year(res, sites, yearly, "CUDA kernel")
    ctx.close()


@pytest.mark.gpu
def test_contrast_step_hand_values_device_update(gpu_ctx):
    import torch
    recs, ress, w, dw = K.contrast_case()
    dev = torch.device("cuda", 0)
    d_res = torch.from_numpy(ress.view(np.uint8).reshape(-1).copy()).to(dev)
    d_traj = torch.from_numpy(recs.view(np.uint8).reshape(-1).copy()).to(dev)
    torch.cuda.synchronize()
    gw = _lib.Weights()
    st = gpu_ctx.update_device(gw, 2, d_res, d_traj)
    _check_tables(gw.table(), w, dw, "device update")
    assert st.n_improvements == 1 and st.n_contrast_applied == 1
