"""CSV export of the best run: layout and formats of utils/csv_export.rs, numbers equal to the replay."""
import os

import numpy as np
import pytest

import oracle_lib as O
from eirgrid_b200 import _abi, _lib

pytestmark = pytest.mark.gpu


def test_best_run_csv_matches_replay_and_reference_layout(gpu_ctx, oracle_world, tmp_path):
    w = _lib.Weights()
    with pytest.raises(_lib.EirgridError):
        gpu_ctx.export_best_run_csv(w, str(tmp_path))  # no best strategy yet
    res, traj, _, _ = gpu_ctx.rollout(w, 512, seed=77)
    w.update(res, traj)
    out = gpu_ctx.export_best_run_csv(w, str(tmp_path / "enhanced_csv"))
    assert os.path.basename(out).replace("_", "").isdigit() and len(os.path.basename(out)) == 15
    text = open(os.path.join(out, "simulation_summary.csv"), encoding="utf-8").read().split("\n")
    assert text[0] == "Simulation Summary" and text[1].startswith("Timestamp,") and text[3] == "Final Metrics"
    assert text[4].startswith("Final Net Emissions (tonnes CO2),") and text[6].startswith("Total Cost (€),")
    ia = text.index("Actions Taken")
    assert text[ia + 1] == "Year,Action Type,Generator Type,Generator ID,Operation %,Offset Type,Estimated Cost (€)"
    iy = text.index("Yearly Summary Metrics")
    assert text[iy + 1].startswith("Year,Population,PowerUsage,PowerGeneration,PowerBalance,PublicOpinion,YearlyCapitalCost")
    # the best episode of the batch, replayed by the oracle: same yearly rows and final metrics
    has, nb, b, nd, d = w.best()
    k = int(np.lexsort((np.arange(len(res)), -res["score"]))[0])
    eres, _, _, eyearly = oracle_world.replay(traj[k:k + 1])
    assert float(text[4].split(",")[1]) == eres["net_emissions"][0]
    assert text[6].split(",")[1] == "%.2f" % eres["total_cost"][0]
    rows = [r.split(",") for r in text[iy + 2:iy + 28]]
    assert [int(r[0]) for r in rows] == list(range(2025, 2051))
    ey = eyearly["y"][0]
    for y, r in enumerate(rows):
        assert int(r[1]) == ey["total_population"][y] and int(r[14]) == ey["active_generators"][y]
        assert r[2] == "%.2f" % ey["total_power_usage"][y] and r[5] == "%.4f" % ey["average_public_opinion"][y]
        assert r[7] == "%.2f" % ey["total_capital_cost"][y] and r[11] == "%.2f" % ey["net_co2_emissions"][y]
        assert r[18] == "%.2f" % ey["total_cost"][y]
    # actions: the additional actions of the best record, in year order
    acts = [r.split(",") for r in text[ia + 2:iy - 1]]
    want = sum(int(traj[k]["n_additional"][y]) for y in range(26))
    assert len(acts) == want
    first_year_with = next(y for y in range(26) if traj[k]["n_additional"][y])
    a0 = int(traj[k]["actions"][first_year_with][traj[k]["n_deficit"][first_year_with]])
    assert int(acts[0][0]) == 2025 + first_year_with
    assert acts[0][1] == ("AddGenerator" if a0 < 45 else "AddCarbonOffset" if a0 < 57 else acts[0][1])
    if a0 < 45:
        assert acts[0][2] == _abi.GEN_TYPES[a0 // 3] and float(acts[0][6]) > 0
    hist = open(os.path.join(out, "improvement_history.csv"), encoding="utf-8").read().strip().split("\n")
    assert hist[0].startswith("Iteration,Score,Net Emissions (tonnes),Total Cost (€)") and len(hist) >= 2
    st = open(os.path.join(out, "yearly_details", "settlements.csv")).read().strip().split("\n")
    assert st[0] == "Year,Name,Longitude,Latitude,Population,PowerUsage" and len(st) == 1 + 26 * 130
    r0 = st[1].split(",")
    assert r0[0] == "2025" and r0[1] == "S000" and int(r0[4]) == 329487 and abs(float(r0[2]) - (-6.2495)) < 1e-3
