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
    k = int(np.lexsort((np.arange(len(res)), -res["score"]))[0])
    rec = _abi.traj_rows(traj[k])  # 26 x (deficit actions, additional actions)
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
    a0 = int(rec[first_year_with][1][0])
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
    # yearly_details/generators.csv: one row per active generator and year (the exporter's second pass, csv_export.rs:813-979)
    gen = open(os.path.join(out, "yearly_details", "generators.csv"), encoding="utf-8").read().strip().split("\n")
    assert gen[0].startswith("Year,Generator ID,Type,Longitude,Latitude,Power Output (MW),Efficiency (%),Operation (%),CO2 Output (tonnes),Is Active,")
    assert gen[0].endswith("Reliability Factor,Planning Time (years),Construction Time (years),Construction Speed")
    grow = [r.split(",") for r in gen[1:]]
    assert all(len(r) == 20 for r in grow)
    per_year = {y: [r for r in grow if int(r[0]) == 2025 + y] for y in range(26)}
    for y in range(26):
        assert len(per_year[y]) == ey["active_generators"][y]
    # 2025: the pre-existing fleet is still Planned (quirk Q1), only plants of the run appear, numbered from the fleet size up
    ids_2025 = [r[1] for r in per_year[0]]
    assert all(i.startswith("Gen_") and i.split("_")[2] == "2025" for i in ids_2025)
    assert [int(i.split("_")[3]) for i in ids_2025] == list(range(59, 59 + len(ids_2025)))
    kinds = [int(a) // 3 for a in list(rec[0][0]) + list(rec[0][1]) if a < 45]
    assert [r[2] for r in per_year[0]] == [_abi.GEN_TYPES[t] for t in kinds]
    r = per_year[0][0]
    h = sum(ord(ch) for ch in r[1])                                   # id-hash coordinates, csv_export.rs:867-869
    x, yy = 5000.0 + (h % 100) / 100.0 * 40000.0, 5000.0 + ((h // 100) % 100) / 100.0 * 40000.0
    assert r[3] == "%.6f" % (-10.6 + 4.7 * (x / 50000.0)) or abs(float(r[3]) - (-10.6 + 4.7 * x / 50000.0)) < 2e-6
    assert abs(float(r[4]) - (51.4 + 4.0 * yy / 50000.0)) < 2e-6
    assert r[6] == "99.00" and r[7] == "10000.00" and r[9] == "true" and r[10] == "2025" and r[11] == "2050" and r[19] == "Normal"
    last = per_year[25]
    ex = [r for r in last if r[1].startswith("Existing_")]
    assert len(ex) == 59 and ex[0][1].split("_")[2] == "0" and ex[0][10] == "0" and ex[0][11] == "25"
    wind = next(r for r in ex if r[2] == "OnshoreWind")
    assert wind[5] == "50.00" and wind[13] == "75000000.00" and wind[16] == "0.35" and wind[17] == "1.50" and wind[18] == "1.25"
    # yearly_details/carbon_offsets.csv: offsets stay Planned on the export map, so the offset columns are zero (csv_export.rs:987-1093)
    off = open(os.path.join(out, "yearly_details", "carbon_offsets.csv"), encoding="utf-8").read().strip().split("\n")
    assert off[0].startswith("Year,Offset ID,Type,X,Y,Size,Capture Efficiency (%),Power Consumption (MW),CO2 Offset (tonnes),Negative CO2 Emissions (tonnes),Cost (€)")
    codes = [(2025 + y, int(a)) for y in range(26) for a in rec[y][1] if 45 <= a < 57]
    orow = [r.split(",") for r in off[1:]]
    assert len(orow) == sum(2051 - yr for yr, _ in codes)
    if codes:
        yr, a = codes[0]
        first = next(r for r in orow if r[1].endswith("_0"))
        name = ["Forest", "Wetland", "ActiveCapture", "CarbonCredit"][(a - 45) // 3]
        assert first[0] == str(yr) and first[1] == "Offset_%s_%d_0" % (name, yr) and first[2] == name
        assert first[5] == "0" and first[6] == "85.00" and first[8] == "0.00" and first[9] == "-0.00" and first[13] == "0.00"
        base = [1e6, 1e6, 1e9, 5e7][(a - 45) // 3] * 1.0185 ** (yr - 2025) * [1.0, 1.2, 1.5][(a - 45) % 3]
        assert abs(float(first[10]) - base) <= 1e-9 * base + 0.006
        assert -10.6 <= float(first[3]) <= -5.9 and 51.4 <= float(first[4]) <= 55.4
    logs = open(os.path.join(out, "operation_logs", "generator_operation_logs.csv"), encoding="utf-8").read().strip().split("\n")
    assert logs == ["Year,Month,Day,Hour,Generator ID,Type,Power Output (MW),Operation %,Actual Output (MW),Weather Factor,CO2 Emissions (tonnes)"]
