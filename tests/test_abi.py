"""CPU tests of the C-ABI library: it loads, exports every symbol the header declares, and its host-side logic
(policy table, JSON checkpoints, sequential update) matches the oracle. No device compute is called here."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import oracle_lib as O
from eirgrid_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "eirgrid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eg_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libeirgrid_b200.so does not export %s" % n
    assert sorted(_lib.EXPORTS) == names


def test_struct_sizes_match_header():
    assert _abi.RESULT_DTYPE.itemsize == 64 and _abi.TRAJ_DTYPE.itemsize == 1088
    assert _abi.SITES_DTYPE.itemsize == 1968 and _abi.YEAR_DTYPE.itemsize == 144
    assert C.sizeof(_abi.WeightsTable) == 8 * (26 * (61 + 15 + 21) + 6) + 16
    assert C.sizeof(_abi.RunCfg) == 32 and C.sizeof(_abi.UpdateStats) == 48


def test_no_gpu_fails_loudly_without_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.EirgridError) as e:
        _lib.Context(0)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_weights_new_matches_oracle_table():
    assert bytes(_lib.Weights().table()) == bytes(O.Weights().table())
    assert [_lib.lib().eg_deficit_key_action(k) for k in range(15)] == [24, 21, 36, 33, 27, 0, 3, 12, 30, 15, 6, 9, 39, 42, 60]


def test_sequential_update_matches_oracle(oracle_world):
    ow, gw = O.Weights(), _lib.Weights()
    for it in range(4):
        res, traj, _, _ = oracle_world.rollout(ow, 48, seed=11, first_episode=it * 48)
        so, sg = ow.update(res, traj), gw.update(res, traj)
        assert bytes(ow.table()) == bytes(gw.table())
        assert (so.n_improvements, so.iterations_without_improvement, so.best_score) == \
               (sg.n_improvements, sg.iterations_without_improvement, sg.best_score)
        bo, bg = ow.best(), gw.best()
        assert bo[0] == bg[0] and all(np.array_equal(a, b) for a, b in zip(bo[1] + bo[2], bg[1] + bg[2]))
    # stagnation regimes: forced contrast (> 800) and randomisation (> 1200)
    for iwi in (850, 1300):
        t = ow.table()
        t.iterations_without_improvement = iwi
        ow.set_table(t)
        gw.set_table(t)
        res, traj, _, _ = oracle_world.rollout(ow, 24, seed=12, first_episode=iwi)
        ow.update(res, traj, rng_seed=5)
        gw.update(res, traj, rng_seed=5)
        assert bytes(ow.table()) == bytes(gw.table())


def test_weights_json_schema_and_roundtrip(tmp_path, oracle_world):
    gw = _lib.Weights()
    res, traj, _, _ = oracle_world.rollout(O.Weights(), 16, seed=3)
    gw.update(res, traj)
    path = str(tmp_path / "run" / "latest_weights.json")
    gw.save_to_file(path)
    d = json.load(open(path))
    # SerializableWeights (ai/learning/serialization.rs:37-51)
    assert list(d.keys()) == ["weights", "learning_rate", "best_metrics", "best_weights", "best_actions", "iteration_count",
                              "iterations_without_improvement", "exploration_rate", "deficit_weights",
                              "best_deficit_actions", "optimization_mode", "improvement_history"]
    assert sorted(d["weights"].keys()) == [str(y) for y in range(2025, 2051)]
    entry = d["weights"]["2025"][0]
    assert list(entry[0].keys()) == ["action_type", "generator_type", "generator_id", "operation_percentage", "offset_type", "cost_multiplier"]
    assert entry[0] == {"action_type": "AddGenerator", "generator_type": "OnshoreWind", "generator_id": None,
                        "operation_percentage": None, "offset_type": None, "cost_multiplier": 100}
    assert len(d["weights"]["2025"]) == 61 and len(d["deficit_weights"]["2030"]) == 15
    assert d["weights"]["2025"][57][0] == {"action_type": "UpgradeEfficiency", "generator_type": None, "generator_id": "",
                                           "operation_percentage": None, "offset_type": None, "cost_multiplier": None}
    assert set(d["best_metrics"].keys()) == {"final_net_emissions", "average_public_opinion", "total_cost", "power_reliability"}
    assert d["iteration_count"] == 16 and d["optimization_mode"] is None and len(d["improvement_history"]) >= 1
    g2 = _lib.Weights.load_from_file(path)
    t1, t2 = gw.table(), g2.table()
    for a, b in zip(t1.arrays()[:2], t2.arrays()[:2]):
        assert np.array_equal(a, b)
    assert t2.has_count_weights == 0  # action_count_weights are not serialised (weights/serialization.rs:474)
    assert (t2.has_best, t2.iteration_count, t2.iterations_without_improvement) == (1, t1.iteration_count, t1.iterations_without_improvement)
    assert list(t2.best_metrics) == list(t1.best_metrics)
    b1, b2 = gw.best(), g2.best()
    assert all(np.array_equal(a, b) for a, b in zip(b1[1] + b1[2], b2[1] + b2[2]))
    # merge == update_weights_from: weights overwritten, iteration_count = max
    fresh = _lib.Weights()
    fresh.update_weights_from(g2)
    tf = fresh.table()
    assert np.array_equal(tf.arrays()[0], t2.arrays()[0]) and tf.iteration_count == t2.iteration_count


def test_load_errors_are_reported(tmp_path):
    with pytest.raises(_lib.EirgridError) as e:
        _lib.Weights.load_from_file(str(tmp_path / "missing.json"))
    assert e.value.code == -2
    bad = tmp_path / "bad.json"
    bad.write_text("{\"weights\": {\"2025\": [[{\"action_type\": \"Teleport\"}, 0.1]]}}")
    with pytest.raises(_lib.EirgridError):
        _lib.Weights.load_from_file(str(bad))
    # keys outside the closed key set (a 200 % multiplier) are skipped, not fatal
    ok = tmp_path / "ok.json"
    ok.write_text(json.dumps({"weights": {"2025": [[{"action_type": "AddGenerator", "generator_type": "Nuclear", "cost_multiplier": 200}, 0.5],
                                                   [{"action_type": "DoNothing"}, 0.25]]},
                              "learning_rate": 0.2, "iteration_count": 7, "iterations_without_improvement": 2, "exploration_rate": 0.2,
                              "deficit_weights": {}}))
    w = _lib.Weights.load_from_file(str(ok))
    t = w.table()
    assert t.weights[0][60] == 0.25 and t.weights[0][15] == 0.03 and t.iteration_count == 7
    assert t.deficit_weights[0][0] == 0.15  # defaults when the file has none (weights/serialization.rs:266-283)


def test_malformed_weight_files_fail_like_serde(tmp_path):
    """What serde_json + load_from_file (weights/serialization.rs:141-205) reject must be an error here too, never a crash:
    truncated or non-JSON text, wrong types, non-JSON number tokens, nesting beyond serde's recursion limit, year keys
    that are not u32, unknown action or generator types in the main table."""
    good = tmp_path / "good.json"
    _lib.Weights().save_to_file(str(good))
    txt = good.read_text()
    assert _lib.Weights.load_from_file(str(good)).table().weights[0][0] == 0.08
    cases = {
        "empty": "", "truncated": txt[: len(txt) // 2], "not_json": "hello", "wrong_type": '{"weights": 5}', "array": "[]",
        "no_weights": "{}", "nan": txt.replace("0.08", "NaN", 1), "plus": txt.replace("0.08", "+0.08", 1),
        "hex": txt.replace("0.08", "0x1p-3", 1), "negative_year": txt.replace('"2025"', '"-5"', 1),
        "bad_action": txt.replace('"AddGenerator"', '"Explode"', 1), "bad_generator": txt.replace('"OnshoreWind"', '"Fusion"', 1),
        "deep_array": "[" * 100000, "deep_object": '{"a":' * 100000,
    }
    doc = json.loads(txt)
    for name, key, value in (("count_1e40", "iteration_count", 1e40), ("count_negative", "iterations_without_improvement", -5),
                             ("count_fraction", "iteration_count", 2.5), ("count_string", "iteration_count", "7")):
        cases[name] = json.dumps(dict(doc, **{key: value}))
    for name, text in cases.items():
        p = tmp_path / (name + ".json")
        p.write_text(text)
        with pytest.raises(_lib.EirgridError) as e:
            _lib.Weights.load_from_file(str(p))
        assert e.value.code == -2, name
    # a year outside 2025..2050 is a valid u32 key: loaded and ignored, as the reference's map would hold it unused
    far = tmp_path / "far_year.json"
    far.write_text(txt.replace('"2025"', '"99999"', 1))
    assert _lib.Weights.load_from_file(str(far)).table().weights[1][0] == 0.08


def test_weight_history_file_is_what_the_animation_tool_reads(tmp_path, oracle_world):
    """save_weight_history (multi_simulation.rs:179-207): a JSON array of {iteration, timestamp, weights: to_json(), best_score};
    tools/visualization/weight_history_animation.py reads entry['weights']['weights'][year][action] and the learning rates."""
    w = _lib.Weights()
    path = str(tmp_path / "weight_history.json")
    open(path, "w").write("[]")
    w.history_append(5, path)
    eres, etraj, _, _ = oracle_world.rollout(O.Weights(), 16, seed=5)
    w.update(eres, etraj)
    w.history_append(10, path)
    hist = json.load(open(path))
    assert [h["iteration"] for h in hist] == [5, 10]
    assert hist[0]["best_score"] == 0.0 and hist[1]["best_score"] > 1.0
    assert re.match(r"\d{4}-\d\d-\d\dT\d\d:\d\d:\d\d[+-]\d\d:\d\d$", hist[0]["timestamp"])
    inner = hist[1]["weights"]
    for key in ("weights", "action_count_weights", "learning_rate", "iteration_count", "iterations_without_improvement",
                "exploration_rate", "force_best_actions", "deficit_weights", "guaranteed_best_actions", "optimization_mode", "best_score"):
        assert key in inner, key
    assert set(inner["weights"].keys()) == {str(y) for y in range(2025, 2051)}
    row = inner["weights"]["2025"]
    assert len(row) == 61 and row["AddGenerator(OnshoreWind, 100%)"] == w.table().weights[0][0]
    assert "AddCarbonOffset(Forest, 150%)" in row and "DoNothing" in row and "UpgradeEfficiency()" in row
    assert len(inner["deficit_weights"]["2030"]) == 15 and len(inner["action_count_weights"]["2040"]) == 21
    assert inner["learning_rate"] == 0.2 and inner["iteration_count"] == 16
