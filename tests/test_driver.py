"""Host driver: checkpoint directory layout and resume rule (CPU), a short checkpointed run with resume (GPU)."""
import datetime
import json
import os

import pytest

from eirgrid_b200 import simulation as S
from eirgrid_b200.__main__ import parse_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "tests", "golden", "ireland_map")


def test_cli_flags_and_defaults_match_reference():
    a = parse_args([])
    assert (a.iterations, a.checkpoint_dir, a.checkpoint_interval, a.progress_interval, a.cache_dir) == (1000, "checkpoints", 5, 10, "cache")
    assert a.parallel and a.enable_energy_sales and a.enable_csv_export  # SetTrue flags with default true: always true
    assert not (a.no_continue or a.force_full_simulation or a.cost_only or a.enable_construction_delays or a.track_weight_history)
    assert a.seed is None
    b = parse_args(["-n", "50", "-c", "ck", "-i", "7", "-r", "3", "-C", "cc", "--seed", "12345", "--cost-only", "--no-continue"])
    assert (b.iterations, b.checkpoint_dir, b.checkpoint_interval, b.progress_interval, b.cache_dir, b.seed) == (50, "ck", 7, 3, "cc", 12345)


def test_run_dir_name_has_literal_2024_prefix():
    assert S.run_dir_name(datetime.datetime(2031, 3, 9, 14, 5, 59)) == "20240309_140559"


def test_resume_rule(tmp_path):
    ck = tmp_path / "checkpoints"
    for name in ("20240101_000000", "20240309_140559", "20260101_000000", "2024_bad", "notes", "20241399_000000"):
        (ck / name).mkdir(parents=True)
    assert os.path.basename(S.find_resume_dir(str(ck))) == "20240309_140559"   # 2026 and month 13 are rejected
    # start iteration: max over names made of digits/'_' only (multi_simulation.rs:391-394), which includes 2026...
    (ck / "20260101_000000" / "checkpoint_iteration.txt").write_text("40\n")
    (ck / "20240309_140559" / "checkpoint_iteration.txt").write_text("25")
    assert S.find_start_iteration(str(ck)) == 40
    assert S.find_resume_dir(str(tmp_path / "none")) is None and S.find_start_iteration(str(tmp_path / "none")) == 0


@pytest.mark.gpu
def test_checkpointed_run_and_resume(tmp_path):
    ck = str(tmp_path / "ck")
    s1 = S.run_multi_simulation(ASSETS, 4096, checkpoint_dir=ck, checkpoint_interval=5, cache_dir=str(tmp_path / "cache"),
                                force_full_simulation=False, batch_size=2048, master_seed=3, log=lambda *a: None)
    d = s1["run_dir"]
    assert sorted(os.listdir(d)) == ["best_weights.json", "checkpoint_iteration.txt", "enhanced_csv", "latest_weights.json"]
    assert open(os.path.join(d, "checkpoint_iteration.txt")).read() == "4096"
    w = json.load(open(os.path.join(d, "latest_weights.json")))
    assert w["iteration_count"] == 4096 and w["best_metrics"] is not None
    # without a location-analysis cache every iteration is a "full run": after the first batch the best strategy is
    # replayed (multi_simulation.rs:444-465)
    assert s1["best_score"] > 1.0
    s2 = S.run_multi_simulation(ASSETS, 6144, checkpoint_dir=ck, cache_dir=str(tmp_path / "cache"), batch_size=2048,
                                master_seed=4, log=lambda *a: None)
    assert s2["start_iteration"] == 4096 and s2["iterations"] == 6144
    assert s2["best_score"] >= s1["best_score"]
