"""The native (C++) host driver host/eirgrid_host.cpp: builds against the C ABI only, fails loudly without a GPU, and on a
GPU runs the batch/checkpoint loop with the reference's run-directory layout and resume rule."""
import json
import os
import subprocess

import pytest

from eirgrid_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "tests", "golden", "ireland_map")
HOST = os.path.join(ROOT, "host", "_build", "eirgrid_host")


@pytest.fixture(scope="module")
def host_binary():
    from eirgrid_b200 import build
    build.build()
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "host")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return HOST


def test_host_links_only_the_c_abi(host_binary):
    out = subprocess.run(["ldd", host_binary], capture_output=True, text=True).stdout
    assert "libeirgrid_b200.so" in out
    src = open(os.path.join(ROOT, "host", "eirgrid_host.cpp")).read()
    assert "cuda_runtime" not in src and "oracle" not in src.replace("// ", "")
    help_text = subprocess.run([host_binary, "--help"], capture_output=True, text=True).stdout
    for flag in ("--iterations", "--no-continue", "--checkpoint-dir", "--checkpoint-interval", "--progress-interval", "--cache-dir",
                 "--force-full-simulation", "--seed", "--cost-only", "--enable-energy-sales", "--enable-construction-delays",
                 "--track-weight-history"):
        assert flag in help_text, flag


def test_host_fails_loudly_without_gpu(host_binary):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([host_binary, "-n", "10", "--assets", ASSETS, "--no-continue"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_host_run_checkpoint_and_resume(host_binary, tmp_path):
    ck = str(tmp_path / "ck")
    cache = str(tmp_path / "cache")
    os.makedirs(cache)
    open(os.path.join(cache, "location_analysis.json"), "w").write("{}")
    cmd = [host_binary, "--assets", ASSETS, "-c", ck, "-C", cache, "--batch-size", "2048", "--master-seed", "7", "-i", "1000"]
    r = subprocess.run(cmd + ["-n", "8192", "--no-continue"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    s1 = json.loads(r.stdout.strip().split("\n")[-1])
    assert s1["iterations"] == 8192 and s1["best_score"] is not None and s1["kernel_launches"] > 0
    run1 = s1["run_dir"]
    assert os.path.basename(run1).startswith("2024") and len(os.path.basename(run1)) == 15
    for f in ("latest_weights.json", "best_weights.json", "checkpoint_iteration.txt"):
        assert os.path.exists(os.path.join(run1, f)), f
    assert open(os.path.join(run1, "checkpoint_iteration.txt")).read().strip() == "8192"
    w = _lib.Weights.load_from_file(os.path.join(run1, "latest_weights.json"))
    assert w.table().iteration_count == 8192 and w.has_best_actions()
    # resume: continues from iteration 8192 with the saved weights; the last 10 % replay the best strategy
    import time
    time.sleep(1.1)  # a new timestamp directory
    r = subprocess.run(cmd + ["-n", "12288"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    s2 = json.loads(r.stdout.strip().split("\n")[-1])
    assert s2["start_iteration"] == 8192 and s2["iterations"] == 12288 and s2["run_dir"] != run1
    assert s2["best_score"] >= s1["best_score"]


@pytest.mark.gpu
def test_host_batch_equals_python_trainer(host_binary, tmp_path):
    """Same seed, same batches: the C++ driver's weights equal the Python BatchTrainer's bit for bit (both call the same
    kernels and the same host update; this pins eg_train_batch_* + eg_update_combine_apply against the torch plumbing)."""
    import numpy as np
    from eirgrid_b200 import trainer as T
    ck = str(tmp_path / "ck")
    cache = str(tmp_path / "cache")
    os.makedirs(cache)
    open(os.path.join(cache, "location_analysis.json"), "w").write("{}")
    # 2 batches of 4096: 0 + 819 < 8192 and 4096 + 819 < 8192, so neither is a replay ("full run") batch
    r = subprocess.run([host_binary, "--assets", ASSETS, "-c", ck, "-C", cache, "--batch-size", "4096", "--master-seed", "11",
                        "-n", "8192", "--no-continue", "-i", "100000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    s = json.loads(r.stdout.strip().split("\n")[-1])
    tr = T.BatchTrainer(4096, seed=11, device=0, asset_dir=ASSETS)
    for _ in range(2):
        tr.step()
    want = tr.weights.table()
    tr.close()
    got = _lib.Weights.load_from_file(os.path.join(s["run_dir"], "latest_weights.json")).table()
    assert got.iteration_count == want.iteration_count == 8192
    assert got.iterations_without_improvement == want.iterations_without_improvement
    assert np.array_equal(np.ctypeslib.as_array(got.weights), np.ctypeslib.as_array(want.weights))
    assert np.array_equal(np.ctypeslib.as_array(got.deficit_weights), np.ctypeslib.as_array(want.deficit_weights))
    assert list(got.best_metrics) == list(want.best_metrics)
