"""Host code of the library (weights, JSON, asset loaders) under AddressSanitizer + UBSan with malformed input.

The harnesses in tests/fuzz/ compile the library's host sources directly with g++ (no CUDA needed) and feed them what a
buggy caller or a damaged file could: records with arbitrary counts and action codes, arbitrary statistics tables and
winner records, truncated / mutated asset files. Any sanitizer report fails the test.
"""
import os
import random
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "eirgrid_b200", "csrc")
FUZZ = os.path.join(ROOT, "tests", "fuzz")
ASSETS = os.path.join(ROOT, "tests", "golden", "ireland_map")
FLAGS = ["-O1", "-g", "-fsanitize=address,undefined,float-cast-overflow,float-divide-by-zero", "-fno-sanitize-recover=undefined", "-std=c++17", "-I" + CSRC,
         "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include"]


def _build(tmp_path, harness, source):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / harness.replace(".cpp", ""))
    r = subprocess.run(["g++"] + FLAGS + [os.path.join(FUZZ, harness), os.path.join(CSRC, source), "-o", exe],
                       capture_output=True, text=True)
    if r.returncode != 0 and ("cannot find -lasan" in r.stderr or "cannot find -lubsan" in r.stderr or "libasan" in r.stderr):
        pytest.skip("sanitizer runtime not installed")
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def _run(exe, args, cwd):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([exe] + args, capture_output=True, text=True, cwd=cwd, env=env, timeout=300)
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    return r.stdout


def test_sequential_update_survives_arbitrary_records(tmp_path):
    out = _run(_build(tmp_path, "fuzz_update.cpp", "weights.cpp"), [], str(tmp_path))
    assert "reload rc=0" in out


def test_batch_apply_survives_arbitrary_statistics_and_winner_records(tmp_path):
    out = _run(_build(tmp_path, "fuzz_batch_apply.cpp", "weights.cpp"), [], str(tmp_path))
    assert out.count("round") == 12


def test_asset_loaders_survive_damaged_files(tmp_path):
    exe = _build(tmp_path, "load_map.cpp", "host_tables.cpp")
    files = [open(os.path.join(ASSETS, f)).read() for f in ("settlements.json", "ireland_generators.csv", "coastline_points.json")]
    good = _run(exe, [os.path.join(ASSETS, f) for f in ("settlements.json", "ireland_generators.csv", "coastline_points.json")], str(tmp_path))
    assert "rc=0" in good and "S=130 E=59 C=200" in good
    rs = random.Random(5)
    variants = []
    for which in range(3):
        text = files[which]
        variants += [(which, ""), (which, text[: len(text) // 2]), (which, "[" * 5000), (which, "{}"), (which, text * 2)]
        for _ in range(6):
            b = bytearray(text.encode())
            for _ in range(15):
                b[rs.randrange(len(b))] = rs.randrange(32, 127)
            variants.append((which, b.decode("latin1")))
    head = files[1].split("\n")[0]
    variants += [(1, head + "\nabc,def,ghi,jkl\n"), (1, head + "\n100,53.0\n"), (1, head + "\n1e400,99.0,-70.0,gas\n"),
                 (0, '{"settlements": [{"name": "x", "lat": "a", "lon": [], "population": -5}]}'),
                 (0, '{"settlements": [{"name": "x", "lat": 53.0, "lon": -7.0, "population": 1e12}]}')]
    names = ["s.json", "g.csv", "c.json"]
    for which, text in variants:
        paths = []
        for k in range(3):
            p = tmp_path / names[k]
            p.write_text(text if k == which else files[k], encoding="latin1")
            paths.append(str(p))
        out = _run(exe, paths, str(tmp_path))
        assert out.startswith("rc=")
