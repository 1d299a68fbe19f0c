#!/usr/bin/env python3
"""Freeze the oracle's outputs: tests/golden/oracle_frozen_r02.npz (round 2: eg_traj / eg_sites store the year rows back to
back; oracle_frozen_r01.npz is the same content in round 1's fixed 40-slots-per-year layout and is still checked).

Everything except the population table, the 2025 generation and the location analysis is "parity unpinned" against the
reference: the oracle's reading of the Rust code is the definition the CUDA path is held to. This fixture
records what that definition yields for a fixed set of inputs, so that any later change of the oracle — or of the compiler /
libm behaviour underneath it — shows up as a failing CPU test instead of silently moving the target.

    python tests/golden/make_oracle_frozen.py      # rewrites the fixture (only after a deliberate change of the oracle)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402
from eirgrid_b200 import _abi  # noqa: E402


def frozen_outputs():
    world = O.World.ireland(fast=True)
    w = O.Weights()
    out = {}
    res, traj, sites, yearly = world.rollout(w, 64, seed=1, first_episode=0)
    out["initial_results"] = res
    out["initial_traj"] = traj
    out["initial_sites_sha1"] = np.frombuffer(hashlib.sha1(sites.tobytes()).digest(), np.uint8)
    out["initial_yearly_sha1"] = np.frombuffer(hashlib.sha1(yearly.tobytes()).digest(), np.uint8)
    w.update(res, traj)                                    # 64 sequential reference-rule updates
    t = w.table()
    out["updated_table_sha1"] = np.frombuffer(hashlib.sha1(bytes(t)).digest(), np.uint8)
    out["updated_weights_2025"] = np.array(t.weights[0][:], np.float64)
    out["updated_weights"], out["updated_deficit_weights"], _ = t.arrays()
    out["updated_iwi"] = np.array([t.iterations_without_improvement, t.iteration_count], np.int64)
    for iwi in (150, 600, 1300):
        t = w.table()
        t.iterations_without_improvement = iwi
        w.set_table(t)
        r2, t2, s2, y2 = world.rollout(w, 32, seed=7 + iwi, first_episode=1000)
        out["iwi%d_results" % iwi] = r2
        out["iwi%d_traj_sha1" % iwi] = np.frombuffer(hashlib.sha1(t2.tobytes()).digest(), np.uint8)
        out["iwi%d_sites_sha1" % iwi] = np.frombuffer(hashlib.sha1(s2.tobytes()).digest(), np.uint8)
    rr, rt, rs, ry = world.replay(traj[:8])                # replaying a record reproduces the episode
    out["replay_results"] = rr
    return out


if __name__ == "__main__":
    o = frozen_outputs()
    np.savez_compressed(os.path.join(HERE, "oracle_frozen_r02.npz"), **o)
    print("wrote", len(o), "arrays;", "mean score %.6f" % o["initial_results"]["score"].mean())
