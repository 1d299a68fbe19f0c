"""Test-side loader for the CPU oracle (oracle/_build/liboracle.so). Test infrastructure only.

The oracle is the checker: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs are the only importers of this module.
"""
import csv
import ctypes as C
import json
import os
import subprocess

import numpy as np

from eirgrid_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
MAP_DIR = os.path.join(ROOT, "tests", "golden", "ireland_map")

FAITHFUL, FAST = 0, 1


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp"))]
    srcs.append(os.path.join(ROOT, "include", "eirgrid_b200.h"))
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_world_new.restype = C.c_void_p
        L.orc_world_new.argtypes = [C.c_int, C.c_double]
        L.orc_weights_new.restype = C.c_void_p
        L.orc_weights_clone.restype = C.c_void_p
        L.orc_weights_clone.argtypes = [C.c_void_p]
        for f in ("orc_world_free", "orc_world_build_fast", "orc_weights_free"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = None
        L.orc_world_add_settlement_raw.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_uint32]
        L.orc_world_add_settlement_xy.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_uint32]
        L.orc_world_add_existing_raw.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_world_add_existing_xy.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_world_add_coast.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_fuel_to_type.argtypes = [C.c_char_p]
        L.orc_world_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.orc_world_settlement.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                           C.POINTER(C.c_uint32)]
        L.orc_world_existing.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double)]
        L.orc_world_prefix.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_world_site_static.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_world_demand.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_world_existing_generation_if_operational.argtypes = [C.c_void_p]
        L.orc_world_existing_generation_if_operational.restype = C.c_double
        L.orc_world_existing_online_year.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_powi.argtypes = [C.c_double, C.c_int]
        L.orc_powi.restype = C.c_double
        L.orc_inflation.argtypes = [C.c_int]
        L.orc_inflation.restype = C.c_double
        L.orc_gen_cost.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_gen_cost.restype = C.c_double
        L.orc_score.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_score.restype = C.c_double
        L.orc_weights_get_table.argtypes = [C.c_void_p, C.POINTER(_abi.WeightsTable)]
        L.orc_weights_set_table.argtypes = [C.c_void_p, C.POINTER(_abi.WeightsTable)]
        L.orc_weights_get_best.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_train_batch.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_abi.RunCfg), C.c_uint64, C.c_uint64, C.c_uint32,
                                      C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(_abi.UpdateStats)]
        L.orc_rollout.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_abi.RunCfg), C.c_uint64, C.c_uint64, C.c_uint32,
                                  C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_replay.argtypes = [C.c_void_p, C.POINTER(_abi.RunCfg), C.c_void_p, C.c_uint32, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint64,
                                 C.POINTER(_abi.UpdateStats)]
        L.orc_location_analysis.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_uint32,
                                            C.c_uint32]
        _lib = L
    return _lib


def read_map_files(map_dir=MAP_DIR):
    """Parse the three asset files (same schemas as aiSimulator/assets) into plain rows."""
    s = json.load(open(os.path.join(map_dir, "settlements.json")))["settlements"]
    settlements = [(e["lat"], e["lon"], int(e["population"])) for e in s]
    with open(os.path.join(map_dir, "ireland_generators.csv")) as f:
        rows = list(csv.reader(f))[1:]
    generators = [(float(r[0]), float(r[1]), float(r[2]), r[3]) for r in rows]
    coast = json.load(open(os.path.join(map_dir, "coastline_points.json")))["grid_coords"]
    return settlements, generators, coast


class World:
    def __init__(self, grid_n=51, step=1000.0):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_world_new(grid_n, step))
        self.grid_n = grid_n
        self.step = step
        self.fast = False

    @classmethod
    def ireland(cls, fast=True, map_dir=MAP_DIR):
        w = cls()
        settlements, generators, coast = read_map_files(map_dir)
        for lat, lon, pop in settlements:
            w.L.orc_world_add_settlement_raw(w.h, lat, lon, pop)
        for cap, lat, lon, fuel in generators:
            t = w.L.orc_fuel_to_type(fuel.encode())
            assert t >= 0, fuel
            w.L.orc_world_add_existing_raw(w.h, cap, lat, lon, t)
        for x, y in coast:
            w.L.orc_world_add_coast(w.h, x, y)
        if fast:
            w.build_fast()
        return w

    @classmethod
    def from_arrays(cls, sx, sy, spop, ex, ey, etype, ecap, cx, cy, grid_n, step, fast=True):
        w = cls(grid_n, step)
        for x, y, p in zip(sx, sy, spop):
            w.L.orc_world_add_settlement_xy(w.h, float(x), float(y), int(p))
        for x, y, t, c in zip(ex, ey, etype, ecap):
            w.L.orc_world_add_existing_xy(w.h, float(c), float(x), float(y), int(t))
        for x, y in zip(cx, cy):
            w.L.orc_world_add_coast(w.h, float(x), float(y))
        if fast:
            w.build_fast()
        return w

    def build_fast(self):
        self.L.orc_world_build_fast(self.h)
        self.fast = True

    def counts(self):
        out = (C.c_uint32 * 3)()
        self.L.orc_world_counts(self.h, out)
        return tuple(out)

    def arrays(self):
        """Loader-transformed map as arrays (feeds eg_map_set)."""
        ns, ne, nc = self.counts()
        sx, sy, sp = np.zeros(ns), np.zeros(ns), np.zeros(ns, np.uint32)
        x, y, p = C.c_double(), C.c_double(), C.c_uint32()
        for i in range(ns):
            self.L.orc_world_settlement(self.h, i, C.byref(x), C.byref(y), C.byref(p))
            sx[i], sy[i], sp[i] = x.value, y.value, p.value
        ex, ey, et, ec = np.zeros(ne), np.zeros(ne), np.zeros(ne, np.uint8), np.zeros(ne)
        t, cap, pl, co = C.c_int(), C.c_double(), C.c_double(), C.c_double()
        for i in range(ne):
            self.L.orc_world_existing(self.h, i, C.byref(x), C.byref(y), C.byref(t), C.byref(cap), C.byref(pl),
                                      C.byref(co))
            ex[i], ey[i], et[i], ec[i] = x.value, y.value, t.value, cap.value
        return sx, sy, sp, ex, ey, et, ec

    def demand(self):
        pop = np.zeros(26, np.uint32)
        usage = np.zeros(26)
        self.L.orc_world_demand(self.h, _abi.ptr(pop), _abi.ptr(usage))
        return pop, usage

    def prefix(self, yidx, rclass):
        out = np.zeros(self.grid_n * self.grid_n)
        assert self.L.orc_world_prefix(self.h, yidx, rclass, _abi.ptr(out)) == 0
        return out

    def site_static(self):
        n = self.grid_n * self.grid_n
        c, o = np.zeros(n), np.zeros(n)
        assert self.L.orc_world_site_static(self.h, _abi.ptr(c), _abi.ptr(o)) == 0
        return c, o

    def rollout(self, weights, n, seed=1, first_episode=0, cfg=None, mode=FAST, literal_scan=False, threads=0,
                want_traj=True, want_sites=True, want_yearly=True):
        cfg = cfg or _abi.RunCfg()
        threads = threads or (os.cpu_count() or 1)
        res = np.zeros(n, _abi.RESULT_DTYPE)
        traj = np.zeros(n, _abi.TRAJ_DTYPE) if want_traj else None
        sites = np.zeros(n, _abi.SITES_DTYPE) if want_sites else None
        yearly = np.zeros(n, _abi.YEARLY_DTYPE) if want_yearly else None
        rc = self.L.orc_rollout(self.h, weights.h, C.byref(cfg), seed, first_episode, n, mode, int(literal_scan),
                                threads, _abi.ptr(res), _abi.ptr(traj), _abi.ptr(sites), _abi.ptr(yearly))
        assert rc == 0
        return res, traj, sites, yearly

    def replay(self, traj_in, cfg=None, mode=FAST, threads=0):
        cfg = cfg or _abi.RunCfg()
        threads = threads or (os.cpu_count() or 1)
        n = len(traj_in)
        res = np.zeros(n, _abi.RESULT_DTYPE)
        traj = np.zeros(n, _abi.TRAJ_DTYPE)
        sites = np.zeros(n, _abi.SITES_DTYPE)
        yearly = np.zeros(n, _abi.YEARLY_DTYPE)
        rc = self.L.orc_replay(self.h, C.byref(cfg), _abi.ptr(np.ascontiguousarray(traj_in)), n, mode, threads,
                               _abi.ptr(res), _abi.ptr(traj), _abi.ptr(sites), _abi.ptr(yearly))
        assert rc == 0
        return res, traj, sites, yearly

    def location_analysis(self, use_loaded_map, half=25, step=2000.0, first=0, n=None):
        side = 2 * half + 1
        n = side * side - first if n is None else n
        out = np.zeros((n, 15))
        assert self.L.orc_location_analysis(self.h, int(use_loaded_map), half, step, _abi.ptr(out), first, n) == 0
        return out

    def __del__(self):
        try:
            self.L.orc_world_free(self.h)
        except Exception:
            pass


class Weights:
    def __init__(self, handle=None):
        self.L = lib()
        self.h = C.c_void_p(handle if handle is not None else self.L.orc_weights_new())

    def clone(self):
        return Weights(self.L.orc_weights_clone(self.h))

    def table(self):
        t = _abi.WeightsTable()
        self.L.orc_weights_get_table(self.h, C.byref(t))
        return t

    def set_table(self, t):
        self.L.orc_weights_set_table(self.h, C.byref(t))

    def best(self):
        """(has_best, best_actions, best_deficit_actions): two lists of 26 uint8 arrays of any length"""
        nb, nd = np.zeros(26, np.uint32), np.zeros(26, np.uint32)
        has = self.L.orc_weights_get_best(self.h, _abi.ptr(nb), None, 0, _abi.ptr(nd), None, 0)
        b, d = np.zeros(max(int(nb.sum()), 1), np.uint8), np.zeros(max(int(nd.sum()), 1), np.uint8)
        self.L.orc_weights_get_best(self.h, _abi.ptr(nb), _abi.ptr(b), b.size, _abi.ptr(nd), _abi.ptr(d), d.size)
        ob, od = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64), np.concatenate([[0], np.cumsum(nd)]).astype(np.int64)
        return (bool(has), [b[ob[y]:ob[y + 1]].copy() for y in range(26)], [d[od[y]:od[y + 1]].copy() for y in range(26)])

    def train_batch(self, world, n, seed=1, first_episode=0, cfg=None, mode=FAST, threads=0, rng_seed=0):
        """One batch of the reference loop with the episodes' own unbounded action lists (no eg_traj between rollout
        and update). Returns (results, records, stats); the records are for comparison only."""
        cfg = cfg or _abi.RunCfg()
        threads = threads or (os.cpu_count() or 1)
        res = np.zeros(n, _abi.RESULT_DTYPE)
        traj = np.zeros(n, _abi.TRAJ_DTYPE)
        st = _abi.UpdateStats()
        rc = self.L.orc_train_batch(world.h, self.h, C.byref(cfg), seed, first_episode, n, mode, threads, rng_seed,
                                    _abi.ptr(res), _abi.ptr(traj), C.byref(st))
        assert rc == 0
        return res, traj, st

    def update(self, results, trajs, replay=False, rng_seed=0):
        st = _abi.UpdateStats()
        rc = self.L.orc_update(self.h, _abi.ptr(np.ascontiguousarray(results)), _abi.ptr(np.ascontiguousarray(trajs)),
                               len(results), int(replay), rng_seed, C.byref(st))
        assert rc == 0
        return st

    def __del__(self):
        try:
            self.L.orc_weights_free(self.h)
        except Exception:
            pass
