"""Synthetic scaled maps for BASELINE config 4 ("synthetic 10x scaled grid").

The shipped Irish map (130 settlements, 59 existing plants, 200 coastline points; loaders settlements_loader.rs:23-42,
generators_loader.rs:133-207) is scaled by `factor`: settlements are resampled with Gaussian jitter around the real ones,
populations are resampled from the empirical distribution, existing plants keep the real fuel mix and capacity
distribution, and the candidate-site grid gets factor x as many sites over the same 50 km box. Deterministic in the seeds.
Returns the arrays `Context.map_set` / `eg_map_set` take (grid metres, already transformed and clamped).
"""
import csv
import json
import math
import os

import numpy as np

MAP_MAX = 50000.0                      # data/poi.rs:11-15 clamp
LON0, LAT0 = -10.6, 51.4               # const_funcs.rs:124-136
X_PER_DEG, Y_PER_DEG = 10638.297872340427, 12500.0
FUEL_TYPE = {"gas": 7, "oil": 8, "wind": 0, "hydro": 10, "coal": 6, "biomass": 9}  # generators_loader.rs:60-116


def to_grid(lat, lon):
    """transform_lat_lon_to_grid + Coordinate::new clamp (const_funcs.rs:124-136, poi.rs:11-15)."""
    x = (np.asarray(lon, dtype=np.float64) - LON0) * X_PER_DEG
    y = (np.asarray(lat, dtype=np.float64) - LAT0) * Y_PER_DEG
    return np.clip(x, 0.0, MAP_MAX), np.clip(y, 0.0, MAP_MAX)


def load_ireland_arrays(asset_dir):
    """The shipped map as arrays: (sx, sy, spop, ex, ey, etype, ecap, cx, cy)."""
    st = json.load(open(os.path.join(asset_dir, "settlements.json")))["settlements"]
    sx, sy = to_grid([s["lat"] for s in st], [s["lon"] for s in st])
    spop = np.array([int(s["population"]) for s in st], dtype=np.uint32)
    ex, ey, et, ec = [], [], [], []
    with open(os.path.join(asset_dir, "ireland_generators.csv")) as f:
        for row in csv.DictReader(f):
            t = FUEL_TYPE.get(row["primary_fuel"].strip().lower())
            if t is None:
                continue
            x, y = to_grid(float(row["latitude"]), float(row["longitude"]))
            ex.append(float(x)); ey.append(float(y)); et.append(t); ec.append(float(row["capacity_mw"]))
    pts = json.load(open(os.path.join(asset_dir, "coastline_points.json")))["grid_coords"]  # map_handler.rs:360-375
    cx = np.clip(np.array([float(p[0]) for p in pts]), 0.0, MAP_MAX)
    cy = np.clip(np.array([float(p[1]) for p in pts]), 0.0, MAP_MAX)
    return (sx, sy, spop, np.array(ex), np.array(ey), np.array(et, dtype=np.uint8), np.array(ec), cx, cy)


MAX_SETTLEMENT_FACTOR = 3


def scaled_map(asset_dir, factor=10, settlement_seed=42, plant_seed=43, jitter_m=2000.0, grid_n=None, settlement_factor=None):
    """factor x existing plants and ~factor x candidate sites over the same box; settlements scaled by `settlement_factor`.

    The reference's placement score is a product over ALL settlements of (1 + pop/1e6) / (1 + d/1e4)
    (gpu/metal_location_search.rs:128-134): 1e-90..1e-46 on the shipped 130 settlements, so it underflows to 0 for every
    site beyond ~4x as many settlements and the search then finds nothing. The settlement count is therefore scaled by
    min(factor, 3) (1e-270..1e-138), which is as far as the reference's own arithmetic allows. Populations are divided by
    the settlement factor and plant capacities by `factor`, so national demand and the existing fleet's total capacity stay
    those of the shipped map and episodes build a comparable number of plants.

    Returns (sx, sy, spop, ex, ey, etype, ecap, cx, cy, grid_n, grid_step)."""
    sx, sy, spop, ex, ey, et, ec, cx, cy = load_ireland_arrays(asset_dir)
    sf = min(factor, MAX_SETTLEMENT_FACTOR) if settlement_factor is None else int(settlement_factor)
    rs = np.random.RandomState(settlement_seed)
    n_s = len(sx) * sf
    src = rs.randint(0, len(sx), n_s)
    nsx = np.clip(sx[src] + rs.normal(0.0, jitter_m, n_s), 0.0, MAP_MAX)
    nsy = np.clip(sy[src] + rs.normal(0.0, jitter_m, n_s), 0.0, MAP_MAX)
    npop = np.maximum(spop[rs.randint(0, len(spop), n_s)] // sf, 1).astype(np.uint32)
    rp = np.random.RandomState(plant_seed)
    n_e = len(ex) * factor
    srcp = rp.randint(0, len(ex), n_e)
    nex = np.clip(ex[srcp] + rp.normal(0.0, jitter_m, n_e), 0.0, MAP_MAX)
    ney = np.clip(ey[srcp] + rp.normal(0.0, jitter_m, n_e), 0.0, MAP_MAX)
    net = et[srcp].astype(np.uint8)
    nec = ec[rp.randint(0, len(ec), n_e)] / float(factor)
    if grid_n is None:
        grid_n = int(round(51 * math.sqrt(factor)))  # 10x sites: 161^2 / 51^2 = 9.97
    step = float(int(MAP_MAX // (grid_n - 1)))     # integer-valued step (eg_map_desc.grid_step)
    return (np.ascontiguousarray(nsx), np.ascontiguousarray(nsy), npop, np.ascontiguousarray(nex), np.ascontiguousarray(ney),
            net, np.ascontiguousarray(nec), cx.copy(), cy.copy(), int(grid_n), step)
