#!/usr/bin/env python3
"""Regenerate tests/golden/location_analysis_scores.npz from the reference's shipped cache/location_analysis.json.

That file is the output of LocationAnalysis::analyze_map (utils/map_handler.rs:61-142) on an EMPTY map with
min_suitability 0.3 (aiSimulator/bin/analyze_locations.rs:19-46): 2601 entries in loop order (i, j in [-25, 25],
step 2000 m, negatives clamped to 0) with the per-type scores that reached the threshold. It is the only
machine-readable output of the reference that pins calculate_generator_suitability.
Stored here as scores[2601, 15] (NaN = type absent, i.e. score < 0.3) plus the coordinates.
Run in the build container only:  python tests/golden/make_location_golden.py
"""
import json
import os

import numpy as np

REF = os.environ.get("EIRGRID_REFERENCE", "/root/reference")
TYPES = ["OnshoreWind", "OffshoreWind", "DomesticSolar", "CommercialSolar", "UtilitySolar", "Nuclear", "CoalPlant",
         "GasCombinedCycle", "GasPeaker", "Biomass", "HydroDam", "PumpedStorage", "BatteryStorage", "TidalGenerator",
         "WaveEnergy"]
d = json.load(open(os.path.join(REF, "cache", "location_analysis.json")))
locs = d["locations"]
scores = np.full((len(locs), 15), np.nan)
xy = np.zeros((len(locs), 2))
for k, e in enumerate(locs):
    xy[k] = (e["coordinate"]["x"], e["coordinate"]["y"])
    for name, v in e["suitability_scores"].items():
        scores[k, TYPES.index(name)] = v
counts = np.array([d["type_counts"].get(t, 0) for t in TYPES])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "location_analysis_scores.npz")
remaining = np.array([d["remaining_spaces"].get(t, 0) for t in TYPES])
multi_xy = np.array([[c["x"], c["y"]] for c, _ in d["multi_type_locations"]])
multi_n_types = np.array([len(ts) for _, ts in d["multi_type_locations"]])
np.savez_compressed(out, scores=scores, xy=xy, type_counts=counts, remaining_spaces=remaining, multi_xy=multi_xy,
                    multi_n_types=multi_n_types, n_exhausted=np.array(len(d["exhausted_types"])),
                    top_level_keys=np.array(sorted(d.keys())))
print("wrote", out, scores.shape, counts)
