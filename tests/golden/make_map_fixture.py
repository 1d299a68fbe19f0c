#!/usr/bin/env python3
"""Regenerate tests/golden/ireland_map/* from the reference's shipped map assets.

The GPU box has no /root/reference, so the three input files the reference's loaders read
(aiSimulator/assets/{settlements.json, ireland_generators.csv, coastline_points.json};
loaders: src/data/settlements_loader.rs:23-42, src/data/generators_loader.rs:133-207,
src/utils/map_handler.rs:360-375) are re-emitted here in the SAME schemas, stripped to the
fields those loaders actually use (settlement names are replaced by an index tag, the
constituent-settlement lists and the unused lat/lon copy of the coastline are dropped).
Numeric values are copied verbatim (repr round-trip), so every derived quantity is bit-identical.

Run in the build container only:  python tests/golden/make_map_fixture.py
"""
import csv, json, os, sys

REF = os.environ.get("EIRGRID_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "aiSimulator", "assets")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ireland_map")


def main():
    os.makedirs(DST, exist_ok=True)
    s = json.load(open(os.path.join(SRC, "settlements.json")))["settlements"]
    out = {"settlements": [
        {"name": "S%03d" % i, "lat": e["lat"], "lon": e["lon"], "population": e["population"],
         "grid_x": e["grid_x"], "grid_y": e["grid_y"]} for i, e in enumerate(s)]}
    json.dump(out, open(os.path.join(DST, "settlements.json"), "w"), indent=1)
    c = json.load(open(os.path.join(SRC, "coastline_points.json")))
    json.dump({"grid_coords": c["grid_coords"]}, open(os.path.join(DST, "coastline_points.json"), "w"))
    with open(os.path.join(SRC, "ireland_generators.csv")) as f, \
            open(os.path.join(DST, "ireland_generators.csv"), "w", newline="") as g:
        rows = list(csv.reader(f))
        w = csv.writer(g, lineterminator="\n")
        for r in rows:
            w.writerow(r[:4])
    print("wrote", DST, len(out["settlements"]), "settlements,", len(rows) - 1, "generators,",
          len(c["grid_coords"]), "coastline points")


if __name__ == "__main__":
    sys.exit(main())
