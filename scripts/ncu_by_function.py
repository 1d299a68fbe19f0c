#!/usr/bin/env python3
"""Aggregate the per-line table printed by ncu_summary.py (--top 2000) by device function of episode.cu.

    python scripts/ncu_summary.py rep --top 2000 | python scripts/ncu_by_function.py
"""
import os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
text = open(os.path.join(ROOT, "eirgrid_b200", "csrc", "episode.cu")).read().split("\n")
marks = []  # (line, name)
for i, l in enumerate(text, 1):
    m = re.match(r"\s*__device__ .*?\b(\w+)\(.*\{", l) or re.match(r"\s*__global__ .*?\b(\w+)\(", l)
    if m:
        marks.append((i, m.group(1)))
    for pat, name in ((r"for \(int k0 = 0; k0 < ns", "place: walk step"), (r"lane 0 holds the step's highest static score", "place: evaluation loop+reduce"),
                      (r"for \(int y = 0; y < EG_NY", "run: year loop"),
                      (r"calculate_yearly_metrics, analysis", "run: yearly metrics"), (r"SimulationMetrics from the 2050", "run: result"),
                      (r"const int cls = __ldg\(&T->acc_class\[t\]\);\s*$", "add_generator: sums")):
        if re.search(pat, l):
            marks.append((i, name))
marks.sort()
def fn(ln):
    name = "other/-1"
    for a, n in marks:
        if a <= ln:
            name = n
    return name
agg = {}
for l in sys.stdin:
    m = re.match(r"\s*(-?\d+)\s+([0-9.]+)% inst\s+([0-9.]+)% samples", l)
    if not m:
        continue
    ln = int(m.group(1))
    x = agg.setdefault(fn(ln) if ln > 0 else "other/-1", [0.0, 0.0])
    x[0] += float(m.group(2)); x[1] += float(m.group(3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-36s inst %5.1f%%  samples %5.1f%%" % (k, v[0], v[1]))
