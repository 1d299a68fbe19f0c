#!/usr/bin/env python3
"""Hot SASS instructions of an .ncu-rep (source page): address, executed count, samples, source line, text.

    python scripts/ncu_sass_hot.py rep.ncu-rep [min_share_percent]
"""
import csv, io, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
src = list(csv.reader(io.StringIO(out)))
h = src[1]
ci, cs, ct, csrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed"), h.index("Source")
rows = src[2:]
tot = sum(int(r[ci] or 0) for r in rows)
tots = sum(int(r[cs] or 0) for r in rows)
base = int(rows[0][0], 16)
prev = None
for n, r in enumerate(rows):
    e = int(r[ci] or 0)
    if 100.0 * e / tot >= thr:
        if prev is not None and n != prev + 1:
            print("   ...")
        prev = n
        print("%6x %5.2f%% %5.2f%%s lanes %4.1f  %s" % (int(r[0], 16) - base, 100.0 * e / tot, 100.0 * int(r[cs] or 0) / tots, int(r[ct] or 0) / max(e, 1), r[csrc][:100]))
