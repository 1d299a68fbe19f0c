#!/usr/bin/env python3
"""SASS opcode histogram per kernel of the built library (cuobjdump -sass on the object files): what the claims in DESIGN.md about
LDGSTS / UBLKCP (cp.async / cp.async.bulk), REDUX, IDP.4A, DMUL-without-DFMA rest on.

    python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["LDGSTS", "UBLKCP", "SYNCS", "REDUX", "CREDUX", "IDP", "DFMA", "DMUL", "DADD", "DSETP", "MUFU", "SHFL", "VOTE", "LDS", "STS", "LDG", "STG", "ATOM", "RED",
         "BAR", "CALL", "FFMA", "HMMA", "UTMALDG", "LDGDEPBAR", "DEPBAR", "MATCH", "WARPSYNC", "ELECT"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "eirgrid_b200", "build", "libeirgrid_b200", "*.o")))
    for obj in objs:
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        func, hist = None, {}
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                func = m.group(1)
                hist[func] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
            if m and func:
                hist[func][m.group(1)] += 1
                full = m.group(1) + m.group(2)
                if m.group(1) in ("IDP", "REDUX", "CREDUX", "SYNCS", "UBLKCP", "LDGSTS", "MUFU"):
                    hist[func]["  " + full] += 1
        for func, h in hist.items():
            total = sum(v for k, v in h.items() if not k.startswith("  "))
            if not total:
                continue
            short = demangle(func)
            short = re.sub(r"\(anonymous namespace\)::", "", short)[:150]
            print("== %s :: %s  (%d SASS instructions)" % (os.path.basename(obj), short, total))
            top = ", ".join("%s %d" % (k, v) for k, v in sorted(((k, v) for k, v in h.items() if not k.startswith("  ")), key=lambda kv: -kv[1])[:14])
            print("   top: " + top)
            watch = ", ".join("%s %d" % (k, h.get(k, 0)) for k in WATCH if h.get(k, 0))
            print("   watched: " + (watch or "-"))
            sub = ", ".join("%s %d" % (k.strip(), v) for k, v in sorted(h.items()) if k.startswith("  "))
            if sub:
                print("   forms: " + sub)
            print("   absent: " + ", ".join(k for k in ("DFMA", "FFMA", "HMMA", "UBLKCP", "LDGSTS", "REDUX", "IDP") if not h.get(k, 0) and not (k == "REDUX" and h.get("CREDUX", 0))))  # CREDUX: the reduction into a uniform register


if __name__ == "__main__":
    main()
