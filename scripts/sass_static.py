#!/usr/bin/env python3
"""Static SASS size of the rollout kernel per device function / source line of episode.cu (needs -lineinfo).

    python scripts/sass_static.py [lib.so] [--lines N] [--mode 0|1|2]
"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else os.path.join(ROOT, "eirgrid_b200", "libeirgrid_b200.so")
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
MODE = sys.argv[sys.argv.index("--mode") + 1] if "--mode" in sys.argv else "1"  # 0 general, 1 lean/plain sampler, 2 lean/stagnation sampler
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("episode")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
text = open(os.path.join(ROOT, "eirgrid_b200", "csrc", "episode.cu")).read().split("\n")
marks = []
for i, l in enumerate(text, 1):
    m = re.match(r"\s*__device__ .*?\b(\w+)\(.*\{", l) or re.match(r"\s*__global__ .*?\b(\w+)\(", l)
    if m:
        marks.append((i, m.group(1)))
def fn(ln):
    name = "?"
    for a, n in marks:
        if a <= ln:
            name = n
    return name
cnt, byfn, on, line = collections.Counter(), collections.Counter(), False, None
for l in dis.split("\n"):
    s = l.strip()
    if s.startswith(".text."):
        on = ("eg_episode_kernelILb0ELi0ELi%sEE" % MODE) in s
    m = re.match(r'//## File "([^"]+)", line (\d+)', s)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if on and re.match(r"/\*[0-9a-f]{4,}\*/", s):
        cnt[line] += 1
        byfn[fn(line[1]) if line and line[0] == "episode.cu" else (line[0] if line else "?")] += 1
tot = sum(cnt.values())
print("rollout kernel: %d instructions = %.1f KB" % (tot, tot * 16 / 1024))
for k, v in byfn.most_common(40):
    print("  %5d  %s" % (v, k))
print("heaviest lines:")
for (f, ln), v in cnt.most_common(nlines):
    print("  %5d  %s:%d  %s" % (v, f, ln, text[ln - 1].strip()[:90] if f == "episode.cu" else ""))
