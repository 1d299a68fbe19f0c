#!/usr/bin/env python3
"""Hot code footprint of the rollout kernel: static SASS instructions executed at least `thr` times in the captured launch,
by device function of episode.cu (the kernel is sensitive to what has to live in the 32 KB L1.5 instruction cache).

    python scripts/ncu_hot_footprint.py rep.ncu-rep [thr=episodes of the launch]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
lib = os.path.join(ROOT, "eirgrid_b200", "libeirgrid_b200.so")
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
h = src[1]
ci = h.index("Instructions Executed")
kname = src[0][1]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("episode")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
m_args = re.search(r"eg_episode_kernel<\(bool\)(\d), \(int\)(\d), \(int\)(\d)>|eg_episode_kernel<(\d), (\d), (\d)>", kname)
a_ = [g for g in m_args.groups() if g is not None] if m_args else ["0", "0", "1"]
fn_ok, line, m = False, None, {}
for l in sass.split("\n"):
    s = l.strip()
    if s.startswith(".text."):
        fn_ok = ("eg_episode_kernelILb%dELi%sELi%sEE" % (a_[0] == "1", a_[1], a_[2])) in s
    mm = re.match(r'//## File "(.*)", line (\d+)', s)
    if mm:
        line = int(mm.group(2)) if mm.group(1).endswith("episode.cu") else -(hash(os.path.basename(mm.group(1))) % 1000) - 2
        fname = os.path.basename(mm.group(1))
        continue
    mm = re.match(r"/\*([0-9a-f]{4,})\*/", s)
    if mm and fn_ok:
        m[int(mm.group(1), 16)] = (line, fname if line is not None and line < 0 else None)
text = open(os.path.join(ROOT, "eirgrid_b200", "csrc", "episode.cu")).read().split("\n")
marks = []
for i, l in enumerate(text, 1):
    mm = re.match(r"\s*__device__ .*?\b(\w+)\(.*\{", l) or re.match(r"\s*__global__ .*?\b(\w+)\(", l)
    if mm:
        marks.append((i, mm.group(1)))
def fn(ln):
    name = "?"
    for a, n in marks:
        if a <= ln:
            name = n
    return name
base = int(src[2][0], 16)
hot, cold, dyn = collections.Counter(), collections.Counter(), collections.Counter()
for r in src[2:]:
    ln, f = m.get(int(r[0], 16) - base, (None, None))
    key = f if f else (fn(ln) if ln and ln > 0 else "?")
    c = int(r[ci] or 0)
    (hot if c >= thr else cold)[key] += 1
    dyn[key] += c
th, tc = sum(hot.values()), sum(cold.values())
print("%s\nstatic %d instructions = %.1f KB; executed >= %d times: %d = %.1f KB" % (kname, th + tc, (th + tc) / 64, thr, th, th / 64))
td = sum(dyn.values()) or 1
for k, v in sorted(hot.items(), key=lambda kv: -kv[1]):
    print("  %-34s hot %4d  (+%4d colder)  %5.1f%% of executed" % (k, v, cold.get(k, 0), 100 * dyn[k] / td))
