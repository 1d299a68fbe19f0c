#!/usr/bin/env python3
"""Per source line: share of a stall reason's samples (default stall_no_inst) next to instruction share.

    python scripts/ncu_stall_regions.py rep.ncu-rep [stall_column] [top]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
col = sys.argv[2] if len(sys.argv) > 2 else "stall_no_inst"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
lib = os.path.join(ROOT, "eirgrid_b200", "libeirgrid_b200.so")
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
h = src[1]
ci, cc = h.index("Instructions Executed"), h.index(col)
kname = src[0][1]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("episode")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
m_args = re.search(r"eg_episode_kernel<\(bool\)(\d), \(int\)(\d), \(int\)(\d)>|eg_episode_kernel<(\d), (\d), (\d)>", kname)
a_ = [g for g in m_args.groups() if g is not None] if m_args else ["0", "0", "1"]
want_replay, want_wide = a_[0] == "1", a_[1] == "1"
fn_ok, line, m = False, None, {}
for l in sass.split("\n"):
    s = l.strip()
    if s.startswith(".text."):
        fn_ok = ("eg_episode_kernelILb%dELi%sELi%sEE" % (want_replay, a_[1], a_[2])) in s
    mm = re.match(r'//## File "(.*)", line (\d+)', s)
    if mm:
        line = int(mm.group(2)) if mm.group(1).endswith("episode.cu") else -1
        continue
    mm = re.match(r"/\*([0-9a-f]{4,})\*/", s)
    if mm and fn_ok:
        m[int(mm.group(1), 16)] = line
base = int(src[2][0], 16)
agg = collections.defaultdict(lambda: [0, 0])
for r in src[2:]:
    a = agg[m.get(int(r[0], 16) - base)]
    a[0] += int(r[ci] or 0)
    a[1] += int(r[cc] or 0)
ti = sum(a[0] for a in agg.values()) or 1
tc = sum(a[1] for a in agg.values()) or 1
text = open(os.path.join(ROOT, "eirgrid_b200", "csrc", "episode.cu")).read().split("\n")
print("%s by source line (total %d samples)" % (col, tc))
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    t = text[ln - 1].strip()[:90] if ln and ln > 0 else str(ln)
    print("  %5s %5.1f%% of %s  %5.1f%% inst | %s" % (ln, 100 * a[1] / tc, col, 100 * a[0] / ti, t))
