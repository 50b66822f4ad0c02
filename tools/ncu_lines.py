"""Joins an ncu --page source (SASS) csv with nvdisasm line info of the SAME build and ranks source lines.

    python tools/ncu_lines.py <report.ncu-rep> <launch_skip> <kernel-substring-in-mangled-name> [top]
Prints per source line: share of issued warp instructions, average active lanes, share of stall samples.
"""
import collections, csv, io, re, subprocess, sys, tempfile, os
rep, skip, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
print(rows[hi - 1][:2])
h = rows[hi]
ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) >= len(h) - 1 and r[0].startswith("0x")]
base = int(data[0][0], 16)
uniq = {}
for r in data:
    uniq.setdefault(int(r[0], 16) - base, r)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(REPO, "luisarender_b200/lib/libb200pt.so")], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
name = None; cur = None; lineof = {}; opof = {}
for l in sass.splitlines():
    if l.startswith(".text."):
        name = l; cur = None; continue
    if name is None or ksub not in name: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m: lineof[int(m.group(1), 16)] = cur; opof[int(m.group(1), 16)] = m.group(2)
mism = sum(1 for off, r in uniq.items() if off not in opof or r[ix["Source"]].split()[0].split(".")[0] not in opof[off])
print("instructions", len(uniq), "mismatching", mism)
I = lambda r, n: int(r[ix[n]])
tot = sum(I(r, "Instructions Executed") for r in uniq.values()); ts = sum(I(r, "# Samples") for r in uniq.values())
tthr = sum(I(r, "Thread Instructions Executed") for r in uniq.values())
print("warp instructions", tot, "avg lanes", round(tthr / tot, 2))
by = collections.defaultdict(lambda: [0, 0, 0])
for off, r in uniq.items():
    k = lineof.get(off)
    b = by[k]; b[0] += I(r, "Instructions Executed"); b[1] += I(r, "Thread Instructions Executed"); b[2] += I(r, "# Samples")
for k, b in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(k):34s} inst {b[0] / tot * 100:5.2f}%  lanes {b[1] / max(b[0], 1):5.1f}  samples {b[2] / ts * 100:5.2f}%")
