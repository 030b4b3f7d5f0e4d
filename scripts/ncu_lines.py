"""Stall samples of one kernel of an .ncu-rep per SOURCE LINE: the report's SASS page (ncu --page source --csv) is joined
with the line table of the shipped cubin (nvdisasm -g).  Build container:  python scripts/ncu_lines.py REP.ncu-rep CUBIN_NAME [top]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, cub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "two-towers-overlords_b200", "lib", "libtt_b200.so")], cwd=tmp,
               capture_output=True)
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f"{cub}.sm_100a.cubin")], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
kernel = rows[0][1]
hdr = rows[1]
iA, iS = hdr.index("Address"), hdr.index("# Samples")
# the function's section in the disassembly
fn = re.search(r"::(\w+)(?:<.*>)?\(", kernel).group(1)
cur, off2line, active = None, {}, False
for ln in sass.split("\n"):
    if ln.startswith("//---") and ".text." in ln:
        active = fn in ln
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        off2line[int(m.group(1), 16)] = cur
data = rows[2:]
base = int(data[0][iA], 16)
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not" not in h]
agg, per, tot = collections.Counter(), collections.defaultdict(collections.Counter), 0
for r in data:
    s = int(r[iS] or 0)
    tot += s
    k = off2line.get(int(r[iA], 16) - base)
    agg[k] += s
    for i in stall:
        v = int(r[i] or 0)
        if v:
            per[k][hdr[i][6:]] += v
print(f"{kernel}: {tot} samples")
cache = {}
for k, s in agg.most_common(top):
    text = ""
    if k:
        for d in (os.path.join(ROOT, "two-towers-overlords_b200", "csrc"),):
            f = os.path.join(d, k[0])
            if os.path.exists(f):
                cache.setdefault(f, open(f).read().split("\n"))
                text = cache[f][k[1] - 1].strip()[:100]
    print(f"{s:6d} {100 * s / tot:5.1f}%  {k[0] if k else '?'}:{k[1] if k else 0}  {dict(per[k].most_common(3))}  {text}")
