"""Aggregate the ncu source page (SASS samples) by CUDA source line using nvdisasm -g line info.
usage: prof_by_line.py <src.csv from `ncu --page source --csv`> <nvdisasm -g -c listing> <mangled kernel name>"""
import collections
import csv
import re
import sys

src_csv, sass, kname = sys.argv[1:4]
txt = open(sass).read()
hdr = "//--------------------- .text.%s " % kname
start = txt.index(hdr)
end = txt.find("//--------------------- .", start + 10)
sec = txt[start:end if end > 0 else len(txt)]
cur = None
seq = []
for ln in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        seq.append((int(m.group(1), 16), cur, m.group(2).strip()))
rows = list(csv.reader(open(src_csv)))
H = rows[1]
data = [r for r in rows[2:] if len(r) > 5 and r[0].startswith("0x")]
# only the first kernel block
first = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) > 5 and r[0].startswith("0x"):
        first.append(r)
data = first
iS, iI, iA, iT = H.index("# Samples"), H.index("Instructions Executed"), H.index("Address"), H.index("Avg. Threads Executed")
base = int(data[0][iA], 16)
byoff = {int(r[iA], 16) - base: (int(r[iS]), int(r[iI])) for r in data}
agg = collections.defaultdict(lambda: [0, 0])
for off, loc, ins in seq:
    if off in byoff:
        s, i = byoff[off]
        agg[loc][0] += s
        agg[loc][1] += i
tot = sum(v[0] for v in agg.values())
toti = sum(v[1] for v in agg.values())
print("sass instrs %d, profiled %d, samples %d, inst %d" % (len(seq), len(data), tot, toti))
src = {}
import os
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csolve_b200", "csrc")
for fn in ("kernels.cu", "contract.cuh"):
    src[fn] = open(os.path.join(root, fn)).read().splitlines()
N = int(sys.argv[4]) if len(sys.argv) > 4 else 40
for loc, v in sorted(agg.items(), key=lambda x: -x[1][0])[:N]:
    fn, l = loc if loc else ("?", 0)
    text = src[fn][l - 1].strip()[:90] if fn in src and 0 < l <= len(src[fn]) else ""
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100 * v[0] / tot, 100 * v[1] / toti, fn, l, text))
