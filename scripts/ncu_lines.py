"""Aggregate the warp-stall samples of an `ncu --page source --print-source cuda,sass --csv` dump per CUDA
source line (the kernel must be compiled with -lineinfo and captured with --import-source on).
usage: python scripts/ncu_lines.py dump.csv [top_n]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
per = defaultdict(lambda: [0, 0, 0, ""])   # samples, not-issued samples, instructions executed
stall_cols = []
stalls = defaultdict(lambda: defaultdict(int))
fname = ""
cur = ("", 0)
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ia, ins, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Warp Stall Sampling (Not-issued Samples)"), hdr.index("Instructions Executed")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
        continue
    if hdr is None:
        continue
    if r[0].isdigit():      # a CUDA line (its own metric cells are '-'); the SASS rows that follow belong to it
        cur = (fname, int(r[0]))
        per[cur][3] = ",".join(r[1:len(r) - len(hdr) + 2]).strip()[:110]
        continue
    if len(r) != len(hdr) or r[0] != "" or not r[2].startswith("0x"):
        continue
    def num(x):
        try: return int(float(x))
        except ValueError: return 0
    per[cur][0] += num(r[ia]); per[cur][1] += num(r[ins]); per[cur][2] += num(r[ie])
    for i, h in stall_cols:
        stalls[cur][h] += num(r[i])
tot = sum(v[0] for v in per.values()) or 1
print("total samples", tot)
for key, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    s = sorted(stalls[key].items(), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% %-14s:%4d inst=%8d  %-40s | %s" % (100.0 * v[0] / tot, key[0], key[1], v[2], " ".join("%s=%d" % (h[6:], n) for h, n in s if n), v[3]))
