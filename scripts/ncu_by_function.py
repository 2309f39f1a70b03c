#!/usr/bin/env python
"""Aggregate an ncu SASS source page by device function.

usage: ncu_by_function.py <source.csv from `ncu -i X.ncu-rep --page source --csv`> <nvdisasm -g -c output of the same
       cubin> <kernel substring> <n_events>
Prints per function: share of stall samples, instructions per event, hot code size (lines executed at least 0.004
times per event), top stall reasons."""
import csv, re, subprocess, sys
from collections import defaultdict
src_csv, dis_txt, kname, n_events = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
fmap = {}; in_k = False; func = "<entry>"
for ln in open(dis_txt):
    if ln.startswith("//---") and ".text." in ln:
        in_k = kname in ln; func = "<entry>"; continue
    if not in_k: continue
    m = re.match(r"^(\$[\w$]+):\s*$", ln)
    if m: func = m.group(1).split("$")[-1]; continue
    m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*)$", ln)
    if m: fmap[int(m.group(1), 16)] = func
names = sorted(set(fmap.values()))
dem = dict(zip(names, subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")))
def short(f):
    d = dem.get(f, f); d = re.sub(r"_INTERNAL_\w+::|\(anonymous namespace\)::", "", d); return d.split("(")[0]
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None; byf = defaultdict(lambda: [0, 0]); st = defaultdict(int); byf_st = defaultdict(lambda: defaultdict(int))
lines = defaultdict(int); linef = {}
ts = te = 0
for r in rows[2:]:
    if len(r) <= iex or not r[ia].startswith("0x"): continue
    a = int(r[ia], 16)
    if base is None: base = a
    f = short(fmap.get(a - base, "?"))
    s = int(float(r[isamp] or 0)); e = int(float(r[iex] or 0))
    byf[f][0] += s; byf[f][1] += e; ts += s; te += e
    lines[(a - base) // 128] += e; linef[(a - base) // 128] = f
    for i in stall:
        v = int(float(r[i] or 0)); st[hdr[i]] += v; byf_st[f][hdr[i]] += v
hot = defaultdict(int)
for l, e in lines.items():
    if e / 8 / n_events >= 0.004: hot[linef[l]] += 128
print("samples %d, instructions executed %d = %.0f per event; hot code %.1f KB" % (ts, te, te / n_events, sum(hot.values()) / 1024))
tot = sum(st.values())
for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]: print("  %-26s %5.1f%%" % (k, 100 * v / tot))
print("function: samples%  instr/event  hotKB | top stalls")
for f, (s, e) in sorted(byf.items(), key=lambda kv: -kv[1][0])[:45]:
    tops = sorted(byf_st[f].items(), key=lambda kv: -kv[1])[:3]; tt = sum(byf_st[f].values()) or 1
    print("  %5.1f%% %7.1f %5.1f  %-30s %s" % (100 * s / ts, e / n_events, hot[f] / 1024, f, " ".join("%s=%.0f%%" % (k[6:], 100 * v / tt) for k, v in tops)))
