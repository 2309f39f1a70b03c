#!/usr/bin/env python
"""Aggregate an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) by device function and by
CUDA source line, using nvdisasm -g -c output of the same cubin for the address -> function/line map.

usage: ncu_by_function.py <source.csv> <nvdisasm.txt> <kernel-substring> [top_lines]
"""
import csv, re, sys
from collections import defaultdict

src_csv, dis_txt, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# ---- nvdisasm map: offset -> (function, line)
fmap = {}
in_k = False; func = None; line = None
for ln in open(dis_txt):
    if ln.startswith("//---") and ".text." in ln:
        in_k = kname in ln
        func = "<entry>"; line = None
        continue
    if not in_k:
        continue
    m = re.match(r"^(\$?[\w$]+):\s*$", ln)
    if m and m.group(1).startswith("$"):
        f = m.group(1)
        f = f.split("$")[-1]
        # demangle lightly: take the identifier before trailing E.. suffix
        m2 = re.search(r"engine_cu_[0-9a-f]+(\d+)([A-Za-z_]\w*)", f)
        func = f
        mm = re.findall(r"(\d+)([A-Za-z_][A-Za-z_0-9]*)", f)
        # pick last plausible name
        names = [n[:int(k)] for k, n in mm if len(n) >= int(k) and int(k) > 2]
        func = names[-1] if names else f
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*)$", ln)
    if m:
        fmap[int(m.group(1), 16)] = (func, line)

rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
byf = defaultdict(lambda: [0, 0, 0]); byl = defaultdict(lambda: [0, 0])
stall_tot = defaultdict(int)
tot_s = tot_e = 0
for r in rows[2:]:
    if len(r) <= iex or not r[ia].startswith("0x"):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    off = a - base
    f, l = fmap.get(off, ("?", None))
    s = int(float(r[isamp] or 0)); e = int(float(r[iex] or 0))
    byf[f][0] += s; byf[f][1] += e; byf[f][2] += 1
    byl[(f, l)][0] += s; byl[(f, l)][1] += e
    tot_s += s; tot_e += e
    for i in stall_cols:
        try: stall_tot[hdr[i]] += int(float(r[i] or 0))
        except ValueError: pass
print("total samples %d, instructions executed %d, static instrs %d" % (tot_s, tot_e, sum(v[2] for v in byf.values())))
print("\nstall samples:")
for k, v in sorted(stall_tot.items(), key=lambda kv: -kv[1])[:10]:
    print("  %-28s %8d  %5.1f%%" % (k, v, 100.0 * v / max(1, sum(stall_tot.values()))))
print("\nby function: samples%  exec%  static-instrs  name")
for f, (s, e, n) in sorted(byf.items(), key=lambda kv: -kv[1][0]):
    print("  %5.1f%%  %5.1f%%  %6d  %s" % (100.0 * s / tot_s, 100.0 * e / tot_e, n, f))
print("\ntop source lines by samples:")
for (f, l), (s, e) in sorted(byl.items(), key=lambda kv: -kv[1][0])[:top]:
    print("  %5.1f%%  %5.1f%%  %-24s %s" % (100.0 * s / tot_s, 100.0 * e / tot_e, f, l))
