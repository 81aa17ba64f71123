#!/usr/bin/env python3
"""Per-source-line / per-function profile of one kernel from an .ncu-rep (--set full, -lineinfo build).
Joins the SASS page of the report (instructions executed + stall samples per address) with nvdisasm's
line table of the same cubin.  usage: ncu_lines.py <report.ncu-rep> <cubin> <source.cu> [top]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, src = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
import os
HELPER_END = 0 if os.environ.get("NCU_LINES_INNERMOST") else 70   # source lines below this are one-line arithmetic helpers (charged to their call site)

# --- line table: kernel text offset -> source line ------------------------------------------------
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
sections = {}   # section name -> {offset: (line charged, sass text, [lines of the whole inline chain, innermost first])}
cur, chain = None, []
mine = src.split("/")[-1]
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        cur = sections.setdefault(m.group(1), {}); chain = []; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if chain is None:
            chain = []
        chain.append(int(m.group(2)) if m.group(1).endswith(mine) else -1)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur is not None:
        if chain is None:              # an instruction without its own line comment inherits the previous one
            chain = last_chain
        last_chain = chain
        # charge the innermost frame that is not a one-line helper / toolkit header
        line = next((x for x in chain if x >= HELPER_END), chain[-1] if chain else 0)
        cur[int(m.group(1), 16)] = (line, m.group(2).strip(), chain)
        chain = None
last_chain = []
# --- function extents of the source (column-0 definitions) ------------------------------------------
text = open(src).read().splitlines()
func_of = [None] * (len(text) + 2)
name, depth_start = None, None
for i, l in enumerate(text, 1):
    if name is None:
        m = re.match(r"^(?:template.*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__|inline|__forceinline__|__noinline__|[\w:<>\*&]+\s)+.*?(\w+)\s*\(", l)
        if m and not l.startswith((" ", "\t", "//", "#", "}")) and not l.rstrip().endswith(";"):
            name = m.group(1)
    if name is not None:
        func_of[i] = name
        if l.startswith("}"):
            name = None

# --- report rows ------------------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
kname = re.search(r'"Kernel Name","([^"(]+)', raw[0]).group(1)
rows = list(csv.reader(raw[1:]))
hdr = rows[0]
ia, ii, isamp, ithr = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
stall_cols = [(k, h) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = rows[1:]
base = int(body[0][ia], 16)
want = os.environ.get("NCU_SECTION")
sec = (next((v for k, v in sections.items() if want in k), None) if want else None) or next((v for k, v in sections.items() if "render_kernel" in k and kname.split("::")[-1] in k), None) or max(sections.values(), key=len)

by_line = defaultdict(lambda: [0, 0, 0])
by_func = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
incl = defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for r in body:
    off = int(r[ia], 16) - base
    ent = sec.get(off, (0, "", []))
    ln = ent[0]
    n, s, t = int(r[ii] or 0), int(r[isamp] or 0), int(r[ithr] or 0)
    for fn in {func_of[x] for x in ent[2] if 0 < x < len(func_of) and func_of[x]}:
        incl[fn][0] += n; incl[fn][1] += s
    tot_i += n; tot_s += s
    by_line[ln][0] += n; by_line[ln][1] += s; by_line[ln][2] += t
    f = func_of[ln] if 0 < ln < len(func_of) and func_of[ln] else f"<line {ln}>"
    by_func[f][0] += n; by_func[f][1] += s; by_func[f][2] += t
    for k, h in stall_cols:
        v = int(r[k] or 0)
        if v:
            by_func[f][3][h] += v

print(f"kernel {kname}: {tot_i} warp instructions, {tot_s} samples, {len(body)} SASS instructions")
print("\nby function (innermost non-helper inlined frame):  inst%  samples%  lanes  top stalls")
for f, (n, s, t, st) in sorted(by_func.items(), key=lambda kv: -kv[1][1])[:top]:
    tops = ", ".join(f"{h[6:]} {100.0 * v / max(s, 1):.0f}%" for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(f"  {f:34s} {100.0 * n / tot_i:6.2f} {100.0 * s / tot_s:7.2f}  {t / max(n, 1):5.1f}  {tops}")
print("\ninclusive (every function on the inline chain):  inst%  samples%")
for f, (n, s) in sorted(incl.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"  {f:34s} {100.0 * n / tot_i:6.2f} {100.0 * s / tot_s:7.2f}")
print("\nby line:  line  inst%  samples%  source")
for ln, (n, s, t) in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:top]:
    code = text[ln - 1].strip()[:110] if 0 < ln <= len(text) else ""
    print(f"  {ln:5d} {100.0 * n / tot_i:6.2f} {100.0 * s / tot_s:7.2f}  {code}")
