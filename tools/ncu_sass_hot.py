#!/usr/bin/env python3
"""Hottest SASS region of an `ncu --page source --print-source sass --csv` export (gzip ok): the window of N
consecutive instructions with the most stall samples, printed with samples, executions and the dominant stall reason.
Usage: python tools/ncu_sass_hot.py file_sass.csv.gz [N] [kernel-name-substring]"""
import csv
import gzip
import io
import sys

path = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 120
want = sys.argv[3] if len(sys.argv) > 3 else ""          # substring of the kernel name (an export can hold several launches)
raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
rows = list(csv.reader(io.StringIO(raw)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
pick = next((i for i in starts if want in rows[i][1]), starts[0])
end = next((i for i in starts if i > pick), len(rows))
rows = rows[pick:end]
print("# kernel:", rows[0][1][:100])
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
samp = [int(r[col["# Samples"]] or 0) for r in body]
execs = [int(r[col["Instructions Executed"]] or 0) for r in body]
tot = sum(samp)
best, bi = -1, 0
acc = sum(samp[:N])
best, bi = acc, 0
for i in range(1, max(1, len(body) - N)):
    acc += samp[i + N - 1] - samp[i - 1]
    if acc > best:
        best, bi = acc, i
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"# {path}: {len(body)} instructions, {tot} samples; hottest {N}-instruction window at #{bi} holds {best} samples")
agg = {h: 0 for h in stall_cols}
for r in body:
    for h in stall_cols:
        agg[h] += int(r[col[h]] or 0)
print("# stall samples over the kernel:", {h[6:]: v for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for i in range(bi, min(len(body), bi + N)):
    r = body[i]
    st = {h[6:]: int(r[col[h]] or 0) for h in stall_cols}
    top = max(st.items(), key=lambda kv: kv[1])
    print(f"{i:5d} {samp[i]:6d} {execs[i]:9d}  {r[col['Source']].strip():60s} {top[0] if top[1] else ''}:{top[1] if top[1] else ''}")
