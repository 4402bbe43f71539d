"""Condense an `ncu --page source --csv --print-source sass` export: per kernel, the SASS lines with the most sampled
stalls / executed instructions (keeps gpurun_out/ small).  usage: python tools/hot_lines.py source.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
out = []
for r in rows:
    if hdr is None or (len(r) > 3 and r[0] in ("Address", "#")):
        if any("Source" in c for c in r):
            hdr = r
            out.append(("HDR", r))
            continue
    if hdr and len(r) == len(hdr):
        out.append(("ROW", r))
if hdr is None:
    print("no source page found")
    sys.exit(0)
idx = {h: i for i, h in enumerate(hdr)}
samp = next((h for h in hdr if h.startswith("# Samples") or h == "Sampling Data (All)" or "Samples" in h), None)
inst = next((h for h in hdr if "Instructions Executed" in h), None)
src = next(h for h in hdr if "Source" in h)
print("columns:", hdr)
data = [r for k, r in out if k == "ROW"]


def num(r, h):
    try:
        return float(r[idx[h]].replace(",", "")) if h else 0.0
    except ValueError:
        return 0.0


tot_s = sum(num(r, samp) for r in data) or 1.0
tot_i = sum(num(r, inst) for r in data) or 1.0
print(f"rows {len(data)} total samples {tot_s:.0f} total instructions executed {tot_i:.0f}")
top = sorted(range(len(data)), key=lambda i: -num(data[i], samp))[:120]
print("--- top lines by samples (index, samples %, inst %, sass)")
for i in sorted(top):
    r = data[i]
    print(f"{i:6d} {100 * num(r, samp) / tot_s:6.2f} {100 * num(r, inst) / tot_i:6.2f}  {r[idx[src]][:110]}")
