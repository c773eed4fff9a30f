#!/usr/bin/env python
"""Summarise an ncu source page of k_pw_umma by warp role (producer / MMA issuer / epilogue) using SASS markers."""
import csv, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
seen = set(); u = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[idx['Address']] in seen: continue
    seen.add(r[idx['Address']]); u.append(r)
def iv(r, c):
    try: return int(float(r[idx[c]] or 0))
    except: return 0
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(iv(r, '# Samples') for r in u)
print("instructions", len(u), "samples", tot)
# top 25 instructions
for r in sorted(u, key=lambda r: -iv(r, '# Samples'))[:25]:
    i = u.index(r)
    st = sorted(((iv(r, c), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {iv(r,'# Samples'):6d} {100*iv(r,'# Samples')/max(tot,1):5.1f}% exec={iv(r,'Instructions Executed'):9d} {r[idx['Source']][:64]:64s} {[(c,v) for v,c in st if v]}")
# regions of 50
print("--- by region (index, samples, warp-instructions executed)")
for i in range(0, len(u), 50):
    s = sum(iv(r, '# Samples') for r in u[i:i+50]); e = sum(iv(r, 'Instructions Executed') for r in u[i:i+50])
    marks = set()
    for r in u[i:i+50]:
        for m in ('UBLKCP', 'UTCHMMA', 'LDTM', 'STG', 'LDG', 'STS', 'RED', 'F2F', 'CCTL', 'BAR.SYNC', 'SYNCS.PHASECHK', 'FENCE', 'UTCBAR', 'LDS'):
            if m in r[idx['Source']]: marks.add(m)
    if s or e: print(f"{i:5d} {s:6d} {e:10d} {sorted(marks)}")
