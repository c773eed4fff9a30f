#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel share table (markdown) for profiles/."""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
tot = 0.0; agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) != len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum": continue
    name = r[idx["Kernel Name"]]
    name = re.sub(r"\(.*\)$", "", name).replace("void ", "").replace("<unnamed>::", "")
    v = float(r[idx["Metric Value"]].replace(",", ""))
    unit = r[idx["Metric Unit"]]
    us = v / 1e3 if unit.startswith("n") else (v if unit.startswith("u") else v * 1e3)
    agg[name][0] += 1; agg[name][1] += us; tot += us
print(f"| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {us:.1f} | {100*us/tot:.1f}% | {us/n:.1f} |")
print(f"\ntotal {tot/1e3:.3f} ms over {sum(n for n, _ in agg.values())} launches (cold-cache, serialised by ncu: compare SHARES)")
