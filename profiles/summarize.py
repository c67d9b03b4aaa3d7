"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.
usage: python profiles/summarize.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows:
    if r is hdr or len(r) <= max(ik, iv, im) or r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("sgm::tc::<unnamed>::", "").replace("sgm::<unnamed>::", "")
    try:
        t = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    tot[name] += t
    cnt[name] += 1
total = sum(tot.values())
print(f"| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"| `{k[:90]}` | {cnt[k]} | {v / 1e3:.1f} | {100 * v / total:.1f}% |")
print(f"\ntotal {total / 1e3:.1f} us over {sum(cnt.values())} launches (ncu per-launch times are cold-cache and serialised: compare shares)")
