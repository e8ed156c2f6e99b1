"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launches per kernel, shares.
usage: python tools/summarize_launches.py launches.csv "header comment" > summary.txt"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt, order = collections.Counter(), collections.Counter(), []
for r in rows[h + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[mu], 1e-6)
    name = re.sub(r"\(.*", "", r[kn]).strip()
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)[:70]
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
ours = sum(v for k, v in tot.items() if "pips::" in k)
print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# per-launch times are cold-cache and serialised (ncu replays each kernel alone): compare SHARES, not absolutes")
print(f"# total {T:.3f} ms over {sum(cnt.values())} launches; pips_b200 kernels: {ours:.3f} ms ({100 * ours / T:.1f}%)")
for k, v in tot.most_common():
    print(f"{v:9.3f} ms {100 * v / T:5.1f}%  x{cnt[k]:4d}  {k}")
