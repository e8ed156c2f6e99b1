"""profiles/sass_summary.txt: per kernel of libpips_b200.so the counts of the SASS mnemonics that prove a Blackwell-native
path (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tile load, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops),
from `cuobjdump -sass` (runs without a GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pips_b200 import _build  # noqa: E402

WANT = ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTCBAR", "SYNCS", "UTCATOM", "FFMA2", "MUFU", "SHFL", "STG", "LDG", "ST.E", "BAR")
lib = _build.build()
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern, counts, order = None, {}, []
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").split("(")[0]
        counts[kern] = collections.Counter()
        order.append(kern)
        continue
    if kern is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        op = m.group(1)
        counts[kern]["_total"] += 1
        for w in WANT:
            if op == w or op.startswith(w + "."):
                counts[kern][w] += 1
cols = [w for w in WANT if any(c[w] for c in counts.values())]
lines = [f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} (sm_100a), source digest {_build.source_digest()[:16]}",
         "# per kernel: static instruction counts of the mnemonics that show tcgen05 / TMEM / TMA / mbarrier use",
         f"{'kernel':58s} {'instr':>7s} " + " ".join(f"{c:>8s}" for c in cols)]
for k in order:
    c = counts[k]
    lines.append(f"{k[:58]:58s} {c['_total']:7d} " + " ".join(f"{c[w]:8d}" for w in cols))
text = "\n".join(lines) + "\n"
dst = os.path.join(ROOT, "profiles", "sass_summary.txt")
open(dst, "w").write(text)
print(text)
