#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cur.csv python tools/profile_step.py bf16x3 fp32 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python - <<'PY'
import csv, collections, re
lines=[l for l in open('gpurun_out/launches_cur.csv') if not l.startswith('==')]
tot=collections.defaultdict(float); cnt=collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1e6 if u=='ns' else v/1e3 if u=='us' else v
    k=re.sub(r'\(.*','',row['Kernel Name'])[:90]; tot[k]+=v; cnt[k]+=1
T=sum(tot.values())
print(f"total {T:.3f} ms over {sum(cnt.values())} launches; pips kernels {sum(v for k,v in tot.items() if 'pips::' in k):.3f} ms")
for k,v in sorted(tot.items(), key=lambda kv:-kv[1])[:40]:
    print(f"{v:9.3f} ms {100*v/T:5.1f}%  x{cnt[k]:4d}  {k}")
PY
