#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg1.csv python tools/profile_cfg1.py > gpurun_out/ncu_cfg1.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections, re
lines=[l for l in open('gpurun_out/launches_cfg1.csv') if not l.startswith('==')]
tot=collections.defaultdict(float); cnt=collections.Counter(); grid={}
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1e6 if u=='ns' else v/1e3 if u=='us' else v
    k=re.sub(r'\(.*','',row['Kernel Name'])[:60]+' grid='+row.get('Grid Size','?'); tot[k]+=v; cnt[k]+=1
T=sum(tot.values())
print(f"total {T:.3f} ms over {sum(cnt.values())} launches")
for k,v in sorted(tot.items(), key=lambda kv:-kv[1])[:45]:
    print(f"{v:9.3f} ms {100*v/T:5.1f}%  x{cnt[k]:4d}  avg {1e3*v/cnt[k]:7.1f} us  {k}")
PY
