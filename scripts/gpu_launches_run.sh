#!/bin/bash
# per-kernel time of the whole W5 run() (ncu time-only pass, after a plain run exits 0)
mkdir -p gpurun_out
timeout 300 python scripts/profile_run.py > gpurun_out/plain_run.log 2>&1 || { tail gpurun_out/plain_run.log; exit 1; }
cat gpurun_out/plain_run.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_run.csv python scripts/profile_run.py > gpurun_out/ncu_run.log 2>&1
tail -2 gpurun_out/ncu_run.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_run.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u in ('us', 'usecond') else (v / 1e6 if u in ('ns', 'nsecond') else v)
    n = r[ki].split('(')[0]; tot[n] += v; cnt[n] += 1
for n, v in sorted(tot.items(), key=lambda x: -x[1]): print('%-40s %6d launches %9.3f ms' % (n[:40], cnt[n], v))
print('total', sum(tot.values()))
PY
