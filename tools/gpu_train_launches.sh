#!/bin/bash
# per-kernel launch list of the Match-Tensor training step (cfg2 shape)
mkdir -p gpurun_out
python tools/train_timing.py 2>/dev/null | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/train_launches.csv python tools/train_timing.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/train_launches.csv')))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
ki, vi = rows[h].index('Kernel Name'), rows[h].index('Metric Value')
per = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) > vi:
        a = per.setdefault(r[ki][:70], [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(',', ''))
tot = sum(v[1] for v in per.values())
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:30]:
    print('%-70s n=%4d %10.1f us/iter %5.1f%%' % (k, v[0], v[1] / 8 / 1e3, 100 * v[1] / tot))
print('total per iteration: %.2f ms' % (tot / 8 / 1e6))
PY
