mkdir -p gpurun_out
python tools/train_timing.py 128 > gpurun_out/train_timing.log 2>&1
python tools/train_timing.py 128 fix >> gpurun_out/train_timing.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_timing.py 128 > gpurun_out/ncu_train.log 2>&1
cat gpurun_out/train_timing.log
python - <<'PY'
import csv, collections
rows = list(csv.DictReader(l for l in open('gpurun_out/train_launches.csv') if l.startswith('"')))
agg = collections.defaultdict(float); cnt = collections.Counter()
for r in rows:
    k = r['Kernel Name'][:60]; agg[k] += float(r['Metric Value']) ; cnt[k] += 1
tot = sum(agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:22]:
    print('%-62s %5d launches %10.1f us/step %5.1f%%' % (k, cnt[k], v / 8 / 1e3, 100 * v / tot))
PY
