#!/bin/bash
# per-launch durations of the speculative seeding kernels at config 3
TAG=${1:-r2u}
B=${2:-256}
mkdir -p gpurun_out
PSCL_SEED_BATCH=$B timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"seed3|RadixSort|pair_cell" --csv --log-file gpurun_out/${TAG}_seed3_launches.csv \
  python tools/time_seed.py 10000 8 100000 2000 > gpurun_out/${TAG}_seed3_ncu.log 2>&1; echo "ncu exit $?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_seed3_launches.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
seq = []
for r in rows:
    name = r[4].split("(")[0][:60]; v = float(r[-1].replace(",", "")); unit = r[-2]
    if unit in ("ns", "nsecond"): v /= 1000.0
    elif unit in ("ms", "msecond"): v *= 1000.0
    a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
    seq.append((name, v))
for k, (n, t, m) in agg.items():
    print(f"{k:62s} n={n:5d} total={t/1000:9.3f} ms  max={m:9.1f} us")
print("first 40 seed3 launches:")
for name, v in [x for x in seq if "seed3" in x[0]][:40]: print(f"   {name:50s} {v:9.1f} us")
print("last 12:")
for name, v in [x for x in seq if "seed3" in x[0]][-12:]: print(f"   {name:50s} {v:9.1f} us")
PY
