#!/bin/bash
# where does an end-to-end pscl_demux_run call spend its time: phase trace + per-kernel durations
TAG=${1:-r2g}
mkdir -p gpurun_out
PSCL_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>&1 >/dev/null | grep "pscl_" | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches exit $?"
python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/${TAG}_launches.csv")))
hdr=None; d=collections.defaultdict(list)
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        try: d[r[hdr.index('Kernel Name')][:64]].append(float(r[hdr.index('Metric Value')].replace(',','')))
        except: pass
for k,v in d.items(): print(f"{k:66s} n={len(v):4d} mean={sum(v)/len(v)/1e3:9.1f} us")
PY
for S in 1 4 15; do
  echo "== PSCL_STAGES=$S"; PSCL_STAGES=$S timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras | python -c "
import sys,json
j=json.loads(sys.stdin.read()); e=j['e2e']; print('e2e %.3g' % e['value'], sorted(e['ms_per_call'])[:5])"
done
