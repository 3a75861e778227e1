#!/bin/bash
# round-2 visit B: all GPU tests (incl. pscl_multi on a repeated device), bench line with strong/extra, variants, ncu
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench.json").read())
    print("value %.3g ms_per_step %.4f kernel_ms %.4f frac %.3f e2e %.3g" % (j["value"], j["ms_per_step"], j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["e2e"]["value"]))
    print("e2e totals", j["e2e"]["repeat_totals_ms"])
    print(json.dumps(j.get("strong"), indent=1)[:3000])
    print(json.dumps(j.get("extra"), indent=1)[:2000])
except Exception as e:
    print("bench parse failed", e)
PY
for V in r8 n384r; do
  for K in auto lane; do
    PSCL_LIB_PATH=$PWD/build/libpscl_$V.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --kernel $K 2> gpurun_out/${TAG}_$V.$K.err | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('variant $V $K ms_per_step', round(j['ms_per_step'],4), 'kernel_ms', round(j['roofline']['kernel_ms'],4), 'frac', round(j['roofline']['frac'],3), 'e2e', '%.3g' % j['e2e']['value'])"
  done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_demux_default -s 3 -c 1 -f -o gpurun_out/${TAG}_prof \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_demux_default -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_lane \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --kernel lane > gpurun_out/${TAG}_ncu_full_lane.log 2>&1; echo "ncu lane exit $?"
# freemuxlet kernels: one capture each of the E-step (nS 8), M-step, seeding, posterior
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fmx_estep|k_fmx_mstep|k_fmx_seed" -c 6 -f -o gpurun_out/${TAG}_prof_fmx \
  python tools/time_seed.py 10000 8 100000 2000 > gpurun_out/${TAG}_ncu_fmx.log 2>&1; echo "ncu fmx exit $?"
ls -la gpurun_out | tail -20
