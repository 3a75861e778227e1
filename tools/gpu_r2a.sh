#!/bin/bash
# round-2 visit A: parity of the reworked k_demux_default, geometry variants, seeding time at size
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_demux_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2a_pytest.log
for V in "$@"; do
  for K in auto lane; do
    PSCL_LIB_PATH=$PWD/build/libpscl_$V.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel $K 2> gpurun_out/r2a_$V.$K.err | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('variant $V $K ms_per_step', round(j['ms_per_step'],4), 'kernel_ms', round(j['roofline']['kernel_ms'],4), 'frac', round(j['roofline']['frac'],3), 'e2e', '%.3g' % j['e2e']['value'])"
  done
done
timeout 300 python tools/time_seed.py 10000 8 100000 2000 > gpurun_out/r2a_seed_cfg3.txt 2>&1; cat gpurun_out/r2a_seed_cfg3.txt
timeout 600 python tools/time_seed.py 20000 16 500000 4000 > gpurun_out/r2a_seed_cfg5s.txt 2>&1; cat gpurun_out/r2a_seed_cfg5s.txt
