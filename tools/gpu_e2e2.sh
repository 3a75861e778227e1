#!/bin/bash
TAG=${1:-r3b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_demux_gpu.py tests/test_fullsize_gpu.py tests/test_golden.py -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest.log
PSCL_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>&1 >/dev/null | grep "pscl_" | tail -4
for S in 1 4 8 15; do
  echo "== PSCL_STAGES=$S"; PSCL_STAGES=$S timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras | python -c "
import sys,json
j=json.loads(sys.stdin.read()); e=j['e2e']; print('e2e %.3g' % e['value'], sorted(e['ms_per_call'])[:6])"
done
