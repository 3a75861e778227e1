#!/bin/bash
TAG=${1:-r3u}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_demux_gpu.py tests/test_fullsize_gpu.py tests/test_golden.py tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/${TAG}_pytest.log
for CFG in "6 3 0" "6 3 1" "8 4 0" "6 2 0" "9 3 0" "12 4 0"; do
  set -- $CFG
  echo "== PSCL_SLICES=$1 PSCL_GROUPS=$2 gaps_only=$3"
  if [ "$3" = "1" ]; then unset PSCL_SLICE_FULL; else export PSCL_SLICE_FULL=1; fi
  PSCL_SLICES=$1 PSCL_GROUPS=$2 PSCL_TIMELINE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/tl.err | python -c "
import sys,json
j=json.loads(sys.stdin.read()); e=j['e2e']; print('e2e %.3g' % e['value'], sorted(e['ms_per_call'])[:6])"; grep timeline gpurun_out/tl.err | sort -t'|' -k10 | head -2
done
