#!/bin/bash
timeout 900 python -m pytest tests/test_demux_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu -k "pipelined or tiny or compact or config2" 2>&1 | tail -3
for CFG in "6 3 r" "6 3 g" "6 2 r" "8 4 r" "4 2 r"; do
  set -- $CFG
  if [ "$3" = "r" ]; then export PSCL_SLICE_READS=1; else unset PSCL_SLICE_READS; fi
  echo "== slices $1 groups $2 mode=$3"; PSCL_TIMELINE=1 PSCL_SLICES=$1 PSCL_GROUPS=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/tl.err | python -c "
import sys,json
j=json.loads(sys.stdin.read()); e=j['e2e']; print('e2e %.3g' % e['value'], sorted(e['ms_per_call'])[:4])"; grep timeline gpurun_out/tl.err | tail -4 | cut -c1-230
done
