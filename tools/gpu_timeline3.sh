#!/bin/bash
for CFG in "2 2 1" "2 2 0" "3 3 1" "2 1 1"; do
  set -- $CFG
  if [ "$3" = "1" ]; then export PSCL_SLICE_FULL=1; else unset PSCL_SLICE_FULL; fi
  echo "== slices $1 groups $2 full=$3"; PSCL_TIMELINE=1 PSCL_SLICES=$1 PSCL_GROUPS=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/tl.err | python -c "
import sys,json
j=json.loads(sys.stdin.read()); e=j['e2e']; print('e2e %.3g' % e['value'], sorted(e['ms_per_call'])[:4])"; grep timeline gpurun_out/tl.err | tail -4 | cut -c1-250
done
