#!/bin/bash
for CFG in "6 2 0" "4 2 0" "5 2 0" "3 3 0" "4 4 0" "6 2 1" "4 2 1"; do
  set -- $CFG
  if [ "$3" = "1" ]; then unset PSCL_SLICE_FULL; else export PSCL_SLICE_FULL=1; fi
  echo "== slices $1 groups $2 gaps_only=$3"; PSCL_SLICES=$1 PSCL_GROUPS=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); e=j['e2e']; print('e2e %.3g' % e['value'], sorted(e['ms_per_call'])[:4])"
done
