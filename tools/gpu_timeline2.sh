#!/bin/bash
for GO in 0 1; do
  if [ "$GO" = "1" ]; then unset PSCL_SLICE_FULL; else export PSCL_SLICE_FULL=1; fi
  echo "== gaps_only=$GO"; PSCL_TIMELINE=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras 2>&1 >/dev/null | grep timeline | tail -8
done
