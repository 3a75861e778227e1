#!/bin/bash
for CFG in "128 8" "128 4" "128 2" "192 4" "256 4" "256 2" "96 4" "64 4"; do
  set -- $CFG
  echo "== B=$1 Q=$2"; PSCL_TRACE=1 PSCL_SEED_BATCH=$1 PSCL_SEED_SPLIT=$2 timeout 300 python tools/time_seed.py 10000 8 100000 2000 2>&1 | grep -E "batches" | sed 's/.*smallest batch 8); //'
done
