#!/bin/bash
# times a bench.py workload for several builds of the library (build/libpscl_<v>.so): bash tools/gpu_variants.sh "<bench args>" v1 v2 ...
ARGS=$1; shift
mkdir -p gpurun_out
for V in "$@"; do
  PSCL_LIB_PATH=$PWD/build/libpscl_$V.so timeout 600 python bench.py $ARGS 2> gpurun_out/var_$V.err | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('variant $V ms_per_step', j['ms_per_step'], j.get('fp64',{}).get('frac'), j['roofline'].get('kernel_ms'))"
done
