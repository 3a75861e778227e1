#!/bin/bash
# ncu --set full captures of the freemuxlet kernels (E-step at nS 8 and the nS 16 tiles, M-step, posterior, batched seeding)
TAG=${1:-r2j}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:"k_fmx_posterior|k_fmx_estep|k_fmx_mstep" -s 1 -c 3 -o gpurun_out/${TAG}_fmx8 python tools/time_seed.py 10000 8 100000 2000 > gpurun_out/${TAG}_fmx8.log 2>&1; echo "fmx8 exit $?"
timeout 900 $NCU -k regex:"k_fmx_estep" -s 4 -c 4 -o gpurun_out/${TAG}_fmx16 python tools/time_seed.py 12000 16 500000 4000 > gpurun_out/${TAG}_fmx16.log 2>&1; echo "fmx16 exit $?"
timeout 600 $NCU -k regex:"k_fmx_seed_" -s 900 -c 3 -o gpurun_out/${TAG}_seed python tools/time_seed.py 10000 8 100000 2000 > gpurun_out/${TAG}_seed.log 2>&1; echo "seed exit $?"
ls -la gpurun_out/${TAG}_*
