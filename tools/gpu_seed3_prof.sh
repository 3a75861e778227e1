#!/bin/bash
TAG=${1:-r2x}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:"k_fmx_seed3_delta" -s 40 -c 1 -o gpurun_out/${TAG}_delta python tools/time_seed.py 10000 8 100000 2000 > gpurun_out/${TAG}_delta.log 2>&1; echo "delta exit $?"
python tools/ncu_summary.py gpurun_out/${TAG}_delta.ncu-rep 0 > gpurun_out/${TAG}_delta_ncu.txt 2>&1; rm -f gpurun_out/${TAG}_delta.ncu-rep
timeout 600 $NCU -k regex:"k_fmx_seed3_commit" -s 20 -c 1 -o gpurun_out/${TAG}_commit python tools/time_seed.py 10000 8 100000 2000 > gpurun_out/${TAG}_commit.log 2>&1; echo "commit exit $?"
python tools/ncu_summary.py gpurun_out/${TAG}_commit.ncu-rep 0 > gpurun_out/${TAG}_commit_ncu.txt 2>&1; rm -f gpurun_out/${TAG}_commit.ncu-rep
timeout 600 $NCU -k regex:"k_fmx_seed3_eval0" -s 20 -c 1 -o gpurun_out/${TAG}_eval0 python tools/time_seed.py 10000 8 100000 2000 > gpurun_out/${TAG}_eval0.log 2>&1; echo "eval0 exit $?"
python tools/ncu_summary.py gpurun_out/${TAG}_eval0.ncu-rep 0 > gpurun_out/${TAG}_eval0_ncu.txt 2>&1; rm -f gpurun_out/${TAG}_eval0.ncu-rep
