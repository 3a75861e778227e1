#!/bin/bash
# ncu --set full captures of the freemuxlet kernels (E-step at nS 8 and the nS 16 tiles, M-step, posterior, batched seeding).
# The reports are summarised on the box (tools/ncu_summary.py) and deleted: three of them exceed what gpurun brings back.
TAG=${1:-r2j}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
summ() {  # summ <name> <launches>
  for i in $(seq 0 $(($2 - 1))); do
    python tools/ncu_summary.py gpurun_out/${TAG}_$1.ncu-rep $i > gpurun_out/${TAG}_$1_k$i.txt 2>&1
  done
  rm -f gpurun_out/${TAG}_$1.ncu-rep
}
timeout 600 $NCU -k regex:"k_fmx_posterior|k_fmx_estep|k_fmx_mstep" -s 1 -c 3 -o gpurun_out/${TAG}_fmx8 python tools/time_seed.py 10000 8 100000 2000 > gpurun_out/${TAG}_fmx8.log 2>&1; echo "fmx8 exit $?"
summ fmx8 3
timeout 900 $NCU -k regex:"k_fmx_estep" -s 4 -c 4 -o gpurun_out/${TAG}_fmx16 python tools/time_seed.py 12000 16 500000 4000 > gpurun_out/${TAG}_fmx16.log 2>&1; echo "fmx16 exit $?"
summ fmx16 4
timeout 600 $NCU -k regex:"k_fmx_seed_" -s 900 -c 3 -o gpurun_out/${TAG}_seed python tools/time_seed.py 10000 8 100000 2000 > gpurun_out/${TAG}_seed.log 2>&1; echo "seed exit $?"
summ seed 3
ls -la gpurun_out/${TAG}_*
