#!/bin/bash
# speculative-batch seeding: parity tests, then timings at config 3 and at the config-5 shape for several batch caps
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_freemux_gpu.py -x -q -m gpu > gpurun_out/${TAG}_pytest_fmx.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/${TAG}_pytest_fmx.log
for B in 32 64 128 256 512; do
  PSCL_TRACE=1 PSCL_SEED_BATCH=$B timeout 300 python tools/time_seed.py 10000 8 100000 2000 > gpurun_out/${TAG}_cfg3_B$B.log 2>&1; echo "cfg3 B=$B exit $?"; grep -E "seed|batches" gpurun_out/${TAG}_cfg3_B$B.log
done
for B in 64 128 256; do
  PSCL_TRACE=1 PSCL_SEED_BATCH=$B timeout 300 python tools/time_seed.py 12000 16 500000 4000 > gpurun_out/${TAG}_cfg5_B$B.log 2>&1; echo "cfg5 B=$B exit $?"; grep -E "seed|batches" gpurun_out/${TAG}_cfg5_B$B.log
done
PSCL_SEED_V2=1 timeout 300 python tools/time_seed.py 10000 8 100000 2000 > gpurun_out/${TAG}_cfg3_v2.log 2>&1; grep -E "seed" gpurun_out/${TAG}_cfg3_v2.log
