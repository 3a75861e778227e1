#!/bin/bash
# product multi-GPU path on N real GPUs: tests + pscl_multi strong scaling (1 and N GPUs)
N=${1:-2}; TAG=${2:-r2n$N}
mkdir -p gpurun_out
IDS=$(seq -s, 0 $((N-1)))
PSCL_TEST_GPUS=$IDS timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_freemux_gpu.py -x -q -m gpu -k "multi or sharded or old_mode" > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -4 gpurun_out/${TAG}_pytest_multi.log
for G in 1 $N; do
  timeout 900 python tools/multi_bench.py $G > gpurun_out/${TAG}_multi_api_$G.json 2> gpurun_out/${TAG}_multi_api_$G.err; echo "multi_bench $G exit $?"; cut -c1-1800 gpurun_out/${TAG}_multi_api_$G.json; tail -3 gpurun_out/${TAG}_multi_api_$G.err
done
