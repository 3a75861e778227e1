#!/bin/bash
# A/B GPU visit for the demuxlet kernels: parity tests of the demuxlet path, then bench.py with each kernel.
# Usage (under gpurun): bash tools/gpu_ab.sh <tag>
TAG=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_demux_gpu.py -x -q -m gpu > gpurun_out/${TAG}_pytest_demux.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_demux.log
tail -15 gpurun_out/${TAG}_pytest_demux.log
for K in lane cls; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel $K > gpurun_out/${TAG}_bench_$K.json 2> gpurun_out/${TAG}_bench_$K.err; echo "bench $K exit $?"
  cat gpurun_out/${TAG}_bench_$K.json; tail -3 gpurun_out/${TAG}_bench_$K.err
done
