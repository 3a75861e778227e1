#!/bin/bash
# GPU visit for k_demux_cls: demuxlet parity tests, bench (ws), one full ncu capture of the kernel.
# Usage (under gpurun): bash tools/gpu_ws.sh <tag> [kernel-regex]
TAG=${1:-ws}
KRE=${2:-k_demux_cls}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_demux_gpu.py -x -q -m gpu > gpurun_out/${TAG}_pytest_demux.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_demux.log
tail -5 gpurun_out/${TAG}_pytest_demux.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel cls > gpurun_out/${TAG}_bench_cls.json 2> gpurun_out/${TAG}_bench_cls.err; echo "bench cls exit $?"
cat gpurun_out/${TAG}_bench_cls.json; tail -3 gpurun_out/${TAG}_bench_cls.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 3 -c 1 -f -o gpurun_out/${TAG}_prof \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --kernel cls > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out | tail -8
