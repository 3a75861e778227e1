#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list and one full capture of the top kernel.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [kernel-regex]
TAG=${1:-r01}
KRE=${2:-k_demux_default}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
cat gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 3 -c 1 -f -o gpurun_out/${TAG}_prof \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"
# the other bench lines of a full round (skip with FAST=1)
if [ -z "$FAST" ]; then
  timeout 120 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/${TAG}_smoke.log
  timeout 300 python bench.py --workload freemux --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_freemux.json 2> gpurun_out/${TAG}_bench_freemux.err; echo "freemux exit $?"
  timeout 300 python bench.py --workload demux64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_demux64.json 2> gpurun_out/${TAG}_bench_demux64.err; echo "demux64 exit $?"
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err; echo "reference arm exit $?"
fi
ls -la gpurun_out
