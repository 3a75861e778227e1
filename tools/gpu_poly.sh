#!/bin/bash
# GPU visit for k_demux_poly: bench (config-4 shape) + launch list + one full ncu capture
TAG=${1:-poly}
mkdir -p gpurun_out
timeout 600 python bench.py --workload demux64 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_demux64.json 2> gpurun_out/${TAG}_bench_demux64.err; echo "bench exit $?"; cat gpurun_out/${TAG}_bench_demux64.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_demux_poly -s 2 -c 1 -f -o gpurun_out/${TAG}_prof \
  python bench.py --workload demux64 --cells 64 --steps 1 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"
