#!/bin/bash
# N-GPU visit (gpurun --gpus N): torchrun bench lines for the three workloads + the reference arm
N=${1:-2}; TAG=${2:-multi}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_demux_${N}gpu.json 2> gpurun_out/${TAG}_demux.err; echo "demux exit $?"; tail -1 gpurun_out/${TAG}_bench_demux_${N}gpu.json
timeout 900 $TR bench.py --gpus $N --workload demux64 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_demux64_${N}gpu.json 2> gpurun_out/${TAG}_demux64.err; echo "demux64 exit $?"; tail -1 gpurun_out/${TAG}_bench_demux64_${N}gpu.json
timeout 900 $TR bench.py --gpus $N --workload freemux --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_freemux_${N}gpu.json 2> gpurun_out/${TAG}_freemux.err; echo "freemux exit $?"; tail -1 gpurun_out/${TAG}_bench_freemux_${N}gpu.json
