#!/bin/bash
# N real GPUs (gpurun --gpus N): pscl_multi tests on distinct devices, the torchrun bench line with `strong`, the product path
N=${1:-2}; TAG=${2:-r2m$N}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
IDS=$(seq -s, 0 $((N-1)))
PSCL_TEST_GPUS=$IDS timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_freemux_gpu.py -x -q -m gpu -k "multi or sharded or old_mode_seeding_sharded" > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -5 gpurun_out/${TAG}_pytest_multi.log
for G in $N; do
  timeout 600 python tools/multi_bench.py $G > gpurun_out/${TAG}_multi_api_$G.json 2> gpurun_out/${TAG}_multi_api_$G.err; echo "multi_bench $G exit $?"; cat gpurun_out/${TAG}_multi_api_$G.json | cut -c1-1500
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline \
  > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "torchrun bench exit $?"; tail -3 gpurun_out/${TAG}_bench_${N}gpu.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print("N=%d value %.3g e2e %.3g" % (j["n_gpus"], j["value"], j["e2e"]["value"]))
    for k, v in (j.get("strong") or {}).items():
        print(k, {a: v.get(a) for a in ("ms", "balance", "ms_per_iter", "estep_ms", "allreduce_ms", "seed_ms", "error")})
except Exception as e:
    print("parse failed", e)
PY
