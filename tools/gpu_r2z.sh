#!/bin/bash
# all GPU tests, bench line (with extras), loader timing at config 2 (files -> files), seeding timings
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/${TAG}_pytest.log
for B in 96 128 192; do
  PSCL_TRACE=1 PSCL_SEED_BATCH=$B timeout 300 python tools/time_seed.py 10000 8 100000 2000 2>&1 | grep -E "batches|seed \(greedy\)"
done
PSCL_TRACE=1 timeout 300 python tools/time_seed.py 12000 16 500000 4000 2>&1 | grep -E "batches|seed \(greedy\)"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench.json").read())
    print("value %.3g ms_per_step %.4f kernel_ms %.4f frac %.3f e2e %.3g" % (j["value"], j["ms_per_step"], j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["e2e"]["value"]))
    print("e2e totals", j["e2e"]["repeat_totals_ms"])
    for k, v in (j.get("strong") or {}).items():
        print(k, {a: v.get(a) for a in ("ms", "balance", "ms_per_iter", "estep_ms", "allreduce_ms", "seed_ms", "error")})
    print(json.dumps(j.get("extra"), indent=1)[:1800])
except Exception as e:
    print("bench parse failed", e)
PY
nproc
timeout 600 python tools/time_loader.py 10000 --run 2>&1 | tail -34
