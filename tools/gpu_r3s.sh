#!/bin/bash
TAG=${1:-r3s}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest.log
python tools/e2e_jitter.py 2>&1 | tail -4 | cut -c1-330
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
j = json.loads(open("gpurun_out/${TAG}_bench.json").read())
print("value %.3g ms_per_step %.4f kernel_ms %.4f frac %.3f e2e %.3g" % (j["value"], j["ms_per_step"], j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["e2e"]["value"]))
print("e2e totals", j["e2e"]["repeat_totals_ms"], j["e2e"]["ms_per_call"])
for k, v in (j.get("strong") or {}).items():
    print(k, {a: v.get(a) for a in ("ms", "balance", "ms_per_iter", "estep_ms", "allreduce_ms", "seed_ms", "seed_call_ms", "error")})
print({k: v for k, v in j["extra"]["freemux_cfg3"].items() if k.endswith("ms") or "seed" in k})
PY
