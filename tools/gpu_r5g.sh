#!/bin/bash
# Final visit of the round: every GPU test, smoke, the bench line, the ncu launch list of the same command and one
# --set full capture of the genotype-class k_demux_default (summarised here; the report stays on the box).
TAG=${1:-r5g}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
j = json.loads(open("gpurun_out/${TAG}_bench.json").read())
print("value %.3g ms_per_step %.4f kernel %s kernel_ms %.4f frac %.3f e2e %.3g" % (j["value"], j["ms_per_step"], j["roofline"]["kernel"], j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["e2e"]["value"]))
print("e2e totals", j["e2e"]["repeat_totals_ms"], sorted(j["e2e"]["ms_per_call"])[:5])
for k, v in (j.get("strong") or {}).items():
    print(k, {a: v.get(a) for a in ("ms", "balance", "ms_per_iter", "estep_ms", "allreduce_ms", "seed_ms", "error")})
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_demux_default -s 3 -c 1 -f -o gpurun_out/${TAG}_cls \
  python tools/time_kernels.py --kernels cls --steps 2 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"
python tools/ncu_summary.py gpurun_out/${TAG}_cls.ncu-rep 0 > gpurun_out/${TAG}_k_demux_default_classes8_ncu.txt 2>&1
rm -f gpurun_out/${TAG}_cls.ncu-rep
head -24 gpurun_out/${TAG}_k_demux_default_classes8_ncu.txt
