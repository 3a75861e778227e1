#!/bin/bash
TAG=${1:-r3n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_freemux_gpu.py -x -q -m gpu -k "team or many_clusters or 2000_cells or sharded" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest.log
echo "== team"; timeout 300 python tools/time_seed.py 12000 16 500000 4000 2>&1 | grep -E "estep|mstep 1|classify 1"
echo "== tiles"; PSCL_ESTEP_TILES=1 timeout 300 python tools/time_seed.py 12000 16 500000 4000 2>&1 | grep -E "estep"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:"k_fmx_estep16_team" -s 1 -c 1 -o gpurun_out/${TAG}_team python tools/time_seed.py 12000 16 500000 4000 > gpurun_out/${TAG}_team.log 2>&1; echo "ncu exit $?"
python tools/ncu_summary.py gpurun_out/${TAG}_team.ncu-rep 0 > gpurun_out/${TAG}_k_fmx_estep16_team_ncu.txt 2>&1; rm -f gpurun_out/${TAG}_team.ncu-rep
head -24 gpurun_out/${TAG}_k_fmx_estep16_team_ncu.txt
