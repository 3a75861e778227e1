#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_demux_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== default (genotype classes)"; timeout 120 python tools/e2e_jitter.py 2>&1 | grep median | cut -c1-200
echo "== PSCL_NO_CLS=1"; PSCL_NO_CLS=1 timeout 120 python tools/e2e_jitter.py 2>&1 | grep median | cut -c1-200
