#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_demux_gpu.py -x -q -m gpu -k "genotype_classes or genotype_dictionary or sharding" 2>&1 | tail -5
timeout 300 python tools/time_kernels.py --kernels dict,cls,dict,cls 2>&1 | tail -6 | tee gpurun_out/r5f_cls_timing.txt
