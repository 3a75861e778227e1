#!/bin/bash
echo "== team"; timeout 300 python tools/time_seed.py 12000 16 500000 4000 2>&1 | grep -E "estep"
timeout 600 python -m pytest tests/test_freemux_gpu.py -x -q -m gpu -k "team or many_clusters or 2000_cells" 2>&1 | tail -2
