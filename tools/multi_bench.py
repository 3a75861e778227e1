#!/usr/bin/env python
"""Strong scaling of the PRODUCT multi-GPU path (pscl_multi_demux_run / pscl_multi_fmx_run: one process, one host thread per
GPU, peer-memory all-reduce) on the same two problems bench.py's `strong` object uses.
    python tools/multi_bench.py <n_gpus> [demux_cells] [fmx_cells]     -> one JSON line"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
from popscle_b200 import Multi, RawGeno, synth
from popscle_b200 import bench_strong as bs

n = int(sys.argv[1])
dc = int(sys.argv[2]) if len(sys.argv) > 2 else bs.DEMUX64_CELLS
fc = int(sys.argv[3]) if len(sys.argv) > 3 else bs.FREEMUX16_CELLS
out = {"n_gpus": n, "api": "pscl_multi_* (single process)"}
with Multi(n_gpu=n) as m:
    c4 = synth.CONFIGS[4]
    s = synth.make_pileup(dc, c4["nv"], c4["V"], c4["kbar"], 20260105)
    raw = RawGeno(gt8=np.ascontiguousarray(s.geno.T.astype(np.uint8)), err=0.1)
    s.plp.compact(); s.plp.compact3()
    runs = []
    for rep in range(3):
        t0 = time.perf_counter()
        rec = m.demux_run(s.plp, raw, None, list(c4["alphas"]), compact=3)
        runs.append(1e3 * (time.perf_counter() - t0))
        t = m.timing()
    out["demux64"] = {"cells": dc, "pairs": s.plp.n_pairs, "wall_ms": runs, "kernel_ms_max": max(t["kernel_ms"]), "kernel_ms": t["kernel_ms"],
                      "upload_ms": t["upload_ms"], "units": t["units"], "doublets": int((rec["type"] == 1).sum())}
    c5 = synth.CONFIGS[5]
    s = synth.make_pileup(fc, c5["nv"], c5["V"], c5["kbar"], 20260106)
    s.plp.compact(); s.plp.compact3()
    o = m.fmx_opts(c5["nv"], early_stop=False, max_iter=bs.FREEMUX16_ITERS)
    for tag, timed in (("freemux16", False), ("freemux16_allreduce_isolated", True)):
        if timed:
            os.environ["PSCL_MULTI_TIME_ALLREDUCE"] = "1"  # two extra stream synchronisations per iteration around the collective
        t0 = time.perf_counter()
        cells, res, _, _ = m.fmx_run(s.plp, o, compact=3)
        wall = 1e3 * (time.perf_counter() - t0)
        t = m.timing()
        out[tag] = {"cells": fc, "pairs": s.plp.n_pairs, "wall_ms": wall, "iters": t["iters"], "seed_ms": t["seed_ms"],
                    "em_ms_per_iter_max": max(t["compute_ms"]) / max(t["iters"], 1), "allreduce_ms": t["allreduce_ms"],
                    "allreduce_bytes": t["allreduce_bytes"], "upload_ms": t["upload_ms"], "setup_ms": t["setup_ms"], "units": t["units"],
                    "singlets": int((cells["type"] == 0).sum())}
print(json.dumps(out), flush=True)
