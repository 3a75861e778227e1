#!/usr/bin/env python
"""BASELINE.json configs[3] at its full size: 50 k cells x 64 samples x 1 M SNPs x the 21-point alpha grid, ONE problem cut
into 8 blocks of 6250 barcodes (same donors and SNPs, block b's cells drawn with cell_seed b).  Run with N = 1 (all 8 blocks on
one GPU, one after the other) and with torchrun at N = 8 (one block per GPU); time = CUDA events around pscl_demux_score,
summed over a rank's blocks, max over ranks.  time(1) / time(8) is the strong-scaling speed-up north_star asks for.
python tools/config4_full.py            |  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/config4_full.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
from popscle_b200 import Context, synth
from popscle_b200.capi import RawGeno, bind_to_device

BLOCKS, CELLS = 8, int(os.environ.get("CFG4_BLOCK_CELLS", 6250))
c4 = synth.CONFIGS[4]
nv, alphas = c4["nv"], list(c4["alphas"])
bind_to_device(local)
stream = torch.cuda.current_stream()
ctx = Context(local, stream=stream.cuda_stream)
mine = [b for b in range(BLOCKS) if b % world == rank]
tot_ms, tot_pairs, tot_reads, tot_dbl, gen_s, up_ms, checked = 0.0, 0, 0, 0, 0.0, 0.0, None
geno_set = False
for b in mine:
    t0 = time.perf_counter()
    s = synth.make_pileup(CELLS, nv, c4["V"], c4["kbar"], 20260105, cell_seed=b)
    s.plp.compact(); s.plp.compact3()
    gen_s += time.perf_counter() - t0
    t0 = time.perf_counter()
    if not geno_set:  # 64 MB of hard calls, mixed into the 1.5 GB table on the device; replicated on every GPU
        ctx.demux_set_geno(RawGeno(gt8=np.ascontiguousarray(s.geno.T.astype(np.uint8)), err=0.1), None, s.plp.n_snps)
        geno_set = True
    d = ctx.upload(s.plp, compact=3)
    ctx.sync()
    up_ms += 1e3 * (time.perf_counter() - t0)
    if b == mine[0]:
        ctx.demux_score(d, alphas, 0.5)  # warm-up pass (kernel load, scratch allocation)
        torch.cuda.synchronize()
    if world > 1 and b == mine[0]:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); ctx.demux_score(d, alphas, 0.5); e1.record(stream); e1.synchronize()
    tot_ms += e0.elapsed_time(e1)
    out = ctx.demux_fetch()
    tot_pairs += s.plp.n_pairs; tot_reads += s.plp.n_reads; tot_dbl += int((out["type"] == 1).sum())
    if rank == 0 and b == 0:  # parity on the block's first cells against the CPU oracle (singlet / doublet call and ids)
        import oracle_py as orc
        from tests.parity import check_demux_parity
        few = s.plp.slice_cells(0, 6)
        ref = orc.demux(few, synth.gt_to_gp(s.geno), None, alphas)
        check_demux_parity(out[:6], None, ref, None, alphas)
        truth_ok = float(np.mean((out["type"] != 0) | (out["sng_best"] == s.truth_d1)))
        checked = {"cells_vs_oracle": 6, "singlets_on_true_donor": truth_ok}
    d.free()
t = torch.tensor([tot_ms, gen_s, up_ms], dtype=torch.float64, device=dev)
c = torch.tensor([tot_ms, float(tot_pairs), float(tot_reads), float(tot_dbl)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(c)
if rank == 0:
    ms, pairs, reads = float(t[0]), float(c[1]), float(c[2])
    na = len(alphas)
    flops = pairs * (18.0 * nv * na + 7.0 * nv * nv * na + 36.0 * (reads / max(pairs, 1)) * na)
    print(json.dumps({"workload": "demuxlet configs[3] at full size: 50k cells x 64 samples x 1M SNPs x 21-point alpha grid, 8 barcode blocks",
                      "n_gpus": world, "cells": BLOCKS * CELLS, "pairs": int(pairs), "base_calls": int(reads), "ms": ms,
                      "balance": float(c[0]) / world / ms, "pairs_per_s": pairs / (ms * 1e-3), "base_calls_per_s": reads / (ms * 1e-3),
                      "fp64_tflops": flops / (ms * 1e-3) / 1e12, "doublets_found": int(c[3]), "generate_s_max": float(t[1]),
                      "upload_ms_max": float(t[2]), "parity": checked,
                      "timing": "CUDA events around pscl_demux_score per block, summed over a rank's blocks, max over ranks"}))
ctx.close()
if world > 1:
    dist.destroy_process_group()
