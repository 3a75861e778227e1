"""bench.py's `strong` object: ONE fixed problem per workload, sharded over the N ranks of the run (one rank per GPU).

  demux64    BASELINE.json configs[3]'s shape (64 samples x 1M SNPs, 21-point alpha grid) on a fixed global cell count:
             barcodes sharded into contiguous ranges balanced by pair count (dist.balanced_cell_ranges), genotype table
             replicated, no collective.  ms = one pass of pscl_demux_score over the rank's shard, CUDA events, max over ranks.
  freemux16  configs[4]'s shape (--nsample 16 x 500k SNPs) on a fixed global cell count: stage 1 + greedy seeding over the
             whole pileup on rank 0, clusters broadcast, then SNP-sharded EM with one NCCL all-reduce of the C x npairs
             partial LLKs per iteration (dist.fmx_em_sharded's sequence, timed phase by phase).

Every rank generates the same synthetic problem (same seed) and keeps its shard; the numbers at N = 1 are the whole problem
on one GPU, so time(N=1) / time(N) is the strong-scaling speed-up.
"""
from __future__ import annotations

import time

import numpy as np

DEMUX64_CELLS = 3072
FREEMUX16_CELLS = 12000
FREEMUX16_ITERS = 5


def _ncu_traffic(key):
    import json
    import os
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")) as f:
            return json.load(f).get(key)
    except (OSError, ValueError):
        return None


def _max_over_ranks(x, world, dev):
    import torch
    import torch.distributed as dist
    t = torch.tensor(x, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def _sum_over_ranks(x, world, dev):
    import torch
    import torch.distributed as dist
    t = torch.tensor(x, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    return [float(v) for v in t.tolist()]


def demux64(ctx, rank, world, dev, cells=DEMUX64_CELLS, steps=2, warmup=1):
    import torch
    import torch.distributed as td
    from . import synth
    from .capi import RawGeno
    from .dist import balanced_cell_ranges
    c4 = synth.CONFIGS[4]
    t0 = time.perf_counter()
    s = synth.make_pileup(cells, c4["nv"], c4["V"], c4["kbar"], 20260105)  # the same problem on every rank
    gen_s = time.perf_counter() - t0
    plp, nv, alphas = s.plp, c4["nv"], list(c4["alphas"])
    c0, c1 = balanced_cell_ranges(plp.cell_ptr, world)[rank]
    mine = plp.slice_cells(c0, c1)
    mine.compact(); mine.compact3()
    stream = torch.cuda.current_stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.demux_set_geno(RawGeno(gt8=np.ascontiguousarray(s.geno.T.astype(np.uint8)), err=0.1), None, plp.n_snps)  # 64 MB, mixed on the device
    d = ctx.upload(mine, compact=3)
    ctx.sync()
    up_ms = 1e3 * (time.perf_counter() - t0)
    try:
        for _ in range(warmup):
            ctx.demux_score(d, alphas, 0.5)
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            a.record(stream)
            ctx.demux_score(d, alphas, 0.5)
            b.record(stream)
            b.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / steps
        out = ctx.demux_fetch()
        kernel = ctx.demux_last_kernel()
    finally:
        d.free()
    ms_max, up_max = _max_over_ranks([ms, up_ms], world, dev)
    ms_sum, pairs_sum, dbl = _sum_over_ranks([ms, mine.n_pairs, float((out["type"] == 1).sum())], world, dev)
    na = len(alphas)
    flops = plp.n_pairs * (18.0 * nv * na + 7.0 * nv * nv * na + 36.0 * (plp.n_reads / max(plp.n_pairs, 1)) * na)
    return {"workload": "demuxlet configs[3] shape: 64 samples x 1M SNPs x 21-point alpha grid, ONE problem sharded by barcodes",
            "cells": cells, "pairs": plp.n_pairs, "base_calls": plp.n_reads, "n_gpus": world,
            "ms": ms_max, "ms_mean_over_ranks": ms_sum / world, "balance": (ms_sum / world) / ms_max if ms_max > 0 else None,
            "pairs_per_s": plp.n_pairs / (ms_max * 1e-3), "base_calls_per_s": plp.n_reads / (ms_max * 1e-3),
            "fp64_tflops": flops / (ms_max * 1e-3) / 1e12, "upload_ms": up_max, "doublets_found": int(dbl),
            "kernel": {4: "k_demux_poly", 2: "k_demux_general"}.get(kernel, str(kernel)), "collective": None,
            "limited_by": "k_demux_poly (FP64 issue); the only serial part is the slowest rank's shard (balance = mean/max of the ranks' times)",
            "timing": f"CUDA events around pscl_demux_score on each rank's shard, mean of {steps} passes, max over ranks",
            "generate_s": gen_s}


def freemux16(ctx, rank, world, dev, cells=FREEMUX16_CELLS, iters=FREEMUX16_ITERS):
    import torch
    import torch.distributed as td
    from . import synth
    from .dist import CudaStep, balanced_snp_ranges
    c5 = synth.CONFIGS[5]
    nS = c5["nv"]
    npairs = nS * (nS + 1) // 2
    t0 = time.perf_counter()
    s = synth.make_pileup(cells, nS, c5["V"], c5["kbar"], 20260106)
    gen_s = time.perf_counter() - t0
    plp = s.plp
    v0, v1 = balanced_snp_ranges(plp.pair_snp, plp.n_snps, world)[rank]
    shard = plp.slice_snps(v0, v1) if world > 1 else plp
    for x in (plp, shard):  # the compact host arrays (ABI 3) are built once, outside every timed region
        x.compact(); x.compact3()
    stream = torch.cuda.current_stream()
    step = CudaStep(ctx, dev)
    o = ctx.fmx_opts(nS, early_stop=False, max_iter=iters)
    C = plp.n_cells
    # ---- stage 1 + greedy seeding over the whole pileup on rank 0, broadcast (cmd_cram_freemux2.cpp:117-261) ----
    st_all = torch.zeros(4 * C, dtype=torch.float64, device=dev)
    cl_all = torch.zeros(C, dtype=torch.int32, device=dev)
    seed_ms = 0.0
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if rank == 0:
        st_np, cl_np = step.seed_whole(plp, o)
        st_all.copy_(torch.from_numpy(st_np)); cl_all.copy_(torch.from_numpy(cl_np))
        torch.cuda.synchronize()
        seed_ms = 1e3 * (time.perf_counter() - t0)
    if world > 1:
        td.broadcast(st_all, 0); td.broadcast(cl_all, 0)
    torch.cuda.synchronize()
    seed_wall_ms = 1e3 * (time.perf_counter() - t0)
    # ---- SNP-sharded EM ----
    t0 = time.perf_counter()
    step.init(shard, o)
    st = step.new_f64(4 * C); llk = step.new_f64(C * npairs); cl = step.new_i32(C)
    step.stage1(st)
    ctx.fmx_seed(st_all.data_ptr(), cl_all.data_ptr(), cl.data_ptr())
    step.mstep(cl)
    torch.cuda.synchronize()
    setup_ms = 1e3 * (time.perf_counter() - t0)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = []
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    res = None
    for it in range(iters):
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record(stream)
        step.estep(it, llk)
        e1.record(stream)
        if world > 1:
            td.all_reduce(llk)
        e2.record(stream)
        res = step.classify(llk, cl)  # synchronises (nchanged read-back)
        step.mstep(None)
        e3.record(stream)
        e3.synchronize()
        marks.append((e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e0.elapsed_time(e3)))
    cells_out = step.fetch()
    step.dplp.free()
    m = np.array(marks[1:] if len(marks) > 1 else marks)  # the first iteration pays NCCL's lazy set-up
    it_ms, es_ms, ar_ms, cm_ms = _max_over_ranks([m[:, 3].mean(), m[:, 0].mean(), m[:, 1].mean(), m[:, 2].mean()], world, dev)
    (es_sum,) = _sum_over_ranks([m[:, 0].mean()], world, dev)
    setup_max, = _max_over_ranks([setup_ms], world, dev)
    sng = cells_out["type"] == 0
    purity = 0.0
    if sng.any():  # clusters are arbitrary labels: fraction of singlets that share their cluster's majority donor
        tab = np.zeros((nS, nS), dtype=np.int64)
        np.add.at(tab, (cells_out["clust"][sng], s.truth_d1[sng]), 1)
        purity = float(tab.max(axis=1).sum() / max(tab.sum(), 1))
    return {"workload": "freemuxlet configs[4] shape: --nsample 16 x 500k SNPs, ONE problem sharded by SNPs, greedy seeding on rank 0",
            "cells": cells, "clusters": nS, "pairs": plp.n_pairs, "base_calls": plp.n_reads, "n_gpus": world, "iters": iters,
            "ms_per_iter": it_ms, "estep_ms": es_ms, "allreduce_ms": ar_ms, "classify_mstep_ms": cm_ms,
            "estep_balance": (es_sum / world) / es_ms if es_ms > 0 else None,
            "allreduce_bytes": int(C * npairs * 8), "collective": "NCCL all-reduce (torch.distributed) of C x npairs FP64 per EM iteration" if world > 1 else None,
            "seed_ms": seed_ms, "seed_call_ms": getattr(step, "last_seed_ms", None), "seed_wall_ms_incl_broadcast": seed_wall_ms, "setup_ms": setup_max,
            "base_calls_per_s": plp.n_reads / (it_ms * 1e-3), "singlets": int(sng.sum()), "singlet_cluster_purity": purity,
            "n_changed_last": int(res.n_changed) if res is not None else None,
            "limited_by": "k_fmx_estep (the nS = 16 row tiles) inside an iteration; for the whole run rank 0's set-up over the whole pileup (upload, SNP-major view, stage 1; seed_ms) of which the speculative-batch seeding itself is seed_call_ms",
            "timing": "CUDA events per phase on each rank, mean over iterations 2.., max over ranks", "generate_s": gen_s}


def freemux_cfg3(ctx, plp, truth, dev, iters=10):
    """configs[2] on one GPU, the way the default command runs it: stage 1, GREEDY SEEDING, 10 forced EM iterations —
    device-resident phase times and the end-to-end pscl_fmx_run."""
    import torch
    from .dist import CudaStep
    nS = 8
    npairs = nS * (nS + 1) // 2
    plp.compact(); plp.compact3()
    stream = torch.cuda.current_stream()
    step = CudaStep(ctx, dev)
    o = ctx.fmx_opts(nS, early_stop=False, max_iter=iters)
    C = plp.n_cells

    def timed(fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0)
    # one untimed pass first (kernel loading, the first allocation of every buffer), as for every other number of the line
    step.init(plp, o)
    st = step.new_f64(4 * C); llk = step.new_f64(C * npairs); cl = step.new_i32(C)
    step.stage1(st); step.seed(st, None, cl); step.mstep(cl); step.estep(0, llk); step.classify(llk, cl); step.mstep(None)
    torch.cuda.synchronize()
    step.dplp.free()
    up_ms = timed(lambda: step.init(plp, o))
    st = step.new_f64(4 * C); llk = step.new_f64(C * npairs); cl = step.new_i32(C)
    s1_ms = timed(lambda: step.stage1(st))
    seed_ms = timed(lambda: step.seed(st, None, cl))
    m0_ms = timed(lambda: step.mstep(cl))
    l0 = ctx.launch_count
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es = []
    a.record(stream)
    for it in range(iters):
        step.estep(it, llk)
        step.classify(llk, cl)
        es.append(ctx.fmx_last_kernel_ms())
        step.mstep(None)
    b.record(stream); b.synchronize()
    em_ms = a.elapsed_time(b)
    launches = ctx.launch_count - l0
    cells = step.fetch()
    step.dplp.free()
    sng = cells["type"] == 0
    tab = np.zeros((nS, nS), dtype=np.int64)
    np.add.at(tab, (cells["clust"][sng], truth[sng]), 1)
    # end to end through the one-call C ABI (compact pinned host arrays in, records out), greedy seeding included
    ctx.fmx_run(plp, ctx.fmx_opts(nS, early_stop=False, max_iter=iters), None, compact=3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.fmx_run(plp, ctx.fmx_opts(nS, early_stop=False, max_iter=iters), None, compact=3)
    e2e_s = time.perf_counter() - t0
    abytes = (72 + 4 + 24 * nS) * plp.n_pairs + 8 * C * npairs
    return {"workload": "freemuxlet configs[2]: 10k cells x nsample 8 x 100k SNPs, greedy seeding + 10 forced EM iterations, 1 GPU",
            "upload_and_snp_major_view_ms": up_ms, "stage1_ms": s1_ms, "seed_ms": seed_ms, "first_mstep_ms": m0_ms,
            "em_ms_total": em_ms, "em_ms_per_iter": em_ms / iters, "estep_ms": float(np.mean(es)), "iters": iters,
            "seed_over_em": seed_ms / em_ms if em_ms > 0 else None,
            "base_calls_per_s_em": plp.n_reads * iters / (em_ms * 1e-3),
            "estep_algorithmic_gbs": None if not es else abytes / (float(np.mean(es)) * 1e-3) / 1e9,
            "estep_dram_traffic_bytes": _ncu_traffic("k_fmx_estep"),  # dram read + write of one E-step from the committed ncu capture
            "e2e_ms": 1e3 * e2e_s, "e2e_base_calls_per_s": plp.n_reads * iters / e2e_s, "gpu_launches": int(launches),
            "singlets": int(sng.sum()), "singlet_cluster_purity": float(tab.max(axis=1).sum() / max(tab.sum(), 1))}
