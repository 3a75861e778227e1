"""bench.py --workload freemux: EM iterations of `popscle freemuxlet` on BASELINE.json configs[2]
(10k cells x --nsample 8 x 100k SNPs, forced iterations, no early stop).

One step = one EM iteration over the resident pileup: posterior table + E-step kernel + per-cell
LLK reduce, [NCCL all-reduce of the C x npairs partial LLKs when SNP-sharded], classify, M-step.
With N > 1 every rank owns its own SNP range of the same cells (weak scaling over SNPs).
"""
from __future__ import annotations

import json
import os
import time

import numpy as np


def main(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from popscle_b200 import Context, _build, synth
    import bench as B

    _build.build_cuda()
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = dict(synth.CONFIGS[3])
    if args.cells:
        cfg["C"] = args.cells
    nS = cfg["nv"]
    npairs = nS * (nS + 1) // 2
    # same cells and donors on every rank (same seed for the cell-level draws), rank-specific SNPs
    s = synth.make_pileup(cfg["C"], nS, cfg["V"], cfg["kbar"], 20260104, snp_seed=1000 * rank)
    plp = s.plp
    stream = torch.cuda.current_stream()
    ctx = Context(local_rank, stream=stream.cuda_stream)
    dev = torch.device("cuda", local_rank)
    C = plp.n_cells
    st = torch.zeros(4 * C, dtype=torch.float64, device=dev)
    llk = torch.zeros(C * npairs, dtype=torch.float64, device=dev)
    cl = torch.zeros(C, dtype=torch.int32, device=dev)
    init = torch.from_numpy(s.truth_d1.astype(np.int32)).to(dev)
    iters = args.steps + max(args.warmup, 3)
    o = ctx.fmx_opts(nS, early_stop=False, max_iter=iters)
    dplp = ctx.upload(plp)
    ctx.fmx_init(dplp, o)
    ctx.fmx_stage1(st.data_ptr())
    if world > 1:
        dist.all_reduce(st)
    # seeding is a sequential chain over cells that needs all SNPs: the benchmark starts
    # the EM from the donors' true labels, i.e. the --init-cluster path
    ctx.fmx_seed(st.data_ptr(), init.data_ptr(), cl.data_ptr())
    ctx.fmx_mstep(cl.data_ptr())

    def em_iter(it):
        ctx.fmx_estep(it, llk.data_ptr())
        if world > 1:
            dist.all_reduce(llk)
        ctx.fmx_classify(llk.data_ptr(), cl.data_ptr())
        ctx.fmx_mstep(None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    W = max(args.warmup, 3)
    for it in range(W):
        em_iter(it)
    barrier()
    sampler = B.make_sampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    k_ms = []
    barrier()
    for i, (a, b) in enumerate(ev):
        flush.zero_()
        a.record(stream)
        ctx.fmx_estep(W + i, llk.data_ptr())
        e_ms = None
        if world > 1:
            dist.all_reduce(llk)
        ctx.fmx_classify(llk.data_ptr(), cl.data_ptr())  # synchronises (nchanged read-back)
        e_ms = ctx.fmx_last_kernel_ms()
        ctx.fmx_mstep(None)
        b.record(stream)
        b.synchronize()
        k_ms.append(e_ms)
    barrier()
    launches = ctx.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    n = torch.tensor([plp.n_reads, plp.n_pairs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n)
    t_ms = float(t.item())
    value = float(n[0].item()) * args.steps / (t_ms * 1e-3)

    # end to end: host pileup (pinned, ABI-3 compact arrays) in, per-cell records out; whole run of 10 forced iterations
    from popscle_b200 import Pileup
    keep, arrs = [], {}
    p32, aq = plp.compact()
    first, d16, n8 = plp.compact3()
    for name, src in (("cell_ptr", plp.cell_ptr), ("cell_first_snp", first), ("pair_snp_delta16", d16), ("pair_nreads8", n8), ("read_aq", aq), ("snp_af", plp.snp_af)):
        t_ = torch.from_numpy(src).pin_memory(); keep.append(t_); arrs[name] = t_.numpy()
    hplp = Pileup(plp.n_cells, plp.n_snps, arrs["cell_ptr"], plp.pair_snp, plp.pair_read_ptr, plp.read_allele, plp.read_qual, arrs["snp_af"])
    hplp._compact = (p32, arrs["read_aq"])
    hplp._compact3 = (arrs["cell_first_snp"], arrs["pair_snp_delta16"], arrs["pair_nreads8"])
    h2d = sum(v.nbytes for v in arrs.values())
    e2e_iters = 10
    init = s.truth_d1.astype(np.int32)
    ctx.fmx_run(hplp, ctx.fmx_opts(nS, early_stop=False, max_iter=e2e_iters), init, compact=3)  # warm the memory pool
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cells, res, _, _ = ctx.fmx_run(hplp, ctx.fmx_opts(nS, early_stop=False, max_iter=e2e_iters), init, compact=3)
    e2e_dt = time.perf_counter() - t0
    if rank == 0:
        peak, peak_src = B.measured_peak_gbs()
        est = float(np.mean(k_ms))
        abytes = B.algorithmic_bytes_fmx_iter(plp, nS)
        ach = abytes / (est * 1e-3) / 1e9
        line = {"metric": "pileup base-calls scored/sec (freemuxlet EM iteration)", "value": value, "unit": "base-calls/s",
                "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": t_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "freemuxlet configs[2]: 10k cells x nsample 8 x 100k SNPs, forced EM iterations",
                           "cells": C, "clusters": nS, "snps_per_gpu": plp.n_snps, "pairs_per_gpu": plp.n_pairs,
                           "base_calls_per_gpu": plp.n_reads, "sharding": f"SNPs x{world}" + (", NCCL all-reduce of C x npairs per iteration" if world > 1 else ""),
                           "l2": "flushed between timed steps (256 MiB memset, untimed)", "init": "--init-cluster (true donors)"},
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": B.ncu_traffic("k_fmx_estep"), "kernel": f"k_fmx_posterior + k_fmx_estep<{nS}> + k_fmx_llk_reduce",
                             "kernel_ms": est, "algorithmic_bytes": abytes, "peak_source": peak_src},
                "e2e": {"value": plp.n_reads * e2e_iters / e2e_dt, "unit": "base-calls/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(160 * C), "api": f"pscl_fmx_run, {e2e_iters} forced iterations incl. upload, SNP-major view, stage 1"},
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
