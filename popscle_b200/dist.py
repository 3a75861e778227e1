"""Multi-GPU partitioning: host-side logic only, no device code.

demuxlet shards BARCODES: every cell's likelihood grid depends only on its own reads and the
read-only genotype table (cmd_cram_demuxlet.cpp:636-1013 carries no cross-cell state), so ranks own
contiguous cell ranges balanced by pair count and never communicate; rank 0 gathers the 160-byte
records.  freemuxlet shards SNPs: llk[c][pair] is a sum over SNPs (cmd_cram_freemux2.cpp:454-455),
so ranks own SNP ranges balanced by pair count, all-reduce the C x npairs partial sums once per EM
iteration, classify redundantly, and run the SNP-local M-step without communication.
"""
from __future__ import annotations

import numpy as np

from .capi import Pileup


def balanced_cell_ranges(cell_ptr: np.ndarray, n: int):
    """n contiguous cell ranges with near-equal pair counts (not near-equal cell counts)."""
    C = len(cell_ptr) - 1
    P = int(cell_ptr[-1])
    cuts = [0]
    for r in range(1, n):
        c = int(np.searchsorted(cell_ptr, P * r / n, side="left"))
        cuts.append(min(max(c, cuts[-1]), C))
    cuts.append(C)
    return [(cuts[i], cuts[i + 1]) for i in range(n)]


def balanced_snp_ranges(pair_snp: np.ndarray, n_snps: int, n: int):
    """n contiguous SNP ranges with near-equal pair counts."""
    cnt = np.bincount(pair_snp, minlength=n_snps)
    cum = np.concatenate([[0], np.cumsum(cnt)])
    P = int(cum[-1])
    cuts = [0]
    for r in range(1, n):
        v = int(np.searchsorted(cum, P * r / n, side="left"))
        cuts.append(min(max(v, cuts[-1]), n_snps))
    cuts.append(n_snps)
    return [(cuts[i], cuts[i + 1]) for i in range(n)]


def demux_sharded(engine, plp: Pileup, gp, has_gp, alphas, doublet_prior, rank: int, world: int, gather=None):
    """This rank's barcode shard through `engine.demux_run`; `gather(records)` (e.g. an
    all_gather_object / gather to rank 0) concatenates the shards in rank order."""
    c0, c1 = balanced_cell_ranges(plp.cell_ptr, world)[rank]
    mine = engine.demux_run(plp.slice_cells(c0, c1), gp, has_gp, alphas, doublet_prior)
    if gather is None:
        return mine
    parts = gather(mine)
    return np.concatenate(parts) if parts is not None else None


def fmx_em_sharded(step, plp_shard: Pileup, opts, init_clust, allreduce, n_cells: int):
    """SNP-sharded EM driver.  `step` exposes the step-level C ABI on this rank's shard
    (init/stage1/seed/mstep/estep/classify/fetch working on array-like buffers), `allreduce(buf)`
    sums a buffer over ranks in place.  Greedy seeding needs all SNPs on one rank, so a sharded
    run starts from init_clust (the --init-cluster path) — SURVEY §8e."""
    if init_clust is None:
        raise ValueError("SNP-sharded freemuxlet needs initial clusters (seed on one rank first, or --init-cluster)")
    nS = opts.n_clusters
    st = step.new_f64(4 * n_cells)
    llk = step.new_f64(n_cells * nS * (nS + 1) // 2)
    cl = step.new_i32(n_cells)
    step.init(plp_shard, opts)
    step.stage1(st)
    allreduce(st)
    step.seed(st, init_clust, cl)
    step.mstep(cl)
    res = None
    for it in range(opts.max_iter):
        step.estep(it, llk)
        allreduce(llk)
        res = step.classify(llk, cl)
        step.mstep(None)
        if not opts.mode_old and opts.early_stop and res.n_changed == 0:
            break
    return step.fetch(), res
