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


def fmx_em_sharded(step, plp_shard: Pileup, opts, init_clust, allreduce, n_cells: int, seed_full=None, bcast=None):
    """SNP-sharded EM driver (one rank per GPU).  `step` exposes the step-level C ABI on this rank's shard
    (init/stage1/seed/mstep/estep/classify/fetch working on array-like buffers), `allreduce(buf)` sums a buffer over
    ranks in place.

    Greedy seeding (cmd_cram_freemux2.cpp:223-260) is a sequential chain over the cells that needs every SNP of a cell,
    so it cannot be sharded: without `init_clust`, ONE rank runs stage 1 + seeding over the whole pileup
    (`seed_full()` -> (stage-1 planes [4*C], clusters [C]); None on the other ranks) and `bcast(arr_or_None)` hands both
    to everybody (400 KB at 100k cells) — SURVEY §8e.  The whole-pileup stage-1 sums are then used as they are, so the
    `.lmix` columns carry the same bits as a single-GPU run."""
    nS = opts.n_clusters
    st = step.new_f64(4 * n_cells)
    llk = step.new_f64(n_cells * nS * (nS + 1) // 2)
    cl = step.new_i32(n_cells)
    seeded_st = None
    if init_clust is None:
        if seed_full is None or bcast is None:
            raise ValueError("SNP-sharded freemuxlet without --init-cluster needs seed_full (stage 1 + greedy seeding over the "
                             "whole pileup on one rank) and bcast")
        got = seed_full()
        seeded_st = bcast(None if got is None else np.ascontiguousarray(got[0], dtype=np.float64))
        init_clust = bcast(None if got is None else np.ascontiguousarray(got[1], dtype=np.int32))
    step.init(plp_shard, opts)
    step.stage1(st)
    if seeded_st is not None:
        step.set_f64(st, seeded_st)
    else:
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


class CudaStep:
    """The step-level C ABI of one Context behind the `step` interface of fmx_em_sharded; buffers are torch tensors on
    the context's device (so that torch.distributed's NCCL all-reduce works on them in place)."""

    def __init__(self, ctx, device):
        import torch
        self.ctx, self.torch, self.dev = ctx, torch, device
        self._keep = []

    def new_f64(self, n):
        return self.torch.zeros(n, dtype=self.torch.float64, device=self.dev)

    def new_i32(self, n):
        return self.torch.zeros(n, dtype=self.torch.int32, device=self.dev)

    def set_f64(self, buf, values):
        buf.copy_(self.torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64)))
        self.torch.cuda.synchronize(self.dev)

    def init(self, plp, opts, compact=3):
        self.dplp = self.ctx.upload(plp, compact=compact)
        self.ctx.fmx_init(self.dplp, opts)

    def stage1(self, st):
        self.ctx.fmx_stage1(st.data_ptr())

    def seed(self, st, init_clust, cl):
        ic = None
        if init_clust is not None:
            ic = self.torch.from_numpy(np.ascontiguousarray(init_clust, dtype=np.int32)).to(self.dev)
            self._keep.append(ic)
        self.ctx.fmx_seed(st.data_ptr(), ic.data_ptr() if ic is not None else None, cl.data_ptr())

    def mstep(self, cl):
        self.ctx.fmx_mstep(cl.data_ptr() if cl is not None else None)

    def estep(self, it, llk):
        self.ctx.fmx_estep(it, llk.data_ptr())

    def classify(self, llk, cl):
        return self.ctx.fmx_classify(llk.data_ptr(), cl.data_ptr())

    def fetch(self):
        return self.ctx.fmx_fetch()[0]

    def seed_whole(self, plp, opts, compact=3):
        """stage 1 + greedy seeding over a whole pileup on this context: (stage-1 planes, clusters) as numpy arrays"""
        d = self.ctx.upload(plp, compact=compact)
        try:
            self.ctx.fmx_init(d, opts)
            st, cl = self.new_f64(4 * plp.n_cells), self.new_i32(plp.n_cells)
            self.ctx.fmx_stage1(st.data_ptr())
            self.ctx.sync()
            import time
            t0 = time.perf_counter()
            self.ctx.fmx_seed(st.data_ptr(), None, cl.data_ptr())
            self.ctx.sync()
            self.last_seed_ms = 1e3 * (time.perf_counter() - t0)  # the seeding call alone (the rest is upload, views, stage 1)
            return st.cpu().numpy(), cl.cpu().numpy()
        finally:
            d.free()
