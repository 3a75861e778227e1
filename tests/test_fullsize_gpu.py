"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle would need minutes
for the whole thing, so it checks a random sample of cells): configs[1] demuxlet and configs[2] freemuxlet,
10k cells x 8 samples x 100k SNPs."""
import numpy as np
import pytest

import oracle_py as orc
from popscle_b200 import synth
from tests.parity import assert_close, check_demux_parity, check_fmx_parity

pytestmark = pytest.mark.gpu

DEFAULT = [0.0, 0.5]


@pytest.fixture(scope="module")
def cfg2():
    s = synth.make_config(2)
    return s, synth.gt_to_gp(s.geno)


def test_config2_kernels_agree_and_match_the_oracle_sample(ctx, cfg2):
    s, gp = cfg2
    plp = s.plp
    outs = {}
    d = ctx.upload(plp, compact=True)
    ctx.demux_set_geno(gp, None, plp.n_snps)
    ctx.demux_keep_grid(True)
    try:
        for name, k in (("lane", 1), ("dict", 6), ("poly", 4)):
            ctx.demux_select_kernel(k)
            ctx.demux_score(d, DEFAULT, 0.5)
            outs[name] = ctx.demux_fetch(want_grid=True)
    finally:
        ctx.demux_select_kernel(0)
        ctx.demux_keep_grid(False)
        d.free()
    rec, grid = outs["lane"]
    assert outs["dict"][0].tobytes() == rec.tobytes() and np.array_equal(outs["dict"][1], grid, equal_nan=True)  # same sums, coded genotypes
    live = ~np.isnan(grid)
    for name in ("poly",):  # an independent formulation of the same sums: ~1e-13 apart
        r2, g2 = outs[name]
        assert np.array_equal(np.isnan(g2), ~live)
        assert_close(g2[live], grid[live], f"{name} vs lane grid", rtol=1e-10)
        assert np.array_equal(r2["type"], rec["type"]) and np.array_equal(r2["sng_best"], rec["sng_best"])
    # the oracle on a random sample of cells
    rng = np.random.default_rng(2)
    for c0 in rng.integers(0, plp.n_cells - 25, 6):
        ref, rgrid = orc.demux(plp, gp, None, DEFAULT, 0.5, int(c0), int(c0) + 25, want_grid=True, n_threads=8)
        check_demux_parity(rec[c0:c0 + 25], grid[c0:c0 + 25], ref, rgrid, DEFAULT)
    # the synthetic truth: singlets go to their donor, doublets are found
    sng = rec["type"] == 0
    assert sng.mean() > 0.85 and (rec["sng_best"][sng] == s.truth_d1[sng]).mean() > 0.999
    dbl = s.truth_d1 != s.truth_d2
    assert (rec["type"][dbl] == 1).mean() > 0.95


def test_config2_sharding_and_sample_permutation(ctx, cfg2):
    s, gp = cfg2
    plp = s.plp
    full = ctx.demux_run(plp, gp, None, DEFAULT, compact=True)
    staged = ctx.demux_run(plp, gp, None, DEFAULT, compact=3)  # ABI-3 delta arrays: copied in slices under the scoring
    assert staged.tobytes() == full.tobytes()
    cuts = [0, 1234, 1235, 6000, plp.n_cells]
    parts = [ctx.demux_run(plp.slice_cells(a, b), gp, None, DEFAULT) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.concatenate(parts).tobytes() == full.tobytes()  # barcode shards are bit-identical
    perm = np.array([5, 2, 7, 0, 3, 6, 1, 4])
    p = ctx.demux_run(plp, gp[:, perm, :], None, DEFAULT, compact=True)
    assert np.array_equal(p["type"], full["type"])
    sng = full["type"] == 0
    assert np.array_equal(perm[p["sng_best"][sng]], full["sng_best"][sng])
    assert_close(p["sng_best_llk"], full["sng_best_llk"], "singlet LLK under sample permutation", rtol=1e-9)
    assert_close(p["dbl_best_llk"], full["dbl_best_llk"], "doublet LLK under sample permutation", rtol=1e-9)


def test_config3_freemuxlet_sharded_equals_unsharded_and_recovers_donors(ctx):
    """configs[2]: 3 forced EM iterations from the true donors; two SNP shards + summed partial LLKs (the
    all-reduce) reproduce the single-image run; clusters stay on their donors."""
    import torch
    from popscle_b200 import Context
    s = synth.make_config(3)
    plp = s.plp
    nS, iters = 8, 3
    o = ctx.fmx_opts(nS, early_stop=False, max_iter=iters)
    init = s.truth_d1.astype(np.int32)
    full, res, _, _ = ctx.fmx_run(plp, o, init, compact=True)
    sng = full["type"] == 0
    assert sng.mean() > 0.8 and (full["clust"][sng] == s.truth_d1[sng]).mean() > 0.999
    shards = [plp.slice_snps(0, 45_000), plp.slice_snps(45_000, plp.n_snps)]
    ctxs = [ctx, Context(0)]
    try:
        npairs = nS * (nS + 1) // 2
        dev = torch.device("cuda", 0)
        st = [torch.zeros(4 * plp.n_cells, dtype=torch.float64, device=dev) for _ in shards]
        llk = [torch.zeros(plp.n_cells * npairs, dtype=torch.float64, device=dev) for _ in shards]
        cl = [torch.zeros(plp.n_cells, dtype=torch.int32, device=dev) for _ in shards]
        init_d = torch.from_numpy(init).to(dev)
        keep = []
        for c, sh, a in zip(ctxs, shards, st):
            d = c.upload(sh, compact=True); keep.append(d)
            c.fmx_init(d, o)
            c.fmx_stage1(a.data_ptr())
            c.sync()
        tot = st[0] + st[1]
        for c, a, k in zip(ctxs, st, cl):
            a.copy_(tot); torch.cuda.synchronize()
            c.fmx_seed(a.data_ptr(), init_d.data_ptr(), k.data_ptr())
            c.fmx_mstep(k.data_ptr())
        for it in range(iters):
            for c, l in zip(ctxs, llk):
                c.fmx_estep(it, l.data_ptr()); c.sync()
            tot = llk[0] + llk[1]
            for c, l, k in zip(ctxs, llk, cl):
                l.copy_(tot); torch.cuda.synchronize()
                c.fmx_classify(l.data_ptr(), k.data_ptr())
                c.fmx_mstep(None)
        a, _, _ = ctxs[0].fmx_fetch()
        b, _, _ = ctxs[1].fmx_fetch()
        assert a.tobytes() == b.tobytes()
        check_fmx_parity(a, full, allow_tied_frac=0.05)
    finally:
        ctxs[1].close()


def test_config3_greedy_seeding_and_em_match_the_oracle_at_full_size(ctx):
    """configs[2] the way the default command runs it — stage 1, greedy seeding over all 10 k cells (speculative batches on the
    device, the serial chain in the oracle), two EM iterations — against the CPU oracle on all host threads: the same initial
    cluster of every cell, the same types and ids, LLKs within tolerance (full size, not 400 cells)."""
    import os
    s = synth.make_config(3)
    kw = dict(early_stop=False, max_iter=2)
    cells, res, _, _ = ctx.fmx_run(s.plp, ctx.fmx_opts(8, **kw), compact=4)
    r = orc.fmx_run(s.plp, orc.fmx_opts(8, **kw), n_threads=max(1, os.cpu_count() or 1))
    assert np.array_equal(cells["init_clust"], r["cells"]["init_clust"])
    check_fmx_parity(cells, r["cells"])
    assert res.n_iter == r["res"].n_iter and res.n_singlet == r["res"].n_singlet
    assert (np.bincount(cells["init_clust"], minlength=8) > 1000).all()
