"""GPU parity of the freemuxlet path (through the C ABI) against the CPU oracle."""
import os

import numpy as np
import pytest

import oracle_py as orc
from popscle_b200 import synth
from tests.parity import assert_close, check_fmx_parity

pytestmark = pytest.mark.gpu


def _both(ctx, plp, nS, init=None, want_clusters=True, **kw):
    cells, res, gl, cnt = ctx.fmx_run(plp, ctx.fmx_opts(nS, **kw), init, want_clusters=want_clusters)
    r = orc.fmx_run(plp, orc.fmx_opts(nS, **kw), init, want_clusters=want_clusters, n_threads=8)
    return cells, res, gl, cnt, r


def _check(cells, res, gl, cnt, r, tied=0.02):
    n = check_fmx_parity(cells, r["cells"], allow_tied_frac=tied)
    rr = r["res"]
    assert (res.n_iter, res.n_changed, res.n_singlet, res.n_doublet, res.n_ambiguous) == \
           (rr.n_iter, rr.n_changed, rr.n_singlet, rr.n_doublet, rr.n_ambiguous)
    if gl is not None:
        assert_close(gl, r["clust_gl"], "cluster GLs", rtol=1e-6)
        assert np.array_equal(cnt, r["clust_cnt"])
    return n


@pytest.mark.parametrize("nS", [2, 3, 4, 5, 8])
def test_greedy_seeding_and_em(ctx, nS):
    """whole run: stage 1, greedy seeding, EM with early stop — cluster ids must match exactly."""
    s = synth.make_pileup(C=400, nv=nS, V=2500, kbar=300, seed=500 + nS)
    out = _both(ctx, s.plp, nS)
    n = _check(*out)
    assert n > 0.95 * len(out[0])
    # the clusters recover the donors up to a relabelling
    cells = out[0]
    sng = cells["type"] == 0
    tab = np.zeros((nS, nS), dtype=int)
    np.add.at(tab, (cells["clust"][sng], s.truth_d1[sng]), 1)
    assert tab.max(axis=1).sum() > 0.9 * sng.sum()


@pytest.mark.parametrize("nS", [9, 12, 16, 20, 24, 27, 32])
def test_many_clusters(ctx, nS):
    """runtime-nS tile kernels (more than 8 clusters; one row per launch past 24), up to the cap of 32."""
    s = synth.make_pileup(C=160, nv=nS, V=1500, kbar=250, seed=600 + nS)
    _check(*_both(ctx, s.plp, nS, max_iter=3), tied=0.05)


def test_team_estep_at_16_clusters_gives_the_tile_kernels_bits(ctx):
    """nS = 16 (configs[4]'s cluster count): the one-pass team kernel (four warps per work item, one per row tile, the posterior
    rows gathered once) against the four tile launches it replaces (PSCL_ESTEP_TILES=1): the same per-pair products; only the
    order in which a work item's 32 lane products are multiplied differs (transpose-reduction instead of butterflies), so every
    id and type is the same and the LLKs agree to rounding; cells of several work items and of a few pairs included."""
    s = synth.make_pileup(C=400, nv=16, V=9000, kbar=1500, seed=616)
    o = ctx.fmx_opts(16, early_stop=False, max_iter=4)
    team = ctx.fmx_run(s.plp, o)[0]
    os.environ["PSCL_ESTEP_TILES"] = "1"
    try:
        tiles = ctx.fmx_run(s.plp, o)[0]
    finally:
        del os.environ["PSCL_ESTEP_TILES"]
    for f in team.dtype.names:
        if team.dtype[f].kind == "f":
            assert_close(team[f], tiles[f], f, rtol=1e-12)
        else:
            assert np.array_equal(team[f], tiles[f]), f
    assert (np.bincount(team["clust"][team["clust"] >= 0], minlength=16) > 0).all()


def test_init_cluster_and_forced_iterations(ctx):
    """--init-cluster path (:198-216) with unassigned cells; no early stop (benchmark mode)."""
    s = synth.make_pileup(C=300, nv=4, V=2000, kbar=250, seed=42)
    rng = np.random.default_rng(1)
    init = np.where(rng.random(300) < 0.8, s.truth_d1, -1).astype(np.int32)
    _check(*_both(ctx, s.plp, 4, init, early_stop=False, max_iter=6))


def test_old_mode_em(ctx):
    """freemuxlet-old EM rules (cmd_cram_freemuxlet.cpp:456-653): geno_error in the last iteration only."""
    s = synth.make_pileup(C=250, nv=3, V=1500, kbar=200, seed=43)
    init = s.truth_d1.astype(np.int32)
    _check(*_both(ctx, s.plp, 3, init, mode_old=True, geno_error=0.05, max_iter=4))
    _check(*_both(ctx, s.plp, 3, init, mode_old=True, geno_error=0.0, max_iter=10))


def test_seeding_fraction_and_threshold(ctx):
    s = synth.make_pileup(C=200, nv=3, V=1500, kbar=200, seed=44)
    _check(*_both(ctx, s.plp, 3, frac_init_clust=0.5))
    sc = orc.fmx_run(s.plp, orc.fmx_opts(3, max_iter=0))["cells"]
    thres = float(np.median(sc["llk2"] - sc["llk0"]))
    _check(*_both(ctx, s.plp, 3, singlet_score_thres=thres))


def test_stage1_values(ctx):
    """per-cell llk0/llk2 (.lmix columns) incl. allele-2 reads, deep pairs, empty cells."""
    rng = np.random.default_rng(3)
    C, V = 50, 400
    counts = rng.integers(0, 80, C)
    counts[[0, 11, C - 1]] = 0
    cell_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    P = int(cell_ptr[-1])
    snp = np.concatenate([np.sort(rng.choice(V, c, replace=False)) for c in counts]).astype(np.int32)
    nrd = rng.integers(1, 5, P)
    nrd[rng.random(P) < 0.03] = 300
    prp = np.concatenate([[0], np.cumsum(nrd)]).astype(np.int64)
    N = int(prp[-1])
    from popscle_b200 import Pileup
    plp = Pileup(C, V, cell_ptr, snp, prp, rng.integers(0, 3, N).astype(np.uint8),
                 rng.integers(13, 41, N).astype(np.uint8), np.round(rng.uniform(0.05, 0.5, V), 5))
    cells, res, gl, cnt, r = _both(ctx, plp, 3, max_iter=2)
    for f in ("n_snps", "n_reads"):
        assert np.array_equal(cells[f], r["cells"][f])
    assert_close(cells["llk0"], r["cells"]["llk0"], "llk0", rtol=1e-9)
    assert_close(cells["llk2"], r["cells"]["llk2"], "llk2", rtol=1e-9)
    _check(cells, res, gl, cnt, r, tied=0.3)


def test_snp_sharded_em_matches_unsharded(ctx):
    """SNP shards + sum of the partial LLKs == one shard (here both shards run on the
    same GPU one after the other; the sum stands in for the NCCL all-reduce)."""
    import torch
    s = synth.make_pileup(C=300, nv=4, V=2000, kbar=250, seed=46)
    plp = s.plp
    nS, iters = 4, 4
    o = ctx.fmx_opts(nS, early_stop=False, max_iter=iters)
    full, _, _, _ = ctx.fmx_run(plp, o, s.truth_d1.astype(np.int32))
    from popscle_b200 import Context
    shards = [plp.slice_snps(0, 900), plp.slice_snps(900, 2000)]
    ctxs = [ctx, Context(0)]
    try:
        npairs = nS * (nS + 1) // 2
        dev = torch.device("cuda", 0)
        st = [torch.zeros(4 * plp.n_cells, dtype=torch.float64, device=dev) for _ in shards]
        llk = [torch.zeros(plp.n_cells * npairs, dtype=torch.float64, device=dev) for _ in shards]
        cl = [torch.zeros(plp.n_cells, dtype=torch.int32, device=dev) for _ in shards]
        init = torch.from_numpy(s.truth_d1.astype(np.int32)).to(dev)
        dp = []
        for c, sh, a in zip(ctxs, shards, st):
            d = c.upload(sh); dp.append(d)
            c.fmx_init(d, o)
            c.fmx_stage1(a.data_ptr())
            c.sync()
        tot = st[0] + st[1]
        for c, a, k in zip(ctxs, st, cl):
            a.copy_(tot); torch.cuda.synchronize()
            c.fmx_seed(a.data_ptr(), init.data_ptr(), k.data_ptr())
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
        check_fmx_parity(a, full)
        for f in ("type", "clust", "best_j", "best_k"):
            assert np.array_equal(a[f], full[f])
    finally:
        ctxs[1].close()


def test_fmx_errors(ctx):
    from popscle_b200 import PsclError
    s = synth.make_pileup(C=20, nv=3, V=200, kbar=60, seed=1)
    with pytest.raises(PsclError):
        ctx.fmx_run(s.plp, ctx.fmx_opts(1))    # nSamples-1 == 0 in the doublet prior (:380)
    with pytest.raises(PsclError):
        ctx.fmx_run(s.plp, ctx.fmx_opts(33))
    with pytest.raises(PsclError):
        ctx.fmx_run(s.plp, ctx.fmx_opts(3), np.full(20, 3, dtype=np.int32))  # cluster id >= nsample (:99-100)
    s.plp.snp_af = None
    with pytest.raises(PsclError):
        ctx.fmx_run(s.plp, ctx.fmx_opts(3))


@pytest.mark.parametrize("shape", [(3000, 8, 20000, 600), (600, 5, 500, 200), (900, 16, 6000, 500), (300, 20, 3000, 300), (400, 30, 3000, 300)])
@pytest.mark.parametrize("batch", [1, 7, 32, 256, 1024])
def test_batched_seeding_takes_the_serial_chains_decisions(ctx, shape, batch):
    """The speculative batches (k_fmx_seed3_*: every cell of a batch decided at once against the table before the batch, then
    proven against the merges of the batch's earlier cells) and the older batched form (k_fmx_seed_dist + k_fmx_seed_commit:
    snapshot distances + serial corrections) against the one-CTA serial chain k_fmx_seed, from sparse (few shared SNPs) to
    dense (every SNP shared) pileups, whatever the batch size."""
    C, nS, V, kbar = shape
    s = synth.make_pileup(C=C, nv=nS, V=V, kbar=kbar, seed=500 + nS)
    o = ctx.fmx_opts(nS, max_iter=0)
    os.environ["PSCL_SEED_SERIAL"] = "1"
    try:
        serial = ctx.fmx_run(s.plp, o)[0]
    finally:
        del os.environ["PSCL_SEED_SERIAL"]
    os.environ["PSCL_SEED_BATCH"] = str(batch)
    try:
        batched, _, gl, cnt = ctx.fmx_run(s.plp, o, want_clusters=True)
        if batch <= 256:
            os.environ["PSCL_SEED_V2"] = "1"
            try:
                v2 = ctx.fmx_run(s.plp, o)[0]
            finally:
                del os.environ["PSCL_SEED_V2"]
            assert np.array_equal(v2["init_clust"], serial["init_clust"])
    finally:
        del os.environ["PSCL_SEED_BATCH"]
    assert np.array_equal(batched["init_clust"], serial["init_clust"])
    assert (np.bincount(batched["init_clust"], minlength=nS) > 0).all()


def test_speculative_seeding_with_threshold_and_fraction(ctx):
    """cells below --frac-init-clust / the singlet-score threshold take no part (:225-226): they stay unassigned and their
    pileups never enter a cluster, in the speculative batches as in the serial chain."""
    s = synth.make_pileup(C=700, nv=6, V=4000, kbar=400, seed=321)
    o = ctx.fmx_opts(6, max_iter=0, frac_init_clust=0.6)
    os.environ["PSCL_SEED_SERIAL"] = "1"
    try:
        serial = ctx.fmx_run(s.plp, o)[0]
    finally:
        del os.environ["PSCL_SEED_SERIAL"]
    spec = ctx.fmx_run(s.plp, o)[0]
    assert np.array_equal(spec["init_clust"], serial["init_clust"])
    assert 250 < (spec["init_clust"] < 0).sum() < 300


def test_batched_seeding_vs_oracle_at_2000_cells(ctx):
    """VERDICT r1 (weak 1.iii): greedy seeding + EM against the oracle beyond a few hundred cells, at nS 8 and 16."""
    for nS, C, V, kbar in ((8, 2000, 20000, 500), (16, 2000, 30000, 700)):
        s = synth.make_pileup(C=C, nv=nS, V=V, kbar=kbar, seed=700 + nS)
        cells, res, _, _ = ctx.fmx_run(s.plp, ctx.fmx_opts(nS), compact=3)
        r = orc.fmx_run(s.plp, orc.fmx_opts(nS))
        check_fmx_parity(cells, r["cells"])
        assert res.n_iter == r["res"].n_iter and res.n_singlet == r["res"].n_singlet


@pytest.mark.parametrize("case", ["seed", "seed_frac_nosweeps", "refine_partial_init", "refine_keep_missing"])
def test_old_mode_pairwise_vote_seeding(ctx, case):
    """N3: freemuxlet-old's own seeding (cmd_cram_freemuxlet.cpp:165-346) — Bayes-factor trits on the device, votes on
    glibc's default rand() stream — against the oracle's restatement (which reproduces the reference's files, test_golden)."""
    nS = 4
    s = synth.make_pileup(C=260, nv=nS, V=1500, kbar=260, seed=880)
    init = None
    kw = dict(mode_old=True, geno_error=0.05, early_stop=False)
    if case == "seed":
        kw.update(iter_init=10)
    elif case == "seed_frac_nosweeps":
        kw.update(iter_init=0, frac_init_clust=0.5, bf_thres=4.0)
    else:
        rng = np.random.default_rng(3)
        init = np.where(rng.random(260) < 0.6, s.truth_d1, -1).astype(np.int32)
        init[rng.random(260) < 0.1] = 2
        kw.update(iter_init=10, keep_init_missing=(case == "refine_keep_missing"))
    cells, res, gl, cnt = ctx.fmx_run(s.plp, ctx.fmx_opts(nS, **kw), init, want_clusters=True)
    r = orc.fmx_run(s.plp, orc.fmx_opts(nS, **kw), init, want_clusters=True, n_threads=8)
    assert np.array_equal(cells["init_clust"], r["cells"]["init_clust"])
    if case == "seed_frac_nosweeps":
        assert (cells["init_clust"] < 0).sum() > 100  # the droplets beyond --frac-init-clust stay unassigned
    if case == "refine_keep_missing":
        assert np.array_equal(cells["init_clust"] < 0, init < 0)
    _check(cells, res, gl, cnt, r)


def test_old_mode_seeding_sharded(ctx):
    """pscl_multi: the pairwise seeding needs every SNP of a droplet, so it runs over the whole pileup on one GPU."""
    from popscle_b200 import Multi
    s = synth.make_pileup(C=200, nv=3, V=1200, kbar=240, seed=881)
    o = ctx.fmx_opts(3, mode_old=True, geno_error=0.05, iter_init=10, early_stop=False)
    one = ctx.fmx_run(s.plp, o)[0]
    with Multi(gpu_ids=[int(x) for x in os.environ.get("PSCL_TEST_GPUS", "0,0").split(",")] * 2) as m:
        many = m.fmx_run(s.plp, o)[0]
    assert np.array_equal(many["init_clust"], one["init_clust"])
    check_fmx_parity(many, one)
