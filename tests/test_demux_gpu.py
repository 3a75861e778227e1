"""GPU parity of the demuxlet path (through the C ABI) against the CPU oracle."""
import os

import numpy as np
import pytest

import oracle_py as orc
from popscle_b200 import synth
from tests.parity import check_demux_parity, assert_close

pytestmark = pytest.mark.gpu

DEFAULT = [0.0, 0.5]


KERNELS = {"auto": 0, "lane": 1, "general": 2, "poly": 4, "dict": 6, "cls": 7}


def _run_both(ctx, s, gp, has_gp, alphas, general=False, dp=0.5):
    """general: False/True (legacy switch) or one of KERNELS' names."""
    which = KERNELS[general] if isinstance(general, str) else (2 if general else 0)
    ctx.demux_select_kernel(which)
    try:
        out, grid = ctx.demux_run(s.plp, gp, has_gp, alphas, dp, want_grid=True)
    finally:
        ctx.demux_select_kernel(0)
    ref, rgrid = orc.demux(s.plp, gp, has_gp, alphas, dp, want_grid=True, n_threads=8)
    return out, grid, ref, rgrid


@pytest.mark.parametrize("nv", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("general", ["lane", "dict", "general", "poly"])
def test_default_grid_parity(ctx, nv, general):
    s = synth.make_pileup(C=300, nv=nv, V=2000, kbar=250, seed=100 + nv)
    gp = synth.gt_to_gp(s.geno)
    out, grid, ref, rgrid = _run_both(ctx, s, gp, None, DEFAULT, general)
    n = check_demux_parity(out, grid, ref, rgrid, DEFAULT)
    assert n > 0.95 * len(out)


def test_config1_tutorial_shape(ctx):
    s = synth.make_config(1)
    gp = synth.gt_to_gp(s.geno)
    out, grid, ref, rgrid = _run_both(ctx, s, gp, None, DEFAULT)
    check_demux_parity(out, grid, ref, rgrid, DEFAULT)
    sng = out["type"] == 0
    assert (out["sng_best"][sng] == s.truth_d1[sng]).mean() > 0.99


@pytest.mark.parametrize("nv,alphas", [(4, [0.0, 0.25, 0.5]), (6, [0.0, 0.1, 0.2, 0.3, 0.4, 0.5]),
                                        (16, [0.0, 0.5]), (12, [0.0, 0.3])])
@pytest.mark.parametrize("kernel", ["auto", "general"])
def test_general_grids(ctx, nv, alphas, kernel):
    """auto = k_demux_poly for these shapes; k_demux_general is the 9-FMA baseline"""
    s = synth.make_pileup(C=120, nv=nv, V=1500, kbar=200, seed=300 + nv)
    gp = synth.gt_to_gp(s.geno)
    out, grid, ref, rgrid = _run_both(ctx, s, gp, None, alphas, kernel)
    check_demux_parity(out, grid, ref, rgrid, alphas)


def test_poly_soft_genotypes_deep_pairs_and_big_cells(ctx):
    """k_demux_poly on soft GP rows, SNPs without GP, allele-2 reads, pairs with up to 60 reads (class D) and
    cells cut into several work items, 7-point alpha grid, 20 samples (tiles with idle threads)."""
    rng = np.random.default_rng(17)
    s = synth.make_pileup(C=10, nv=20, V=20000, kbar=3000, seed=1717)
    plp = s.plp
    nrd = np.diff(plp.pair_read_ptr)
    deep = rng.random(plp.n_pairs) < 0.02
    nrd2 = np.where(deep, rng.integers(4, 60, plp.n_pairs), nrd)
    prp = np.concatenate([[0], np.cumsum(nrd2)]).astype(np.int64)
    N = int(prp[-1])
    from popscle_b200 import Pileup
    p2 = Pileup(plp.n_cells, plp.n_snps, plp.cell_ptr, plp.pair_snp, prp, rng.choice([0, 0, 0, 1, 1, 2], N).astype(np.uint8),
                rng.integers(13, 41, N).astype(np.uint8), plp.snp_af)
    s2 = synth.Synth(p2, s.geno, s.af, s.truth_d1, s.truth_d2, 1717)
    gp = rng.dirichlet([0.5, 0.5, 0.5], size=(plp.n_snps, 20)).astype(np.float32).astype(np.float64)
    has = (rng.random(plp.n_snps) > 0.2).astype(np.uint8)
    alphas = [0.0, 0.05, 0.1, 0.2, 0.3, 0.4, 0.5]
    out, grid, ref, rgrid = _run_both(ctx, s2, gp, has, alphas, "poly")
    check_demux_parity(out, grid, ref, rgrid, alphas)
    out, grid, ref, rgrid = _run_both(ctx, s2, gp, has, alphas, "general")
    check_demux_parity(out, grid, ref, rgrid, alphas)


def test_config4_shape_small(ctx):
    """nv=64, 21-point alpha grid (config 4's shape) on a few cells."""
    alphas = [0.025 * i for i in range(21)]
    s = synth.make_pileup(C=12, nv=64, V=3000, kbar=150, seed=404)
    gp = synth.gt_to_gp(s.geno)
    for kernel in ("auto", "general"):
        out, grid, ref, rgrid = _run_both(ctx, s, gp, None, alphas, kernel)
        check_demux_parity(out, grid, ref, rgrid, alphas)


def test_missing_genotypes_and_other_alleles(ctx):
    """SNPs without GP contribute nothing but count in NUM.SNPS (SURVEY 8a note 8); allele 2 reads
    are skipped (note 5); soft GP rows."""
    s = synth.make_pileup(C=150, nv=5, V=1200, kbar=200, seed=55)
    rng = np.random.default_rng(5)
    gp = rng.dirichlet([0.4, 0.4, 0.4], size=(s.plp.n_snps, 5)).astype(np.float32).astype(np.float64)
    has = (rng.random(s.plp.n_snps) > 0.3).astype(np.uint8)
    s.plp.read_allele[rng.random(s.plp.n_reads) < 0.2] = 2
    for general in ("lane", "dict", "general", "poly"):
        out, grid, ref, rgrid = _run_both(ctx, s, gp, has, DEFAULT, general)
        check_demux_parity(out, grid, ref, rgrid, DEFAULT)


def test_deep_pairs_and_empty_cells(ctx):
    """pairs with hundreds of reads (underflow guard), cells without any pair, 1-pair cells."""
    rng = np.random.default_rng(9)
    C, V, nv = 40, 300, 4
    counts = rng.integers(0, 60, C)
    counts[[0, 7, C - 1]] = 0
    counts[3] = 1
    cell_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    P = int(cell_ptr[-1])
    snp = np.concatenate([np.sort(rng.choice(V, c, replace=False)) for c in counts]).astype(np.int32)
    nrd = rng.integers(1, 6, P)
    nrd[rng.random(P) < 0.05] = 400
    prp = np.concatenate([[0], np.cumsum(nrd)]).astype(np.int64)
    N = int(prp[-1])
    from popscle_b200 import Pileup
    plp = Pileup(C, V, cell_ptr, snp, prp, rng.integers(0, 3, N).astype(np.uint8), rng.integers(13, 41, N).astype(np.uint8),
                 rng.uniform(0.05, 0.5, V))
    geno = rng.integers(0, 3, (nv, V)).astype(np.int8)
    gp = synth.gt_to_gp(geno)
    s = synth.Synth(plp, geno, plp.snp_af, None, None, 9)
    for general in ("lane", "dict", "general", "poly"):
        out, grid, ref, rgrid = _run_both(ctx, s, gp, None, DEFAULT, general)
        check_demux_parity(out, grid, ref, rgrid, DEFAULT, allow_tied_frac=0.06)  # 2 of 35: the one-pair cell and one 400-read cell are true ties


@pytest.mark.parametrize("kernel", ["lane", "dict"])
def test_sharding_is_bit_identical(ctx, kernel):
    """barcode shards reproduce the unsharded records bit for bit; so do partial-grid batches."""
    s = synth.make_pileup(C=400, nv=8, V=3000, kbar=300, seed=77)
    gp = synth.gt_to_gp(s.geno)
    ctx.demux_select_kernel(KERNELS[kernel])
    full = ctx.demux_run(s.plp, gp, None, DEFAULT)
    parts = [ctx.demux_run(s.plp.slice_cells(a, b), gp, None, DEFAULT) for a, b in ((0, 130), (130, 131), (131, 400))]
    assert np.concatenate(parts).tobytes() == full.tobytes()
    ctx.set_partial_budget(1 << 20)
    try:
        d = ctx.upload(s.plp)
        ctx.demux_set_geno(gp, None, s.plp.n_snps)
        ctx.demux_score(d, DEFAULT, 0.5, 50, 333)
        sub = ctx.demux_fetch()
    finally:
        ctx.set_partial_budget(1 << 30)
        ctx.demux_select_kernel(0)
    assert sub.tobytes() == full[50:333].tobytes()


def test_symmetry_properties(ctx):
    """LLK[j,k,a] == LLK[k,j,1-a]; permuting samples permutes the grid."""
    alphas = [0.0, 0.25, 0.5, 0.75]
    s = synth.make_pileup(C=60, nv=5, V=800, kbar=150, seed=21)
    gp = synth.gt_to_gp(s.geno)
    _, grid = ctx.demux_run(s.plp, gp, None, alphas, want_grid=True)
    for j in range(5):
        for k in range(5):
            if j != k:
                assert_close(grid[:, j, k, 1], grid[:, k, j, 3], "alpha symmetry", rtol=1e-9)
                assert_close(grid[:, j, k, 2], grid[:, k, j, 2], "alpha 0.5 symmetry", rtol=1e-9)
    perm = np.array([3, 0, 4, 1, 2])
    _, grid2 = ctx.demux_run(s.plp, gp[:, perm, :], None, alphas, want_grid=True)
    for n in range(1, 4):
        a = grid2[:, :, :, n]
        b = grid[:, perm][:, :, perm][:, :, :, n]
        off = ~np.eye(5, dtype=bool)
        assert_close(a[:, off], b[:, off], "sample permutation", rtol=1e-9)


def test_errors(ctx):
    from popscle_b200 import PsclError
    s = synth.make_pileup(C=20, nv=3, V=200, kbar=60, seed=1)
    gp = synth.gt_to_gp(s.geno)
    with pytest.raises(PsclError):
        ctx.demux_run(s.plp, gp, None, [0.0])           # nAlpha == 1 divides by zero in the reference
    with pytest.raises(PsclError):
        ctx.demux_run(s.plp, gp[:, :1, :], None, DEFAULT)  # nv == 1
    bad = synth.make_pileup(C=20, nv=3, V=200, kbar=60, seed=1).plp
    bad.read_qual[0] = 99
    with pytest.raises(PsclError):
        ctx.demux_run(bad, gp, None, DEFAULT)


@pytest.mark.parametrize("kernel", ["lane", "dict"])
def test_cells_larger_than_one_work_item(ctx, kernel):
    """cells with > 2048 pairs are cut into several work items whose partial grids are summed in item order."""
    s = synth.make_pileup(C=24, nv=8, V=30000, kbar=5000, seed=88)
    assert np.diff(s.plp.cell_ptr).max() > 4096
    gp = synth.gt_to_gp(s.geno)
    out, grid, ref, rgrid = _run_both(ctx, s, gp, None, DEFAULT, kernel)
    check_demux_parity(out, grid, ref, rgrid, DEFAULT)


def test_genotype_dictionary_is_bit_identical_and_falls_back(ctx):
    """k_demux_default reads 8-bit dictionary codes instead of genotype rows when the table holds <= 256 distinct triples
    (hard calls: 3 per combination of genotype counts); the records are bit-identical to the row-gather kernel's, SNPs
    without GP included.  A table with more triples (soft GP, several error rates) silently takes the rows."""
    s = synth.make_pileup(C=200, nv=7, V=2500, kbar=300, seed=606)
    rng = np.random.default_rng(6)
    has = (rng.random(s.plp.n_snps) > 0.2).astype(np.uint8)

    ran = []

    def run(gp, which, has_gp=None):
        ctx.demux_select_kernel(which)
        try:
            out = ctx.demux_run(s.plp, gp, has_gp, DEFAULT, want_grid=True)
            ran.append(ctx.demux_last_kernel())
            return out
        finally:
            ctx.demux_select_kernel(0)

    gp = synth.gt_to_gp(s.geno)
    assert 3 < len(np.unique(gp.reshape(-1, 3), axis=0)) <= 256
    for h in (None, has):
        lane, lgrid = run(gp, 1, h)
        for which in (0, 6):
            rec, grid = run(gp, which, h)
            assert ran[-1] in ((6,) if which == 6 else (6, 7))
            assert rec.tobytes() == lane.tobytes() and np.array_equal(grid, lgrid, equal_nan=True)
    # a missing call replaced by a flat prior adds one triple: still coded
    gp4 = gp.copy()
    gp4[rng.random(gp4.shape[:2]) < 0.1] = [0.3, 0.3, 0.4]
    lane, lgrid = run(gp4, 1)
    rec, grid = run(gp4, 0)
    assert ran[-2:] == [1, 6]  # (a fourth triple on a SNP: not the genotype-class variant)
    assert rec.tobytes() == lane.tobytes() and np.array_equal(grid, lgrid, equal_nan=True)
    # exactly 256 distinct triples still fit, 257 do not; either way nothing changes in the output
    for n_trip in (256, 257):
        gpe = gp.copy()
        flat = gpe.reshape(-1, 3)
        flat[:] = flat[0]
        extra = rng.dirichlet([1.0, 1.0, 1.0], size=n_trip - 1)
        flat[rng.permutation(len(flat))[: 40 * (n_trip - 1)].reshape(40, -1)] = extra[None, :, :]
        assert len(np.unique(flat, axis=0)) == n_trip
        lane, lgrid = run(gpe, 1)
        rec, grid = run(gpe, 0)
        assert ran[-1] in ((6, 7) if n_trip == 256 else (1,))
        assert rec.tobytes() == lane.tobytes() and np.array_equal(grid, lgrid, equal_nan=True)
    soft = rng.dirichlet([0.4, 0.4, 0.4], size=(s.plp.n_snps, 7))
    lane, lgrid = run(soft, 1)
    rec, grid = run(soft, 6)
    assert ran[-1] == 1
    assert rec.tobytes() == lane.tobytes() and np.array_equal(grid, lgrid, equal_nan=True)


def test_raw_genotype_forms_are_mixed_on_the_device_bit_for_bit(ctx):
    """ABI 4: hard calls (uint8) or the reader's float32 posteriors plus the genotype error rate go in, and the library
    does the mixing of sc_drop_seq.cpp:287-315 on the device; the records are bit-identical to those from the table the
    host mixes (synth.gt_to_gp restates those lines).  Per-SNP error rates (--geno-error-coeff), SNPs without GP,
    clamping at 0.999, err = 0, soft posteriors, and an invalid hard-call code are covered."""
    from popscle_b200 import PsclError, RawGeno
    s = synth.make_pileup(C=150, nv=6, V=1500, kbar=200, seed=808)
    rng = np.random.default_rng(8)
    V, nv = s.plp.n_snps, 6
    gt8 = np.ascontiguousarray(s.geno.T.astype(np.uint8))
    onehot = np.zeros((V, nv, 3), dtype=np.float32)
    onehot[np.arange(V)[:, None], np.arange(nv)[None, :], s.geno.T.astype(np.int64)] = 1.0
    has = (rng.random(V) > 0.2).astype(np.uint8)
    for err in (0.1, 0.0, 0.02, 5.0):
        want = ctx.demux_run(s.plp, synth.gt_to_gp(s.geno, err), None, DEFAULT, want_grid=True)
        for raw in (RawGeno(gt8=gt8, err=err), RawGeno(gp_f32=onehot, err=err)):
            got = ctx.demux_run(s.plp, raw, None, DEFAULT, want_grid=True)
            assert got[0].tobytes() == want[0].tobytes() and np.array_equal(got[1], want[1], equal_nan=True), err
    want = ctx.demux_run(s.plp, synth.gt_to_gp(s.geno), has, DEFAULT)
    assert ctx.demux_run(s.plp, RawGeno(gt8=gt8), has, DEFAULT).tobytes() == want.tobytes()
    # per-SNP error rates, soft float32 posteriors, a non-default alpha grid (k_demux_poly)
    err_snp = rng.uniform(-0.05, 1.2, V)
    soft = rng.dirichlet([0.5, 0.5, 0.5], size=(V, nv)).astype(np.float32)
    mixed = np.empty((V, nv, 3))
    for v in range(V):  # sc_drop_seq.cpp:288-315 restated per SNP
        avg = np.full(3, 1e-10)
        for j in range(nv):
            avg = avg + soft[v, j].astype(np.float64)
        avg = avg / (avg[0] + avg[1] + avg[2])
        e = min(max(err_snp[v], 0.0), 0.999)
        mixed[v] = (1 - e) * soft[v].astype(np.float64) + e * avg if e > 0 else soft[v]
    for alphas in (DEFAULT, [0.0, 0.2, 0.5]):
        want = ctx.demux_run(s.plp, mixed, None, alphas)
        got = ctx.demux_run(s.plp, RawGeno(gp_f32=soft, err_snp=err_snp), None, alphas)
        assert got.tobytes() == want.tobytes()
    bad = gt8.copy()
    bad[17, 3] = 3
    with pytest.raises(PsclError):
        ctx.demux_run(s.plp, RawGeno(gt8=bad), None, DEFAULT)
    ctx.demux_run(s.plp, RawGeno(gt8=gt8), None, DEFAULT)  # the context recovers


def test_compact_pileup_inputs_are_equivalent(ctx):
    """ABI 2 compact host arrays (u32 read offsets, allele<<6|qual) give bit-identical records, for demuxlet and
    freemuxlet; a byte with allele code 3 is rejected."""
    from popscle_b200 import PsclError
    s = synth.make_pileup(C=150, nv=6, V=1500, kbar=200, seed=4242)
    gp = synth.gt_to_gp(s.geno)
    a = ctx.demux_run(s.plp, gp, None, DEFAULT)
    b = ctx.demux_run(s.plp, gp, None, DEFAULT, compact=True)
    assert a.tobytes() == b.tobytes()
    fa = ctx.fmx_run(s.plp, ctx.fmx_opts(3))[0]
    fb = ctx.fmx_run(s.plp, ctx.fmx_opts(3), compact=True)[0]
    assert fa.tobytes() == fb.tobytes()
    bad = synth.make_pileup(C=20, nv=3, V=200, kbar=60, seed=1).plp
    bad.compact()[1][5] |= 0xC0
    with pytest.raises(PsclError):
        ctx.demux_run(bad, synth.gt_to_gp(synth.make_pileup(C=20, nv=3, V=200, kbar=60, seed=1).geno), None, DEFAULT, compact=True)


def test_delta_coded_pileup_inputs_are_equivalent(ctx):
    """ABI 3 delta-coded pair arrays (first SNP per cell, 16-bit gaps, 8-bit base-call counts) decode on the device to the
    same image: bit-identical records for demuxlet and freemuxlet, with empty cells and a cell larger than one work
    item; deltas that leave [0, n_snps) and counts that do not sum to n_reads are rejected."""
    from popscle_b200 import PsclError
    s = synth.make_pileup(C=40, nv=6, V=30000, kbar=1500, seed=4243)
    assert np.diff(s.plp.cell_ptr).max() > 2048
    q = s.plp
    cp = np.insert(np.insert(q.cell_ptr, 3, q.cell_ptr[3]), 3, q.cell_ptr[3])  # two empty cells in the middle, one at the end
    cp = np.append(cp, cp[-1])
    plp = type(q)(len(cp) - 1, q.n_snps, cp, q.pair_snp, q.pair_read_ptr, q.read_allele, q.read_qual, q.snp_af)
    assert plp.compact3() is not None
    gp = synth.gt_to_gp(s.geno)
    a = ctx.demux_run(plp, gp, None, DEFAULT)
    b = ctx.demux_run(plp, gp, None, DEFAULT, compact=3)
    assert a.tobytes() == b.tobytes()
    fa = ctx.fmx_run(plp, ctx.fmx_opts(3))[0]
    fb = ctx.fmx_run(plp, ctx.fmx_opts(3), compact=3)[0]
    assert fa.tobytes() == fb.tobytes()
    # staged run: the SNP gaps cross in slices of whole cells on a second stream, each slice decoded and scored as it lands
    try:
        for n in ("2", "3", "8", "50"):
            os.environ["PSCL_STAGES"] = n
            assert ctx.demux_run(plp, gp, None, DEFAULT, compact=3).tobytes() == a.tobytes(), n
        bad = synth.make_pileup(C=20, nv=3, V=200, kbar=60, seed=1)
        bad.plp.compact3()[1][bad.plp.cell_ptr[15] + 1] = 60000
        with pytest.raises(PsclError):
            ctx.demux_run(bad.plp, synth.gt_to_gp(bad.geno), None, DEFAULT, compact=3)
    finally:
        os.environ.pop("PSCL_STAGES", None)
    empty = synth.make_pileup(C=4, nv=2, V=50, kbar=20, seed=2)
    empty.plp.cell_ptr[:] = 0
    e = type(empty.plp)(4, 50, empty.plp.cell_ptr, empty.plp.pair_snp[:0], np.zeros(1, np.int64), empty.plp.read_allele[:0], empty.plp.read_qual[:0], None)
    assert ctx.demux_run(e, synth.gt_to_gp(empty.geno), None, DEFAULT, compact=3).tobytes() == ctx.demux_run(e, synth.gt_to_gp(empty.geno), None, DEFAULT).tobytes()
    for what in ("delta", "count"):
        bad = synth.make_pileup(C=20, nv=3, V=200, kbar=60, seed=1)
        first, d16, n8 = bad.plp.compact3()
        if what == "delta":
            d16[bad.plp.cell_ptr[5] + 1] = 60000
        else:
            n8[7] += 1
        with pytest.raises(PsclError):
            ctx.demux_run(bad.plp, synth.gt_to_gp(bad.geno), None, DEFAULT, compact=3)


@pytest.mark.parametrize("nv,na", [(2, 2), (3, 3), (9, 5), (17, 9), (33, 17), (5, 32)])
def test_poly_shapes(ctx, nv, na):
    """every plane-count bucket of k_demux_poly (4/8/16/21/32) and sample counts that leave idle tile threads"""
    alphas = [0.5 * i / (na - 1) for i in range(na)]
    s = synth.make_pileup(C=30, nv=nv, V=900, kbar=150, seed=500 + nv)
    gp = synth.gt_to_gp(s.geno)
    out, grid, ref, rgrid = _run_both(ctx, s, gp, None, alphas, "poly")
    check_demux_parity(out, grid, ref, rgrid, alphas)


@pytest.mark.parametrize("nv", [2, 3, 5, 8])
def test_genotype_classes_are_bit_identical_and_fall_back(ctx, nv):
    """The genotype-class variant of the dictionary kernel (a SNP of a hard-call table holds at most three distinct triples, so a
    pair's factors are computed once per class and picked per accumulator) evaluates the same expressions on the same doubles:
    records and grids are bit-identical to the dictionary and row-gather kernels', SNPs without GP included.  A SNP with a
    fourth triple (a missing call among hard ones) sends the whole table to the dictionary kernel."""
    s = synth.make_pileup(C=150, nv=nv, V=3000, kbar=400, seed=700 + nv)
    rng = np.random.default_rng(nv)
    has = (rng.random(s.plp.n_snps) > 0.2).astype(np.uint8)
    ran = []

    def run(gp, which, has_gp=None, compact=None):
        ctx.demux_select_kernel(which)
        try:
            out = ctx.demux_run(s.plp, gp, has_gp, DEFAULT, want_grid=True) if compact is None else \
                ctx.demux_run(s.plp, gp, has_gp, DEFAULT, want_grid=True, compact=compact)
            ran.append(ctx.demux_last_kernel())
            return out
        finally:
            ctx.demux_select_kernel(0)

    gp = synth.gt_to_gp(s.geno)
    for h in (None, has):
        lane, lgrid = run(gp, 1, h)
        dic, dgrid = run(gp, 6, h)
        rec, grid = run(gp, 7, h)
        assert ran[-3:] == [1, 6, 7]
        assert rec.tobytes() == dic.tobytes() == lane.tobytes()
        assert np.array_equal(grid, dgrid, equal_nan=True) and np.array_equal(grid, lgrid, equal_nan=True)
    # every compact host form ends in the same arrays on the device
    base, _ = run(gp, 6)
    for compact in (3, 6):
        rec2, _ = run(gp, 7, None, compact)
        assert ran[-1] == 7 and rec2.tobytes() == base.tobytes(), compact
    # monomorphic SNPs (one class) and two-class SNPs are part of any table; force a few of each
    g1 = s.geno.copy()
    g1[:50] = 0
    g1[50:100, 1:] = 2
    gp1 = synth.gt_to_gp(g1)
    lane, _ = run(gp1, 1)
    rec, _ = run(gp1, 7)
    assert ran[-1] == 7 and rec.tobytes() == lane.tobytes()
    if nv >= 4:
        gp4 = gp.copy()
        gp4[rng.random(gp4.shape[:2]) < 0.1] = [0.3, 0.3, 0.4]
        assert max(len(np.unique(r, axis=0)) for r in gp4[:500]) == 4
        lane, lgrid = run(gp4, 1)
        rec, grid = run(gp4, 7)
        assert ran[-1] == 6
        assert rec.tobytes() == lane.tobytes() and np.array_equal(grid, lgrid, equal_nan=True)


@pytest.mark.parametrize("kernel", ["lane", "dict", "poly", "general"])
def test_degenerate_pileups(ctx, kernel):
    """one cell with one pair; no pair at all; no SNP with genotypes"""
    from popscle_b200 import Pileup
    nv, V = 4, 50
    rng = np.random.default_rng(3)
    gp = synth.gt_to_gp(rng.integers(0, 3, (nv, V)).astype(np.int8))
    one = Pileup(1, V, [0, 1], [7], [0, 2], np.array([0, 1], np.uint8), np.array([30, 20], np.uint8), rng.uniform(0.1, 0.5, V))
    s1 = synth.Synth(one, None, None, None, None, 0)
    out, grid, ref, rgrid = _run_both(ctx, s1, gp, None, DEFAULT, kernel)
    check_demux_parity(out, grid, ref, rgrid, DEFAULT, allow_tied_frac=1.0)  # ONE cell with ONE pair: everything ties
    empty = Pileup(3, V, [0, 0, 0, 0], np.zeros(0, np.int32), [0], np.zeros(0, np.uint8), np.zeros(0, np.uint8), rng.uniform(0.1, 0.5, V))
    ctx.demux_select_kernel(KERNELS[kernel])
    try:
        o = ctx.demux_run(empty, gp, None, DEFAULT)
    finally:
        ctx.demux_select_kernel(0)
    assert len(o) == 3 and (o["n_snps"] == 0).all()
    s = synth.make_pileup(C=20, nv=nv, V=V, kbar=20, seed=5)
    out, grid, ref, rgrid = _run_both(ctx, s, gp, np.zeros(V, np.uint8), DEFAULT, kernel)
    assert_close(out["sng_best_llk"], ref["sng_best_llk"], "no genotypes: singlet LLK")
    assert_close(out["dbl_best_llk"], ref["dbl_best_llk"], "no genotypes: doublet LLK")


@pytest.mark.parametrize("slices,groups", [(2, 1), (3, 3), (6, 2), (7, 3), (40, 5)])
def test_pipelined_run_is_bit_identical(ctx, slices, groups):
    """The default form of pscl_demux_run at size: every copy on the copy stream (counts, base-calls, then the SNP gaps in
    slices of whole cells with an event behind each), gaps decoded slice by slice, cells scored in groups of slices.  Cells
    are independent (cmd_cram_demuxlet.cpp:636) and a cell's pairs are summed in the same order: the records must be the
    bytes of the one-shot run, for hard calls (dictionary kernel) and soft posteriors (row gather), whatever the cuts;
    malformed gaps are still an error, and the context survives."""
    from popscle_b200 import PsclError, RawGeno
    s = synth.make_pileup(C=700, nv=8, V=6000, kbar=900, seed=1234 + slices)
    hard = RawGeno(gt8=np.ascontiguousarray(s.geno.T.astype(np.uint8)), err=0.1)
    rng = np.random.default_rng(slices)
    soft = rng.dirichlet([0.3, 0.3, 0.3], size=(6000, 8))
    for gp in (hard, soft):
        for compact in (3, 4):
            ref = ctx.demux_run(s.plp, gp, None, DEFAULT, compact=compact)
            os.environ["PSCL_SLICES"], os.environ["PSCL_GROUPS"] = str(slices), str(groups)
            try:
                got = ctx.demux_run(s.plp, gp, None, DEFAULT, compact=compact)
                # ... and with the counts and base-calls sliced too (ABI 7's cell_read_ptr; one k_decode_cells launch per slice)
                os.environ["PSCL_SLICE_FULL"] = "1"
                full = ctx.demux_run(s.plp, gp, None, DEFAULT, compact=compact)
                del os.environ["PSCL_SLICE_FULL"]
                # ... and with only the base-calls sliced beside the gaps (ordinary decoders on a slice's ranges)
                os.environ["PSCL_SLICE_READS"] = "1"
                reads = ctx.demux_run(s.plp, gp, None, DEFAULT, compact=compact)
            finally:
                del os.environ["PSCL_SLICES"], os.environ["PSCL_GROUPS"]
                os.environ.pop("PSCL_SLICE_FULL", None)
                os.environ.pop("PSCL_SLICE_READS", None)
            assert got.tobytes() == ref.tobytes(), (type(gp).__name__, compact)
            assert full.tobytes() == ref.tobytes(), (type(gp).__name__, compact, "full")
            assert reads.tobytes() == ref.tobytes(), (type(gp).__name__, compact, "reads")
    bad = synth.make_pileup(C=60, nv=3, V=400, kbar=90, seed=1)
    first, d8, gbig, cbp, n2, nbig, nbp = bad.plp.compact4()
    os.environ["PSCL_SLICES"] = "3"
    try:
        os.environ["PSCL_SLICE_FULL"] = "1"
        try:  # a count that disagrees with the cells' base-call offsets
            n2[bad.plp.cell_ptr[30] // 4] ^= 0x3 << (2 * (bad.plp.cell_ptr[30] % 4))
            with pytest.raises(PsclError):
                ctx.demux_run(bad.plp, synth.gt_to_gp(bad.geno), None, DEFAULT, compact=4)
            n2[bad.plp.cell_ptr[30] // 4] ^= 0x3 << (2 * (bad.plp.cell_ptr[30] % 4))
        finally:
            del os.environ["PSCL_SLICE_FULL"]
        d8[bad.plp.cell_ptr[45] + 1] = 254  # a gap that walks the SNP id past n_snps
        d8[bad.plp.cell_ptr[45] + 2] = 254
        with pytest.raises(PsclError):
            ctx.demux_run(bad.plp, synth.gt_to_gp(bad.geno), None, DEFAULT, compact=4)
        ok = synth.make_pileup(C=60, nv=3, V=400, kbar=90, seed=2)
        assert ctx.demux_run(ok.plp, synth.gt_to_gp(ok.geno), None, DEFAULT, compact=4).tobytes() == \
            ctx.demux_run(ok.plp, synth.gt_to_gp(ok.geno), None, DEFAULT).tobytes()
    finally:
        del os.environ["PSCL_SLICES"]


def test_staged_run_falls_back_when_the_partial_grids_do_not_fit(ctx):
    """a staged image can only be scored in one batch; with a small scratch budget pscl_demux_run must take
    the plain upload instead of failing with PSCL_ESTATE."""
    s = synth.make_pileup(C=3000, nv=8, V=4000, kbar=300, seed=91)
    raw_gp = synth.gt_to_gp(s.geno)
    ref = ctx.demux_run(s.plp, raw_gp, None, DEFAULT, compact=3)
    os.environ["PSCL_STAGES"] = "3"
    try:
        staged = ctx.demux_run(s.plp, raw_gp, None, DEFAULT, compact=3)
        ctx.set_partial_budget(1 << 20)  # 1 MiB: 1024 work items per batch, the pileup has 3000
        small = ctx.demux_run(s.plp, raw_gp, None, DEFAULT, compact=3)
    finally:
        del os.environ["PSCL_STAGES"]
        ctx.set_partial_budget(1 << 30)
    assert staged.tobytes() == ref.tobytes() and small.tobytes() == ref.tobytes()


def test_a_slice_that_never_arrives_is_reported_and_the_context_survives(built):
    """Fault injection: the last slice's flag is withheld (PSCL_FAULT=drop_stage_flag); the staged kernel gives up after
    PSCL_STAGE_TIMEOUT_MS, pscl_demux_run returns PSCL_ECUDA, and the same context scores the next call correctly."""
    from popscle_b200 import Context, PsclError
    s = synth.make_pileup(C=400, nv=4, V=2000, kbar=200, seed=92)
    gp = synth.gt_to_gp(s.geno)
    os.environ["PSCL_STAGE_TIMEOUT_MS"] = "50"
    try:
        c = Context(0)
    finally:
        del os.environ["PSCL_STAGE_TIMEOUT_MS"]
    try:
        good = c.demux_run(s.plp, gp, None, DEFAULT, compact=3)
        os.environ["PSCL_STAGES"] = "3"
        os.environ["PSCL_FAULT"] = "drop_stage_flag"
        try:
            with pytest.raises(PsclError) as ei:
                c.demux_run(s.plp, gp, None, DEFAULT, compact=3)
            assert ei.value.code == -3 and "never reached the device" in str(ei.value)
        finally:
            del os.environ["PSCL_FAULT"]
        again = c.demux_run(s.plp, gp, None, DEFAULT, compact=3)  # staged, flags intact
        del os.environ["PSCL_STAGES"]
        assert again.tobytes() == good.tobytes()
    finally:
        os.environ.pop("PSCL_STAGES", None)
        c.close()


def test_device_memory_exhaustion_is_reported_and_the_context_survives(built):
    """Fault injection (pscl_debug_fail_alloc): an allocation failing anywhere inside a run comes back as PSCL_ENOMEM,
    and the same context then produces the right records again."""
    from popscle_b200 import Context, PsclError
    s = synth.make_pileup(C=400, nv=4, V=2000, kbar=200, seed=92)
    gp = synth.gt_to_gp(s.geno)
    with Context(0) as c:
        good = c.demux_run(s.plp, gp, None, DEFAULT, compact=3)
        hit = 0
        for nth in range(1, 60):
            c.debug_fail_alloc(nth)
            try:
                got = c.demux_run(s.plp, gp, None, DEFAULT, compact=3)
            except PsclError as e:
                assert e.code == -4, (nth, str(e))
                hit += 1
                continue
            finally:
                c.debug_fail_alloc(0)
            assert got.tobytes() == good.tobytes()
            break
        assert hit >= 8
        assert c.demux_run(s.plp, gp, None, DEFAULT, compact=3).tobytes() == good.tobytes()
        o = c.fmx_opts(4)
        fgood = c.fmx_run(s.plp, o)[0]
        hit = 0
        for nth in (1, 5, 9, 14, 20, 27, 33):
            c.debug_fail_alloc(nth)
            try:
                c.fmx_run(s.plp, o)
            except PsclError as e:
                assert e.code == -4, (nth, str(e))
                hit += 1
            finally:
                c.debug_fail_alloc(0)
        assert hit >= 5
        assert c.fmx_run(s.plp, o)[0].tobytes() == fgood.tobytes()


def test_tiny_pileup_form_abi6_is_equivalent(ctx):
    """ABI 6: 8-bit SNP gaps and 2-bit base-call counts with the large values on the side (1.25 B per pair) — decoded on the
    device (k_decode_snp8 / k_expand_counts2) or, staged, inside the scoring kernel — give the records of the wide arrays bit
    for bit: sparse SNPs (many large gaps), deep pairs (large counts), cells cut into several work items, freemuxlet."""
    from popscle_b200 import PsclError
    s = synth.make_pileup(C=260, nv=8, V=120000, kbar=700, seed=96)  # mean gap 170: a third of the gaps are markers
    plp = s.plp
    rng = np.random.default_rng(4)
    nrd = np.diff(plp.pair_read_ptr)
    nrd = np.where(rng.random(plp.n_pairs) < 0.03, rng.integers(4, 200, plp.n_pairs), nrd)
    prp = np.concatenate([[0], np.cumsum(nrd)]).astype(np.int64)
    N = int(prp[-1])
    from popscle_b200 import Pileup
    p2 = Pileup(plp.n_cells, plp.n_snps, plp.cell_ptr, plp.pair_snp, prp, rng.choice([0, 0, 1, 2], N).astype(np.uint8),
                rng.integers(13, 41, N).astype(np.uint8), plp.snp_af)
    first, d8, gbig, cbp, n2, nbig, nbp = p2.compact4()
    assert len(gbig) > 1000 and len(nbig) > 1000
    gp = synth.gt_to_gp(s.geno)
    wide = ctx.demux_run(p2, gp, None, DEFAULT)
    assert ctx.demux_run(p2, gp, None, DEFAULT, compact=4).tobytes() == wide.tobytes()
    for stages in ("1", "2", "5", "16"):
        os.environ["PSCL_STAGES"] = stages
        try:
            assert ctx.demux_run(p2, gp, None, DEFAULT, compact=4).tobytes() == wide.tobytes(), stages
        finally:
            del os.environ["PSCL_STAGES"]
    # a cell of 7000 pairs: later work items start inside the cell and count the markers before them
    s3 = synth.make_pileup(C=6, nv=5, V=200000, kbar=7000, seed=97)
    gp3 = synth.gt_to_gp(s3.geno)
    w3 = ctx.demux_run(s3.plp, gp3, None, DEFAULT)
    os.environ["PSCL_STAGES"] = "3"
    try:
        assert ctx.demux_run(s3.plp, gp3, None, DEFAULT, compact=4).tobytes() == w3.tobytes()
    finally:
        del os.environ["PSCL_STAGES"]
    # other kernels and freemuxlet decode it on upload
    al = [0.0, 0.25, 0.5]
    assert ctx.demux_run(p2, gp, None, al, compact=4).tobytes() == ctx.demux_run(p2, gp, None, al).tobytes()
    fa = ctx.fmx_run(p2, ctx.fmx_opts(4, max_iter=2, early_stop=False))[0]
    fb = ctx.fmx_run(p2, ctx.fmx_opts(4, max_iter=2, early_stop=False), compact=4)[0]
    assert fa.tobytes() == fb.tobytes()
    # malformed side lists are input errors, not crashes
    for what in ("gap", "count", "short"):
        bad = synth.make_pileup(C=30, nv=3, V=50000, kbar=80, seed=98)
        first, d8, gbig, cbp, n2, nbig, nbp = bad.plp.compact4()
        if what == "gap":
            gbig[3] = 60000
        elif what == "count":
            n2[2] ^= 0x03
        else:
            cbp[-1] -= 1
        with pytest.raises(PsclError):
            ctx.demux_run(bad.plp, synth.gt_to_gp(bad.geno), None, DEFAULT, compact=4)
