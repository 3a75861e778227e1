"""pscl_multi (several GPUs in one process) through the C ABI.  On the one-GPU test box the same device is
listed twice or three times: the sharding, the host threads, the peer-memory all-reduce and the rank-0 seeding are the
same code that runs on distinct GPUs (NVLink only changes where the peer pointers point)."""
import os

import numpy as np
import pytest

import oracle_py as orc
from popscle_b200 import Multi, RawGeno, synth
from tests.parity import check_demux_parity, check_fmx_parity, assert_close

pytestmark = pytest.mark.gpu
DEFAULT = [0.0, 0.5]


def _ids(n):
    """device list of n entries: the same GPU n times on a one-GPU box, or PSCL_TEST_GPUS=0,1,... cycled on a real multi-GPU box"""
    have = [int(x) for x in os.environ.get("PSCL_TEST_GPUS", "0").split(",")]
    return [have[i % len(have)] for i in range(n)]


@pytest.fixture(scope="module")
def multi3(built):
    m = Multi(gpu_ids=_ids(3))
    yield m
    m.close()


@pytest.mark.parametrize("compact", [False, True, 3])
def test_barcode_sharded_demuxlet_is_bit_identical(ctx, multi3, compact):
    s = synth.make_pileup(C=700, nv=6, V=3000, kbar=300, seed=61)
    gp = synth.gt_to_gp(s.geno)
    one, grid1 = ctx.demux_run(s.plp, gp, None, DEFAULT, want_grid=True, compact=compact)
    many, gridn = multi3.demux_run(s.plp, gp, None, DEFAULT, want_grid=True, compact=compact)
    assert many.tobytes() == one.tobytes() and np.array_equal(grid1, gridn, equal_nan=True)
    t = multi3.timing()
    assert t["n_gpus"] == 3 and sum(t["units"]) == s.plp.n_pairs
    assert max(t["units"]) - min(t["units"]) < 2 * 2000  # balanced by pairs: within one cell of each other
    ref, rgrid = orc.demux(s.plp, gp, None, DEFAULT, want_grid=True, n_threads=8)
    check_demux_parity(many, gridn, ref, rgrid, DEFAULT)


def test_sharded_demuxlet_other_shapes_and_raw_genotypes(ctx, multi3):
    s = synth.make_pileup(C=90, nv=12, V=1500, kbar=200, seed=62)
    al = [0.0, 0.125, 0.25, 0.375, 0.5]
    raw = RawGeno(gt8=np.ascontiguousarray(s.geno.T.astype(np.uint8)), err=0.1)
    one = ctx.demux_run(s.plp, raw, None, al, compact=3)
    many = multi3.demux_run(s.plp, raw, None, al, compact=3)
    assert many.tobytes() == one.tobytes()
    # more GPUs than cells with pairs, empty cells at the cuts
    from popscle_b200 import Pileup
    tiny = Pileup(4, 1500, [0, 0, 3, 3, 3], s.plp.pair_snp[:3], s.plp.pair_read_ptr[:4] - s.plp.pair_read_ptr[0],
                  s.plp.read_allele[:int(s.plp.pair_read_ptr[3])], s.plp.read_qual[:int(s.plp.pair_read_ptr[3])], None)
    a = ctx.demux_run(tiny, raw, None, al)
    b = multi3.demux_run(tiny, raw, None, al)
    assert a.tobytes() == b.tobytes()


def test_wide_genotype_table_travels_gpu_to_gpu(ctx, built):
    """>= 32 MB of FP64 table: one H2D copy, then peer copies along a binary tree (here: same-device copies)."""
    s = synth.make_pileup(C=60, nv=8, V=180_000, kbar=400, seed=63)
    gp = synth.gt_to_gp(s.geno)
    assert gp.nbytes >= 32 << 20
    one = ctx.demux_run(s.plp, gp, None, DEFAULT)
    with Multi(gpu_ids=_ids(5)) as m:
        many = m.demux_run(s.plp, gp, None, DEFAULT)
        assert many.tobytes() == one.tobytes()
        os.environ["PSCL_NO_GENO_TREE"] = "1"
        try:
            assert m.demux_run(s.plp, gp, None, DEFAULT).tobytes() == one.tobytes()
        finally:
            del os.environ["PSCL_NO_GENO_TREE"]


@pytest.mark.parametrize("nS", [3, 8])
def test_snp_sharded_freemuxlet_seeds_on_one_gpu_and_matches_single(ctx, multi3, nS):
    s = synth.make_pileup(C=400, nv=nS, V=4000, kbar=300, seed=70 + nS)
    o = ctx.fmx_opts(nS)
    one, r1, gl1, cnt1 = ctx.fmx_run(s.plp, o, want_clusters=True, compact=3)
    many, rn, gln, cntn = multi3.fmx_run(s.plp, o, want_clusters=True, compact=3)
    # the seeding ran over the whole pileup on the first GPU: identical clusters and identical .lmix sums
    assert np.array_equal(many["init_clust"], one["init_clust"])
    assert many["llk0"].tobytes() == one["llk0"].tobytes() and many["llk2"].tobytes() == one["llk2"].tobytes()
    assert rn.n_iter == r1.n_iter and rn.n_singlet == r1.n_singlet and rn.n_changed == r1.n_changed
    check_fmx_parity(many, one)  # LLKs differ by the summation order of the three SNP ranges only
    assert_close(many["best_llk"], one["best_llk"], "best LLK", rtol=1e-11)
    assert np.array_equal(cntn, cnt1)
    assert_close(gln, gl1, "cluster pileups", rtol=1e-9)
    ref = orc.fmx_run(s.plp, orc.fmx_opts(nS))
    check_fmx_parity(many, ref["cells"])
    t = multi3.timing()
    assert t["iters"] == rn.n_iter and sum(t["units"]) == s.plp.n_pairs and t["seed_ms"] > 0


def test_snp_sharded_freemuxlet_from_init_clusters_allreduces_stage1(ctx, multi3):
    s = synth.make_pileup(C=300, nv=4, V=3000, kbar=250, seed=75)
    init = s.truth_d1.astype(np.int32).copy()
    init[::7] = -1
    o = ctx.fmx_opts(4, early_stop=False, max_iter=3)
    one, r1, _, _ = ctx.fmx_run(s.plp, o, init)
    many, rn, _, _ = multi3.fmx_run(s.plp, o, init)
    check_fmx_parity(many, one)
    assert_close(many["llk0"], one["llk0"], "stage-1 llk0 (all-reduced)", rtol=1e-12)
    assert np.array_equal(many["n_reads"], one["n_reads"]) and np.array_equal(many["n_snps"], one["n_snps"])
    with pytest.raises(Exception):
        multi3.fmx_run(s.plp, o, np.full(300, 9, np.int32))  # cluster id >= n_clusters


def test_cli_hosts_take_gpus(ctx, tmp_path):
    """`popscle demuxlet --gpus` / PSCL_GPU_IDS through the C++ host and the Python mirror: same files as one GPU."""
    import subprocess
    from popscle_b200 import _build, cli
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "demux_gt")
    exe = _build.build_host()
    outs = {}
    for tag, env in (("one", {}), ("two", {"PSCL_GPU_IDS": ",".join(map(str, _ids(2)))})):
        o = str(tmp_path / tag)
        subprocess.check_call([exe, "demuxlet", "--plp", "p", "--vcf", "g.vcf.gz", "--field", "GT", "--out", o], cwd=gold,
                              env=dict(os.environ, **env), stderr=subprocess.DEVNULL)
        outs[tag] = open(o + ".best").read()
    assert outs["one"] == outs["two"]
    os.environ["PSCL_GPU_IDS"] = ",".join(map(str, _ids(2)))
    cwd = os.getcwd()
    os.chdir(gold)
    try:
        cli.demuxlet(["--plp", "p", "--vcf", "g.vcf.gz", "--field", "GT", "--out", str(tmp_path / "py")])
    finally:
        os.chdir(cwd)
        del os.environ["PSCL_GPU_IDS"]
    assert open(str(tmp_path / "py") + ".best").read() == outs["one"]
    gold = os.path.join(os.path.dirname(gold), "fmx_default")
    import gzip
    res = {}
    for tag, env in (("one", {}), ("two", {"PSCL_GPU_IDS": ",".join(map(str, _ids(2)))})):
        o = str(tmp_path / ("f" + tag))
        subprocess.check_call([exe, "freemuxlet", "--plp", "p", "--nsample", "4", "--out", o, "--seed", "1"], cwd=gold,
                              env=dict(os.environ, **env), stderr=subprocess.DEVNULL)
        res[tag] = gzip.open(o + ".clust1.samples.gz", "rt").read()
    a = [l.split("\t") for l in res["one"].splitlines()]
    b = [l.split("\t") for l in res["two"].splitlines()]
    assert len(a) == len(b) > 100 and all(x[:6] == y[:6] for x, y in zip(a, b))  # ids, counts, types, best guesses


def test_sharded_demuxlet_on_the_tiny_pileup_form(ctx, multi3):
    """ABI 6 arrays through pscl_multi_demux_run: the shards point into the caller's arrays at pair offsets that are not
    multiples of 4 (two-bit counts) or 1024 (the large-count index)."""
    s = synth.make_pileup(C=333, nv=8, V=90000, kbar=401, seed=64)
    nrd = np.diff(s.plp.pair_read_ptr)
    nrd[::11] = 6
    from popscle_b200 import Pileup
    prp = np.concatenate([[0], np.cumsum(nrd)]).astype(np.int64)
    N = int(prp[-1])
    rng = np.random.default_rng(1)
    p2 = Pileup(s.plp.n_cells, s.plp.n_snps, s.plp.cell_ptr, s.plp.pair_snp, prp, rng.choice([0, 1, 2], N).astype(np.uint8),
                rng.integers(13, 41, N).astype(np.uint8), s.plp.snp_af)
    gp = synth.gt_to_gp(s.geno)
    one = ctx.demux_run(p2, gp, None, DEFAULT)
    assert multi3.demux_run(p2, gp, None, DEFAULT, compact=4).tobytes() == one.tobytes()
    assert multi3.fmx_run(p2, ctx.fmx_opts(3, max_iter=1, early_stop=False), compact=4)[0]["n_reads"].tolist() == \
        ctx.fmx_run(p2, ctx.fmx_opts(3, max_iter=1, early_stop=False))[0]["n_reads"].tolist()
