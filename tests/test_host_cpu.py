"""Host logic on CPU: file round trips, loader semantics, CLI errors, partitioning, and the N>1
paths over gloo (world_size 2) with the oracle standing in for the device."""
import os
import sys

import numpy as np
import pytest

import oracle_py as orc
from popscle_b200 import cli, dist, plpio, synth
from tests.engines import OracleEngine
from tests.parity import assert_close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plp_round_trip_and_bq_window(tmp_path):
    s = synth.make_pileup(C=40, nv=3, V=300, kbar=80, seed=5, cap_bq=40)
    sites = plpio.default_sites(300, s.af, seed=5)
    bcs = plpio.write_plp(str(tmp_path / "p"), s.plp, sites)
    L = plpio.load_plp(str(tmp_path / "p"))  # library defaults minBQ=1 capBQ=60
    assert L.barcodes == bcs
    for f in ("cell_ptr", "pair_snp", "pair_read_ptr", "read_allele", "read_qual"):
        assert np.array_equal(getattr(L.plp, f), getattr(s.plp, f)), f
    assert_close(L.plp.snp_af, s.af, "af", rtol=1e-12)
    L2 = plpio.load_plp(str(tmp_path / "p"), min_bq=20, cap_bq=30)  # sc_drop_seq.cpp:361-369
    keep = s.plp.read_qual >= 20
    assert L2.plp.n_reads == int(keep.sum())
    assert L2.plp.read_qual.max() == 30 and L2.plp.read_qual.min() == 20
    assert L2.plp.n_pairs <= s.plp.n_pairs  # pairs whose reads are all filtered vanish


def test_cel_filters_and_header_errors(tmp_path):
    s = synth.make_pileup(C=30, nv=3, V=200, kbar=60, seed=6, cap_bq=40)
    sites = plpio.default_sites(200, s.af, seed=6)
    bcs = plpio.write_plp(str(tmp_path / "p"), s.plp, sites)
    npair = np.diff(s.plp.cell_ptr)
    thr = int(np.median(npair))
    L = plpio.load_plp(str(tmp_path / "p"), min_snp=thr)
    assert L.barcodes == [b for b, n in zip(bcs, npair) if n >= thr]
    L = plpio.load_plp(str(tmp_path / "p"), group_list=bcs[5:12])
    assert L.barcodes == bcs[5:12]
    import gzip
    with gzip.open(tmp_path / "bad.cel.gz", "wt") as f:
        f.write("#DROPLET_ID\tBARCODE\tNUM.READ\tNUM.UMI\tNUM.SNP\n")
    for ext in ("var.gz", "plp.gz"):
        os.link(tmp_path / f"p.{ext}", tmp_path / f"bad.{ext}")
    with pytest.raises(ValueError, match="malformed or outdated"):
        plpio.load_plp(str(tmp_path / "bad"))


def test_cli_usage_errors(tmp_path):
    with pytest.raises(cli.UsageError, match="Missing required"):
        cli.demuxlet(["--plp", "x"], engine=OracleEngine())
    with pytest.raises(cli.UsageError, match="Missing required"):
        cli.freemuxlet(["--plp", "x", "--out", "y"], engine=OracleEngine())
    with pytest.raises(cli.UsageError, match="more than once"):
        cli.demuxlet(["--plp", "x", "--plp", "y"], engine=OracleEngine())
    with pytest.raises(cli.UsageError, match="Cannot recognize"):
        cli.freemuxlet(["--nope"], engine=OracleEngine())
    assert cli.main(["demuxlet", "--sam", "a.bam", "--vcf", "v", "--out", "o"]) == 134
    assert cli.main(["frobnicate"]) == 2


def test_balanced_partitions():
    s = synth.make_pileup(C=500, nv=3, V=4000, kbar=300, seed=7)
    for n in (1, 2, 3, 8):
        r = dist.balanced_cell_ranges(s.plp.cell_ptr, n)
        assert r[0][0] == 0 and r[-1][1] == 500 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        pairs = [int(s.plp.cell_ptr[b] - s.plp.cell_ptr[a]) for a, b in r]
        assert max(pairs) - min(pairs) <= 2 * int(np.diff(s.plp.cell_ptr).max())
        v = dist.balanced_snp_ranges(s.plp.pair_snp, 4000, n)
        assert v[0][0] == 0 and v[-1][1] == 4000 and all(a[1] == b[0] for a, b in zip(v, v[1:]))
        cnt = np.bincount(s.plp.pair_snp, minlength=4000)
        ps = [int(cnt[a:b].sum()) for a, b in v]
        assert sum(ps) == s.plp.n_pairs and max(ps) - min(ps) <= 2 * int(cnt.max())
    tiny = dist.balanced_cell_ranges(np.array([0, 5], dtype=np.int64), 4)  # more ranks than cells
    assert sum(b - a for a, b in tiny) == 1


def test_slices_are_consistent():
    s = synth.make_pileup(C=60, nv=3, V=500, kbar=100, seed=8)
    parts = [s.plp.slice_cells(a, b) for a, b in dist.balanced_cell_ranges(s.plp.cell_ptr, 3)]
    assert sum(p.n_pairs for p in parts) == s.plp.n_pairs and sum(p.n_reads for p in parts) == s.plp.n_reads
    sh = [s.plp.slice_snps(a, b) for a, b in dist.balanced_snp_ranges(s.plp.pair_snp, 500, 3)]
    assert sum(p.n_pairs for p in sh) == s.plp.n_pairs and all(p.n_cells == 60 for p in sh)


# ---- world_size 2 over gloo -------------------------------------------------------------------
def _worker_demux(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import torch.distributed as td
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    s = synth.make_pileup(C=90, nv=4, V=600, kbar=120, seed=31)
    gp = synth.gt_to_gp(s.geno)

    def gather(rec):
        out = [None] * world
        td.all_gather_object(out, rec)
        return out
    got = dist.demux_sharded(OracleEngine(1), s.plp, gp, None, [0.0, 0.5], 0.5, rank, world, gather)
    if rank == 0:
        full = orc.demux(s.plp, gp, None, [0.0, 0.5])
        q.put(got.tobytes() == full.tobytes())
    td.destroy_process_group()


class _OracleStep:
    """step-level freemuxlet surface over the oracle's pieces (host-logic test only)"""
    def __init__(self):
        self.new_f64 = lambda n: np.zeros(n)
        self.new_i32 = lambda n: np.zeros(n, dtype=np.int32)

    def set_f64(self, buf, values):
        buf[:] = values

    def init(self, plp, opts):
        self.plp, self.o = plp, opts
        self.nS = opts.n_clusters
        self.pair_gl = np.array([orc.fmx_pair_pileup(plp.read_allele[a:b], plp.read_qual[a:b])[0]
                                 for a, b in zip(plp.pair_read_ptr[:-1], plp.pair_read_ptr[1:])]).reshape(-1, 9)
        self.pair_cell = np.repeat(np.arange(plp.n_cells), np.diff(plp.cell_ptr))

    def stage1(self, st):
        C = self.plp.n_cells
        st[:] = 0
        st[2 * C:3 * C] = np.diff(self.plp.cell_ptr)

    def seed(self, st, init, cl):
        cl[:] = init
        self.types = np.where(init >= 0, 0, -1)
        self.cells = np.zeros(self.plp.n_cells, dtype=orc.FMX_CELL_DTYPE)
        for f in ("best_j", "best_k"):
            self.cells[f] = -1

    def mstep(self, cl):
        if cl is not None:
            self.member = np.array(cl)
        V, nS = self.plp.n_snps, self.nS
        gl = np.ones((V, nS, 9))
        for p in range(self.plp.n_pairs):  # ascending cell id inside every SNP
            j = self.member[self.pair_cell[p]]
            if j >= 0:
                v = self.plp.pair_snp[p]
                gl[v, j], _, _ = orc.fmx_merge(gl[v, j], [0, 0, 0], 0.0, self.pair_gl[p], [0, 0, 0], 0.0)
        self.clust_gl = gl

    def estep(self, it, llk):
        llk[:] = orc.fmx_estep(self.plp, self.pair_gl, self.clust_gl, self.nS, self.o.geno_error).ravel()

    def classify(self, llk, cl):
        nS = self.nS
        L = llk.reshape(self.plp.n_cells, -1)
        diag = [j * (j + 1) // 2 + j for j in range(nS)]
        sng = L[:, diag]
        off = np.ones(L.shape[1], bool); off[diag] = False
        best = sng.argmax(1)
        srt = np.sort(sng, 1)
        is_dbl = L[:, off].max(1) > srt[:, -1] + 2
        is_sng = ~is_dbl & (srt[:, -1] > srt[:, -2] + 2)
        new_types = np.where(is_dbl, 1, np.where(is_sng, 0, 2))
        changed = int(((new_types != self.types) | (is_sng & (self.cells["best_j"] != best))).sum())
        self.types = new_types
        self.cells["best_j"] = self.cells["best_k"] = best
        self.member = np.where(is_sng, best, -1)
        cl[:] = self.member

        class R: pass
        r = R(); r.n_changed = changed
        return r

    def fetch(self):
        return self.types.copy(), self.member.copy()


def _worker_fmx(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import torch
    import torch.distributed as td
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    s = synth.make_pileup(C=50, nv=3, V=400, kbar=120, seed=32)
    v0, v1 = dist.balanced_snp_ranges(s.plp.pair_snp, 400, world)[rank]

    def allreduce(buf):
        t = torch.from_numpy(buf)
        td.all_reduce(t)
    o = orc.fmx_opts(3, max_iter=4)
    init = s.truth_d1.astype(np.int32)
    (types, member), res = dist.fmx_em_sharded(_OracleStep(), s.plp.slice_snps(v0, v1), o, init, allreduce, 50)
    # without --init-cluster: greedy seeding over the whole pileup on rank 0 only, clusters broadcast
    calls = []

    def seed_full():
        if rank != 0:
            return None
        calls.append(1)
        r = orc.fmx_run(s.plp, orc.fmx_opts(3, max_iter=0))["cells"]
        st = np.concatenate([r["llk0"], r["llk2"], r["n_snps"].astype(np.float64), r["n_reads"].astype(np.float64)])
        return st, r["init_clust"]

    def bcast(a):
        box = [a]
        td.broadcast_object_list(box, src=0)
        return box[0]
    (types2, member2), _ = dist.fmx_em_sharded(_OracleStep(), s.plp.slice_snps(v0, v1), o, None, allreduce, 50, seed_full=seed_full, bcast=bcast)
    assert len(calls) == (1 if rank == 0 else 0)
    if rank == 0:
        (t1, m1), r1 = dist.fmx_em_sharded(_OracleStep(), s.plp, o, init, lambda b: None, 50)
        ref = orc.fmx_run(s.plp, o, init)["cells"]
        ok = bool(np.array_equal(types, t1) and np.array_equal(member, m1) and np.array_equal(types, ref["type"])
                  and np.array_equal(member, np.where(ref["type"] == 0, ref["best_j"], -1)))
        ref2 = orc.fmx_run(s.plp, o)["cells"]  # greedy seeding + EM, unsharded
        ok = ok and bool(np.array_equal(types2, ref2["type"]) and np.array_equal(member2, np.where(ref2["type"] == 0, ref2["best_j"], -1)))
        try:
            dist.fmx_em_sharded(_OracleStep(), s.plp, o, None, lambda b: None, 50)
            ok = False
        except ValueError:
            pass
        q.put(ok)
    td.destroy_process_group()


def _spawn(fn, port):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=fn, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    ok = q.get(timeout=240)
    for p in ps:
        p.join(60)
    assert all(p.exitcode == 0 for p in ps)
    return ok


def test_gloo_barcode_sharded_demuxlet(built):
    assert _spawn(_worker_demux, 29611)


def test_gloo_snp_sharded_freemuxlet(built):
    assert _spawn(_worker_fmx, 29612)


def test_vcf_without_gt_and_bad_allele_numbers(tmp_path, built):
    """GT is only required when a site filter or --field GT reads it (bcf_filter_arg.h:110-113), and an
    allele number beyond the ALT list is an input error, not an out-of-bounds write — in both hosts."""
    import gzip
    import subprocess
    from popscle_b200 import _build, plpio
    s = synth.make_pileup(C=12, nv=2, V=40, kbar=20, seed=4)
    sites = plpio.default_sites(40, s.af, seed=4)
    plpio.write_plp(str(tmp_path / "p"), s.plp, sites)
    head = "##fileformat=VCFv4.2\n##contig=<ID=1>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\tB\n"

    def vcf(name, fmt, cell):
        with gzip.open(tmp_path / name, "wt") as f:
            f.write(head)
            for v in range(40):
                f.write(f"{sites.chrom[v]}\t{int(sites.pos[v])}\t.\t{sites.ref[v]}\t{sites.alt[v]}\t.\tPASS\t.\t{fmt}\t{cell(v, 0)}\t{cell(v, 1)}\n")
    vcf("gp_only.vcf.gz", "GP", lambda v, j: "0.8,0.15,0.05" if (v + j) % 2 else "0.1,0.2,0.7")
    vcf("bad_gt.vcf.gz", "GT", lambda v, j: "0/3" if v == 7 else "0/1")
    exe = _build.build_host()
    # no GT column: fine once nothing needs it ...
    L = plpio.load_plp(str(tmp_path / "p"), str(tmp_path / "gp_only.vcf.gz"), field="GP", min_mac=0, min_callrate=0.0)
    assert L.geno.has_gp.sum() == 40 and L.geno.gp.shape == (40, 2, 3)
    r = subprocess.run([exe, "demuxlet", "--plp", "p", "--vcf", "gp_only.vcf.gz", "--field", "GP", "--min-mac", "0", "--min-callrate", "0",
                        "--out", "o", "--dry-run"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0 and '"has_gp": 40' in r.stdout, r.stderr
    # ... and the reference's error while the default filters are on
    with pytest.raises(ValueError, match="Cannot find the field GT"):
        plpio.load_plp(str(tmp_path / "p"), str(tmp_path / "gp_only.vcf.gz"), field="GP")
    r = subprocess.run([exe, "demuxlet", "--plp", "p", "--vcf", "gp_only.vcf.gz", "--field", "GP", "--out", "o", "--dry-run"], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode == 134 and "Cannot find the field GT" in r.stderr
    with pytest.raises(ValueError, match="beyond the ALT list"):
        plpio.load_plp(str(tmp_path / "p"), str(tmp_path / "bad_gt.vcf.gz"), field="GT")
    r = subprocess.run([exe, "demuxlet", "--plp", "p", "--vcf", "bad_gt.vcf.gz", "--field", "GT", "--out", "o", "--dry-run"], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode == 134 and "beyond the ALT list" in r.stderr


def test_parallel_ingest_bgzf_gzip_and_thread_counts_agree(tmp_path, built):
    """N1: the C++ loader inflates BGZF blocks and parses the PLP table on several threads.  The flat image and every compact
    form must not depend on the container (BGZF as dsc-pileup writes it / one gzip stream / plain text), on the thread count,
    or on the row order of the table (SNP-major, cell-major, shuffled), and a corrupt block is an error, not garbage."""
    import gzip
    import json
    import subprocess
    from popscle_b200 import _build, plpio
    s = synth.make_pileup(C=150, nv=3, V=900, kbar=260, seed=11)
    sites = plpio.default_sites(900, s.af, seed=11)
    exe = _build.build_host()
    for sub in ("bg", "gz", "txt", "cellmajor", "shuffled"):
        (tmp_path / sub).mkdir()
    plpio.write_plp(str(tmp_path / "bg" / "p"), s.plp, sites, bgzf=True)
    plpio.write_plp(str(tmp_path / "gz" / "p"), s.plp, sites)
    plpio.write_vcf(str(tmp_path / "ref.vcf.gz"), sites, ["A", "B", "C"], geno=s.geno)
    text = plpio.plp_text(s.plp)
    assert gzip.open(tmp_path / "bg" / "p.plp.gz").read() == text and len(text) > 3 * 0xff00  # several BGZF blocks
    rows = text.decode().splitlines()
    head, body = rows[0], rows[1:]
    rng = np.random.default_rng(5)
    variants = {"txt": body, "cellmajor": sorted(body, key=lambda r: (int(r.split("\t")[0]), int(r.split("\t")[1]))),
                "shuffled": [body[i] for i in rng.permutation(len(body))]}
    for sub, lines in variants.items():
        for ext in ("cel", "var"):
            (tmp_path / sub / f"p.{ext}.gz").write_bytes((tmp_path / "gz" / f"p.{ext}.gz").read_bytes())
        (tmp_path / sub / "p.plp.gz").write_bytes(("\n".join([head] + lines) + "\n").encode())  # plain text under the .gz name

    def run(sub, threads):
        r = subprocess.run([exe, "demuxlet", "--plp", str(tmp_path / sub / "p"), "--vcf", str(tmp_path / "ref.vcf.gz"), "--field", "GT", "--out", "o",
                            "--dry-run"], cwd=tmp_path, capture_output=True, text=True, env=dict(os.environ, PSCL_LOADER_THREADS=str(threads)))
        assert r.returncode == 0, r.stderr
        return json.loads(r.stdout.strip().splitlines()[-1])
    want = run("gz", 1)
    assert want["pairs"] == s.plp.n_pairs and want["tiny_form"] == 1 and want["compact_form"] == 3
    for sub in ("bg", "gz", "txt", "cellmajor", "shuffled"):
        for threads in (1, 3, 16):
            assert run(sub, threads) == want, (sub, threads)
    # a flipped byte inside a BGZF block's deflate stream
    raw = bytearray((tmp_path / "bg" / "p.plp.gz").read_bytes())
    raw[len(raw) // 2] ^= 0x5a
    (tmp_path / "bg" / "p.plp.gz").write_bytes(bytes(raw))
    r = subprocess.run([exe, "demuxlet", "--plp", str(tmp_path / "bg" / "p"), "--vcf", str(tmp_path / "ref.vcf.gz"), "--field", "GT", "--out", "o",
                        "--dry-run"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 134 and ("Corrupt" in r.stderr or "not in" in r.stderr or "Cannot access" in r.stderr), r.stderr


def test_vectorised_plp_text_equals_row_formatting():
    """plpio.plp_text assembles the .plp.gz table with array operations; it must be the text per-row formatting gives
    (multi-digit ids, pairs of 1..8 base-calls, empty cells, SNP-major order with droplet ids ascending inside a SNP)."""
    s = synth.make_pileup(C=230, nv=3, V=12000, kbar=150, seed=77)
    plp = s.plp
    C = plp.n_cells
    pair_cell = np.repeat(np.arange(C), np.diff(plp.cell_ptr))
    order = np.lexsort((pair_cell, plp.pair_snp))
    al = (plp.read_allele + ord("0")).astype(np.uint8).tobytes()
    bq = (plp.read_qual + 33).astype(np.uint8).tobytes()
    prp = plp.pair_read_ptr
    rows = [plpio.PLP_HEADER]
    for p in order:
        a, b = int(prp[p]), int(prp[p + 1])
        rows.append(f"{int(pair_cell[p])}\t{int(plp.pair_snp[p])}\t{al[a:b].decode()}\t{bq[a:b].decode()}")
    assert plpio.plp_text(plp) == ("\n".join(rows) + "\n").encode()
