"""Golden vectors written by the reference itself (oracle/_ref/popscle_ref = the reference's own
demuxlet / freemuxlet translation units, see tests/golden/make_golden.py): the same argv is replayed
through popscle_b200.cli with the CPU oracle (must reproduce the files byte for byte) and with the
CUDA library (ids exact, LLK columns within 1e-4 relative — BASELINE.json's tolerance)."""
import gzip
import json
import os
import shutil

import pytest

from popscle_b200 import cli
from tests.engines import OracleEngine

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(d for d in os.listdir(GOLD) if os.path.isdir(os.path.join(GOLD, d)))
LLK_RTOL = 1e-4       # north_star tolerance on LLK columns
PRINT_ATOL = 0.0101   # two %.2lf roundings


def _replay(case, tmp_path, engine):
    src = os.path.join(GOLD, case)
    work = tmp_path / case
    shutil.copytree(src, work)
    argv = json.load(open(work / "cmd.json"))["argv"]
    argv = [("out" if a == "ref" and argv[i - 1] == "--out" else a) for i, a in enumerate(argv)]
    cwd = os.getcwd()
    os.chdir(work)
    try:
        rc = cli.COMMANDS[argv[0]](argv[1:], engine=engine)
    finally:
        os.chdir(cwd)
    assert rc == 0
    pairs = []
    for fn in sorted(os.listdir(work)):
        if fn.startswith("ref."):
            ours = work / ("out." + fn[4:])
            if not ours.exists():
                ours = work / ("out." + fn[4:] + ".gz")
            assert ours.exists(), f"{case}: output {fn[4:]} was not written"
            txt = gzip.open(ours, "rt").read() if str(ours).endswith(".gz") else open(ours).read()
            txt = "".join(l for l in txt.splitlines(True) if not l.startswith("##fileDate="))
            pairs.append((fn, open(work / fn).read(), txt))
    assert pairs
    return pairs


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference_files(case, tmp_path, built):
    for fn, ref, ours in _replay(case, tmp_path, OracleEngine()):
        if ref != ours:
            a, b = ref.split("\n"), ours.split("\n")
            bad = [i for i, (x, y) in enumerate(zip(a, b)) if x != y]
            raise AssertionError(f"{case}/{fn}: {len(bad)} of {len(a)} lines differ (lengths {len(a)}/{len(b)}); first:\nREF {a[bad[0]][:300]}\nOUR {b[bad[0]][:300]}")


def _num(s):
    try:
        return float(s)
    except ValueError:
        return None


def _close(a, b, rtol, atol):
    if a == b or (a != a and b != b):
        return True
    return abs(a - b) <= rtol * max(1.0, abs(b)) + atol


def _cmp_guess(x, y):
    """'j,k,alpha' (demuxlet) or 'j,k' (freemuxlet): unordered pair at alpha 0.5 / for cluster pairs"""
    fx, fy = x.split(","), y.split(",")
    if len(fx) != len(fy):
        return False
    if len(fx) == 3:
        if fx[2] != fy[2]:
            return False
        return (fx[:2] == fy[:2]) or (fx[2] == "0.50" and sorted(fx[:2]) == sorted(fy[:2]))
    return fx == fy


def _compare_table(ref, ours, what):
    a, b = ref.rstrip("\n").split("\n"), ours.rstrip("\n").split("\n")
    assert len(a) == len(b), f"{what}: {len(a)} vs {len(b)} lines"
    hdr = a[0].split("\t")
    assert a[0] == b[0]
    for ln, (x, y) in enumerate(zip(a[1:], b[1:]), 2):
        fx, fy = x.split("\t"), y.split("\t")
        assert len(fx) == len(fy), f"{what}:{ln}"
        for h, u, v in zip(hdr, fx, fy):
            if h.endswith(".GUESS") and "," in u:
                assert _cmp_guess(u, v), f"{what}:{ln} {h}: {u} vs {v}"
                continue
            nu, nv = _num(u), _num(v)
            if nu is None or nv is None or h in ("INT_ID", "NUM.SNPS", "NUM.READS", "NSNPs", "NREADs", "SNG.BEST.GUESS", "SNG.NEXT.GUESS"):
                assert u == v, f"{what}:{ln} {h}: {u} vs {v}"
            elif "POSTERIOR" in h:
                # %.2lg / %.5lf of probabilities (or of a log, cmd_cram_demuxlet.cpp:949): 2 significant digits
                assert _close(nu, nv, 0.06, 1e-5), f"{what}:{ln} {h}: {u} vs {v}"
            else:
                assert _close(nu, nv, LLK_RTOL, PRINT_ATOL), f"{what}:{ln} {h}: {u} vs {v}"


def _compare_vcf(ref, ours, what):
    a, b = ref.rstrip("\n").split("\n"), ours.rstrip("\n").split("\n")
    assert len(a) == len(b), f"{what}: {len(a)} vs {len(b)} lines"
    for ln, (x, y) in enumerate(zip(a, b), 1):
        if x == y:
            continue
        assert not x.startswith("#"), f"{what}:{ln} header differs"
        fx, fy = x.split("\t"), y.split("\t")
        assert fx[:9] == fy[:9], f"{what}:{ln} site columns differ"
        for u, v in zip(fx[9:], fy[9:]):
            pu, pv = u.split(":"), v.split(":")
            assert pu[0] == pv[0] and pu[2] == pv[2] and pu[3] == pv[3], f"{what}:{ln} GT/DP/AD: {u} vs {v}"
            assert abs(int(pu[1]) - int(pv[1])) <= 1, f"{what}:{ln} GQ: {u} vs {v}"   # int() truncation of -10 log10
            for s, t in zip(pu[4].split(","), pv[4].split(",")):
                assert abs(int(s) - int(t)) <= 1, f"{what}:{ln} PL: {u} vs {v}"
            for s, t in zip(pu[5].split(","), pv[5].split(",")):
                assert _close(float(s), float(t), 1e-2, 1e-12), f"{what}:{ln} GP: {u} vs {v}"


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_matches_reference_files(case, tmp_path, ctx):
    for fn, ref, ours in _replay(case, tmp_path, ctx):
        if fn.endswith(".vcf"):
            _compare_vcf(ref, ours, f"{case}/{fn}")
        else:
            _compare_table(ref, ours, f"{case}/{fn}")


# ---- the C++ CLI host (popscle_b200/host -> popscle_b200/popscle) -----------------------------------
def _fnv1a(*arrays):
    h = 1469598103934665603
    for a in arrays:
        for b in a.tobytes():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def _host_argv(case, work):
    argv = json.load(open(os.path.join(GOLD, case, "cmd.json")))["argv"]
    return [("out" if a == "ref" and argv[i - 1] == "--out" else a) for i, a in enumerate(argv)]


@pytest.fixture(scope="session")
def host_exe(built):
    from popscle_b200 import _build
    exe = _build.build_host()
    assert exe and os.path.exists(exe)
    return exe


@pytest.mark.parametrize("case", CASES)
def test_cpp_host_loader_matches_python_loader(case, tmp_path, host_exe):
    """`--dry-run` (test hook): the C++ loader's flat image has the same checksum as plpio.load_plp's."""
    import subprocess
    import numpy as np
    from popscle_b200 import plpio
    work = tmp_path / case
    shutil.copytree(os.path.join(GOLD, case), work)
    argv = _host_argv(case, work)
    r = subprocess.run([host_exe] + argv + ["--dry-run"], cwd=work, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    o = cli._parse(argv[1:], {"demuxlet": cli.DEMUXLET_SPEC, "freemuxlet": cli.FREEMUXLET_SPEC, "freemuxlet-old": cli.FREEMUXLET_OLD_SPEC}[argv[0]])
    cwd = os.getcwd()
    os.chdir(work)
    try:
        if argv[0] == "demuxlet":
            L = plpio.load_plp(o["plp"], o["vcf"], field=o["field"], geno_error_offset=o["geno-error-offset"],
                               geno_error_coeff=o["geno-error-coeff"], r2_info=o["r2-info"], sm_list=o["sm"] or None,
                               min_bq=o["min-BQ"], cap_bq=o["cap-BQ"], min_read=o["min-total"], min_umi=o["min-umi"], min_snp=o["min-snp"],
                               group_list=cli._read_list(o["group-list"]) if o["group-list"] else None)
        elif argv[0] == "freemuxlet":
            L = plpio.load_plp(o["plp"], None, min_bq=o["min-BQ"], cap_bq=o["cap-BQ"])
        else:
            L = plpio.load_plp(o["plp"], None)
    finally:
        os.chdir(cwd)
    p = L.plp
    assert (got["cells"], got["snps"], got["pairs"], got["reads"]) == (p.n_cells, p.n_snps, p.n_pairs, p.n_reads)
    assert got["pileup_fnv1a"] == _fnv1a(p.cell_ptr, p.pair_snp, p.pair_read_ptr, p.read_allele, p.read_qual, p.snp_af)
    if L.geno is not None:
        assert got["samples"] == len(L.geno.samples) and got["has_gp"] == int(L.geno.has_gp.sum())
        assert got["geno_fnv1a"] == _fnv1a(np.ascontiguousarray(L.geno.gp), L.geno.has_gp)
        # ABI 4 raw genotype forms, as geno_view() hands them to the library
        assert got["gt8"] == (1 if L.geno.gt8 is not None else 0) == (1 if o["field"] == "GT" else 0)
        assert got["raw_geno_fnv1a"] == _fnv1a(L.geno.gt8 if L.geno.gt8 is not None else L.geno.gp_f32, L.geno.err_snp)
    # ABI 2/3 compact pileup arrays, as view() hands them to the library
    p32, aq = p.compact()
    c3 = p.compact3()
    assert got["compact_form"] == (3 if c3 is not None else 2)
    assert got["compact_fnv1a"] == _fnv1a(p32, aq, *(c3 if c3 is not None else ()))
    # ABI 6 forms: 8-bit gaps / 2-bit counts with the large values on the side, palette-indexed base-calls
    c4, pr = p.compact4(), p.packed_reads()
    assert got["tiny_form"] == (1 if c4 is not None else 0) and got["read_bits"] == (pr[2] if pr is not None and c4 is not None else 0)
    if c4 is not None:
        assert got["tiny_fnv1a"] == _fnv1a(*c4, *((pr[0][:(p.n_reads * pr[2] + 7) // 8 + 1], pr[1]) if pr is not None else ()))


def test_cpp_host_errors(host_exe, tmp_path):
    import subprocess
    r = subprocess.run([host_exe, "demuxlet", "--plp", "x"], capture_output=True, text=True)
    assert r.returncode == 134 and "FATAL ERROR" in r.stderr and "Missing required option(s)" in r.stderr
    r = subprocess.run([host_exe, "freemuxlet", "--plp", "x", "--out", "y"], capture_output=True, text=True)
    assert r.returncode == 134 and "--nsample" in r.stderr
    r = subprocess.run([host_exe, "demuxlet", "--nope", "1"], capture_output=True, text=True)
    assert r.returncode == 134 and "Cannot recognize" in r.stderr
    assert subprocess.run([host_exe, "dsc-pileup"], capture_output=True).returncode == 2


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cpp_host_matches_reference_files(case, tmp_path, host_exe):
    """the drop-in binary end to end: files in, CUDA likelihoods, reference-format files out"""
    import subprocess
    work = tmp_path / case
    shutil.copytree(os.path.join(GOLD, case), work)
    r = subprocess.run([host_exe] + _host_argv(case, work), cwd=work, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    n = 0
    for fn in sorted(os.listdir(work)):
        if not fn.startswith("ref."):
            continue
        ours = work / ("out." + fn[4:])
        if not ours.exists():
            ours = work / ("out." + fn[4:] + ".gz")
        assert ours.exists(), fn
        txt = gzip.open(ours, "rt").read() if str(ours).endswith(".gz") else open(ours).read()
        txt = "".join(l for l in txt.splitlines(True) if not l.startswith("##fileDate="))
        ref = open(work / fn).read()
        (_compare_vcf if fn.endswith(".vcf") else _compare_table)(ref, txt, f"{case}/{fn} (C++ host)")
        n += 1
    assert n > 0


def test_cpp_plp_parser_line_endings(tmp_path, host_exe):
    """The in-place PLP row parser of the C++ host: a last row without a newline, CRLF line ends and rows that straddle
    the reader's 1 MiB buffer give the same image; an empty line ends the file as tsv_reader does (tsv_reader.cpp:37-41)."""
    import gzip
    import subprocess
    case = "demux_gt"
    argv = _host_argv(case, tmp_path)

    def run(mutate):
        work = tmp_path / mutate.__name__
        shutil.copytree(os.path.join(GOLD, case), work)
        plp = work / "p.plp.gz"
        text = gzip.open(plp, "rt").read()
        with gzip.open(plp, "wt", newline="") as f:
            f.write(mutate(text))
        r = subprocess.run([host_exe] + argv + ["--dry-run"], cwd=work, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        return json.loads(r.stdout)

    def same(t): return t
    def no_final_newline(t): return t.rstrip("\n")
    def crlf(t): return t.replace("\n", "\r\n")
    def padded(t):  # > 1 MiB of leading whitespace on one row: the row straddles (and outgrows) the read buffer
        lines = t.split("\n")
        lines[5] = " " * (3 << 20) + lines[5]
        return "\n".join(lines)
    def cut_at_empty_line(t):
        lines = t.split("\n")
        return "\n".join(lines[:40] + [""] + lines[40:])

    def cell_major(t):  # already in the image's order: no reordering at all
        lines = t.rstrip("\n").split("\n")
        body = sorted(lines[1:], key=lambda r: (int(r.split()[0]), int(r.split()[1])))
        return "\n".join(lines[:1] + body) + "\n"
    def reversed_rows(t):  # neither order: the generic stable sort
        lines = t.rstrip("\n").split("\n")
        return "\n".join(lines[:1] + lines[1:][::-1]) + "\n"

    base = run(same)
    for m in (no_final_newline, crlf, padded, cell_major, reversed_rows):
        got = run(m)
        assert got["pileup_fnv1a"] == base["pileup_fnv1a"] and got["pairs"] == base["pairs"], m.__name__
    short = run(cut_at_empty_line)
    assert 0 < short["reads"] < base["reads"]
    # a (cell, SNP) listed on two rows is one pair with the reads of both (std::map semantics); the scatter path (SNP-major
    # file) and the sorted path (cell-major file) agree on it
    def dup_snp_major(t):
        lines = t.rstrip("\n").split("\n")
        return "\n".join(lines[:8] + [lines[7]] + lines[8:]) + "\n"
    def dup_cell_major(t): return cell_major(dup_snp_major(t))
    d1, d2 = run(dup_snp_major), run(dup_cell_major)
    assert d1["pairs"] == base["pairs"] and d1["reads"] > base["reads"]
    assert d1["pileup_fnv1a"] == d2["pileup_fnv1a"] != base["pileup_fnv1a"]
