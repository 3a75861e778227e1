"""The C-ABI library builds without a GPU, exports every symbol include/*.h declares, and refuses
to run without a device (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from popscle_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "popscle_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pscl_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported(built):
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(capi.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/popscle_b200.h but not exported: {missing}"
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (pscl_[a-z0-9_]+)", out))
    assert set(names) <= exported
    assert exported <= set(names), f"exported but not declared in the header: {sorted(exported - set(names))}"
    assert sorted(capi.EXPORTED_SYMBOLS) == names


def test_library_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "--list-elf", capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "popscle_b200.h"\n#include <stddef.h>\nint main(void){ pscl_demux_cell c; pscl_fmx_cell f; return sizeof(c)==160 && sizeof(f)==160 && sizeof(pscl_pileup)==208 && offsetof(pscl_pileup, cell_read_ptr)==200 && offsetof(pscl_pileup, read_packed)==176 && offsetof(pscl_pileup, read_bits)==192 && offsetof(pscl_pileup, read_aq)==80 && offsetof(pscl_pileup, pair_nreads8)==104 && offsetof(pscl_pileup, pair_snp_delta8)==112 && offsetof(pscl_pileup, nreads_big_ptr)==152 && offsetof(pscl_pileup, n_nreads_big)==168 && sizeof(pscl_geno)==56 && offsetof(pscl_geno, geno_err)==48 && sizeof(pscl_fmx_opts)==80 && offsetof(pscl_fmx_opts, seed)==56 && offsetof(pscl_fmx_opts, bf_thres)==64 && offsetof(pscl_fmx_opts, keep_init_missing)==76 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    assert subprocess.call([str(exe)]) == 0
    assert capi.DEMUX_CELL_DTYPE.itemsize == 160 and capi.FMX_CELL_DTYPE.itemsize == 160


def test_no_cpu_fallback(built):
    """Without a GPU pscl_create must fail with PSCL_ENODEV and the Python host must raise."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = capi.load_library()
    assert lib.pscl_abi_version() == 7
    h = ctypes.c_void_p()
    err = ctypes.create_string_buffer(256)
    rc = lib.pscl_create(0, ctypes.byref(h), err, len(err))
    assert rc == -2 and not h.value and b"no CPU fallback" in err.value
    with pytest.raises(capi.PsclError):
        capi.Context(0)
    from popscle_b200 import cli
    with pytest.raises(capi.PsclError):
        gold = os.path.join(ROOT, "tests", "golden", "demux_gt")
        cwd = os.getcwd()
        os.chdir(gold)
        try:
            cli.demuxlet(["--plp", "p", "--vcf", "g.vcf.gz", "--field", "GT", "--out", "/tmp/never_written"])
        finally:
            os.chdir(cwd)


def test_product_never_imports_the_oracle():
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "popscle_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".cpp", ".h")):
                t = open(os.path.join(dp, f)).read()
                if re.search(r"oracle_py|popscle_oracle|liboracle|oracle/_ref", t):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_pileup_struct_matches_the_ctypes_mirror():
    """ABI 2/3: the compact arrays sit behind the wide ones; the ctypes mirror has the same layout."""
    import ctypes as C
    from popscle_b200 import capi
    assert C.sizeof(capi.CPileup) == 208 and capi.CPileup.cell_read_ptr.offset == 200 and capi.CPileup.read_packed.offset == 176 and capi.CPileup.read_bits.offset == 192 and capi.CPileup.pair_snp_delta8.offset == 112 and capi.CPileup.nreads_big_ptr.offset == 152 and capi.CPileup.n_nreads_big.offset == 168
    assert capi.CPileup.cell_first_snp.offset == 88 and capi.CPileup.pair_snp_delta16.offset == 96 and capi.CPileup.pair_nreads8.offset == 104
    assert capi.CPileup.pair_read_ptr32.offset == 72 and capi.CPileup.read_aq.offset == 80
    assert C.sizeof(capi.CFmxOpts) == 80 and capi.CFmxOpts.bf_thres.offset == 64 and capi.CFmxOpts.keep_init_missing.offset == 76 and capi.CFmxOpts.randomize_singlet_score.offset == 52 and capi.CFmxOpts.seed.offset == 56
    assert C.sizeof(capi.CGeno) == 56 and capi.CGeno.gp_f32.offset == 24 and capi.CGeno.geno_err.offset == 48
    from popscle_b200 import synth
    s = synth.make_pileup(C=5, nv=2, V=50, kbar=50, seed=3)
    p32, aq = s.plp.compact()
    assert p32.dtype == np.uint32 and (p32 == s.plp.pair_read_ptr).all()
    assert ((aq >> 6) == s.plp.read_allele).all() and ((aq & 63) == s.plp.read_qual).all()
    cs = s.plp.c_struct(compact=True)
    assert cs.pair_read_ptr is None and cs.read_allele is None and cs.read_aq == aq.ctypes.data
    # ABI 3: delta-coded pair arrays decode back to pair_snp / pair_read_ptr
    first, d16, n8 = s.plp.compact3()
    assert d16.dtype == np.uint16 and n8.dtype == np.uint8 and first.dtype == np.int32
    snp = np.empty(s.plp.n_pairs, dtype=np.int64)
    for c in range(s.plp.n_cells):
        b, e = s.plp.cell_ptr[c], s.plp.cell_ptr[c + 1]
        snp[b:e] = first[c] + np.cumsum(d16[b:e].astype(np.int64))
    assert (snp == s.plp.pair_snp).all()
    assert (np.concatenate([[0], np.cumsum(n8.astype(np.int64))]) == s.plp.pair_read_ptr).all()
    cs = s.plp.c_struct(compact=3)
    assert cs.pair_snp is None and cs.pair_read_ptr32 is None and cs.pair_nreads8 == n8.ctypes.data
    # ABI 6: 8-bit gaps / 2-bit counts with the large values on the side decode back to the same arrays
    big = synth.make_pileup(C=40, nv=2, V=60000, kbar=120, seed=6).plp  # mean gap 500: many markers
    nrd = np.diff(big.pair_read_ptr); nrd[::7] = 9; nrd[5] = 255
    big.pair_read_ptr = np.concatenate([[0], np.cumsum(nrd)]).astype(np.int64)
    big.read_allele = np.zeros(int(big.pair_read_ptr[-1]), np.uint8); big.read_qual = np.full(int(big.pair_read_ptr[-1]), 30, np.uint8)
    first4, d8, gbig, cbp, n2, nbig, nbp = big.compact4()
    assert d8.dtype == np.uint8 and gbig.dtype == np.uint32 and (d8 == 255).sum() == len(gbig) > 100 and cbp[-1] == len(gbig)
    gap = d8.astype(np.int64); gap[d8 == 255] = gbig
    for c in range(big.n_cells):
        b, e = big.cell_ptr[c], big.cell_ptr[c + 1]
        assert (first4[c] + np.cumsum(gap[b:e]) == big.pair_snp[b:e]).all() and cbp[c] == (d8[:b] == 255).sum()
    f = np.stack([(n2 >> k) & 3 for k in (0, 2, 4, 6)], 1).ravel()[:big.n_pairs].astype(np.int64)
    assert (f == 0).sum() == len(nbig) and nbp[-1] == len(nbig) and all(nbp[k] == (f[:1024 * k] == 0).sum() for k in range(len(nbp) - 1))
    f[f == 0] = nbig
    assert (f == nrd).all()
    cs = big.c_struct(compact=4)
    assert cs.pair_snp is None and cs.pair_snp_delta16 is None and cs.pair_nreads8 is None and cs.pair_snp_delta8 == d8.ctypes.data
    assert cs.n_gap_big == len(gbig) and cs.n_nreads_big == len(nbig) and cs.nreads_big_ptr == nbp.ctypes.data
    wide = synth.make_pileup(C=3, nv=2, V=50, kbar=10, seed=4).plp
    wide.n_snps = 200000
    wide.pair_snp = wide.pair_snp.copy(); wide.pair_snp[wide.cell_ptr[1] - 1] = 199999  # a gap >= 65536: no delta form
    assert wide.compact3() is None
    cs = wide.c_struct(compact=3)
    assert cs.pair_snp is not None and cs.pair_snp_delta16 is None


def test_raw_genotype_forms_build_the_right_struct():
    """ABI 4: RawGeno fills gt8 / gp_f32 / geno_err(_snp) and leaves gp NULL; the plain table fills gp only."""
    from popscle_b200 import RawGeno
    from popscle_b200.capi import Context
    gt8 = np.zeros((7, 3), dtype=np.uint8)
    g, keep, V, nv = Context._geno(RawGeno(gt8=gt8, err=0.05), None)
    assert (V, nv) == (7, 3) and g.gp is None and g.gp_f32 is None and g.gt8 == keep[0].ctypes.data and g.geno_err == 0.05
    f32 = np.zeros((7, 3, 3), dtype=np.float32)
    es = np.linspace(0, 1, 7)
    g, keep, V, nv = Context._geno(RawGeno(gp_f32=f32, err_snp=es), np.ones(7, np.uint8))
    assert g.gt8 is None and g.gp_f32 == keep[0].ctypes.data and g.geno_err_snp == keep[1].ctypes.data and g.has_gp == keep[2].ctypes.data
    g, keep, V, nv = Context._geno(np.zeros((7, 3, 3)), None)
    assert g.gp == keep[0].ctypes.data and g.gt8 is None and g.gp_f32 is None and g.geno_err_snp is None


def test_every_environment_switch_is_documented():
    """each getenv("PSCL_...") of the library and the C++ host is named in INTEGRATION.md, the header or DESIGN.md"""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    for f in glob.glob(os.path.join(root, "popscle_b200", "csrc", "*")) + glob.glob(os.path.join(root, "popscle_b200", "host", "*")):
        if os.path.isfile(f):
            names |= set(re.findall(r'getenv\("(PSCL_[A-Z0-9_]+)"\)', open(f, errors="ignore").read()))
    assert len(names) >= 15
    docs = "".join(open(os.path.join(root, f)).read() for f in ("INTEGRATION.md", "DESIGN.md", os.path.join("include", "popscle_b200.h")))
    missing = sorted(n for n in names if n not in docs)
    assert not missing, missing
