"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never by anything under popscle_b200/.  Builds the library on first use (gcc, a few seconds).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")

DEMUX_CELL_DTYPE = np.dtype([
    ("n_snps", "<i4"), ("type", "<i4"),
    ("best_j", "<i4"), ("best_k", "<i4"), ("best_a", "<i4"),
    ("next_j", "<i4"), ("next_k", "<i4"), ("next_a", "<i4"),
    ("sng_best", "<i4"), ("sng_next", "<i4"),
    ("dbl_best_j", "<i4"), ("dbl_best_k", "<i4"), ("dbl_best_a", "<i4"),
    ("dbl_next_j", "<i4"), ("dbl_next_k", "<i4"), ("dbl_next_a", "<i4"),
    ("best_llk", "<f8"), ("next_llk", "<f8"), ("best_pp", "<f8"), ("sng_pp", "<f8"),
    ("sng_best_llk", "<f8"), ("sng_next_llk", "<f8"), ("sng_only_pp", "<f8"),
    ("dbl_best_llk", "<f8"), ("dbl_next_llk", "<f8"), ("sum_llk", "<f8"), ("sng_llk", "<f8"),
    ("reserved_", "<f8"),
])
FMX_CELL_DTYPE = np.dtype([
    ("n_snps", "<i4"), ("n_reads", "<i4"), ("type", "<i4"), ("clust", "<i4"),
    ("best_j", "<i4"), ("best_k", "<i4"), ("next_j", "<i4"), ("next_k", "<i4"),
    ("sng_best", "<i4"), ("sng_next", "<i4"),
    ("dbl_best_j", "<i4"), ("dbl_best_k", "<i4"), ("dbl_next_j", "<i4"), ("dbl_next_k", "<i4"),
    ("init_clust", "<i4"), ("reserved_", "<i4"),
    ("best_llk", "<f8"), ("next_llk", "<f8"), ("best_pp", "<f8"), ("sng_pp", "<f8"), ("sng_only_pp", "<f8"),
    ("sng_best_llk", "<f8"), ("sng_next_llk", "<f8"), ("dbl_best_llk", "<f8"), ("dbl_next_llk", "<f8"),
    ("sum_llk", "<f8"), ("llk0", "<f8"), ("llk2", "<f8"),
])


class OPileup(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("n_snps", C.c_int32), ("n_pairs", C.c_int64), ("n_reads", C.c_int64),
                ("cell_ptr", C.c_void_p), ("pair_snp", C.c_void_p), ("pair_read_ptr", C.c_void_p),
                ("read_allele", C.c_void_p), ("read_qual", C.c_void_p), ("snp_af", C.c_void_p)]


class OFmxOpts(C.Structure):
    _fields_ = [("n_clusters", C.c_int32), ("doublet_prior", C.c_double), ("geno_error", C.c_double),
                ("max_iter", C.c_int32), ("early_stop", C.c_int32), ("frac_init_clust", C.c_double),
                ("singlet_score_thres", C.c_double), ("mode_old", C.c_int32),
                ("randomize_singlet_score", C.c_int32), ("seed", C.c_int32),
                ("bf_thres", C.c_double), ("iter_init", C.c_int32), ("keep_init_missing", C.c_int32)]


class OFmxResult(C.Structure):
    _fields_ = [("n_iter", C.c_int32), ("n_changed", C.c_int32), ("n_singlet", C.c_int32),
                ("n_doublet", C.c_int32), ("n_ambiguous", C.c_int32)]


_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "popscle_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp = C.c_void_p
        L.orc_phred2err.restype = C.c_double
        L.orc_phred2err.argtypes = [C.c_int]
        L.orc_phred2mat.restype = C.c_double
        L.orc_phred2mat.argtypes = [C.c_int]
        L.orc_log_add.restype = C.c_double
        L.orc_log_add.argtypes = [C.c_double, C.c_double]
        L.orc_demux_pair_pg.restype = None
        L.orc_demux_pair_pg.argtypes = [vp, vp, C.c_int64, C.c_int, vp, vp]
        L.orc_demux.restype = C.c_int
        L.orc_demux.argtypes = [C.POINTER(OPileup), C.c_int, vp, vp, C.c_int, vp, C.c_double, C.c_int, C.c_int, vp, vp, C.c_int]
        L.orc_fmx_pair_pileup.restype = C.c_double
        L.orc_fmx_pair_pileup.argtypes = [vp, vp, C.c_int64, C.c_double, vp, vp]
        L.orc_fmx_merge.restype = None
        L.orc_fmx_merge.argtypes = [vp, vp, vp, vp, vp, C.c_double]
        L.orc_fmx_run.restype = C.c_int
        L.orc_fmx_run.argtypes = [C.POINTER(OPileup), C.POINTER(OFmxOpts), vp, vp, vp, vp, C.POINTER(OFmxResult), vp, vp, C.c_int]
        L.orc_fmx_run_aux.restype = C.c_int
        L.orc_fmx_run_aux.argtypes = [C.POINTER(OPileup), C.POINTER(OFmxOpts), vp, vp, vp, vp, C.POINTER(OFmxResult), C.c_int, vp, vp]
        L.orc_fmx_estep.restype = C.c_int
        L.orc_fmx_estep.argtypes = [C.POINTER(OPileup), vp, vp, C.c_int, C.c_double, C.c_int, C.c_int, vp, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data if a is not None else None


def demux(plp, gp, has_gp, alphas, doublet_prior=0.5, cell_begin=0, cell_end=None, want_grid=False, n_threads=1):
    """plp: popscle_b200.capi.Pileup (or anything with the same attributes)."""
    L = lib()
    gp = np.ascontiguousarray(gp, dtype=np.float64)
    nv = gp.shape[1]
    al = np.ascontiguousarray(alphas, dtype=np.float64)
    hg = np.ascontiguousarray(has_gp, dtype=np.uint8) if has_gp is not None else None
    ce = plp.n_cells if cell_end is None else cell_end
    out = np.zeros(ce - cell_begin, dtype=DEMUX_CELL_DTYPE)
    grid = np.zeros((ce - cell_begin, nv, nv, len(al))) if want_grid else None
    cs = plp.c_struct(OPileup)
    rc = L.orc_demux(C.byref(cs), nv, _p(gp), _p(hg), len(al), _p(al), doublet_prior, cell_begin, ce, _p(out), _p(grid), n_threads)
    assert rc == 0, rc
    return (out, grid) if want_grid else out


def demux_pair_pg(allele, qual, alphas):
    L = lib()
    a = np.ascontiguousarray(allele, dtype=np.uint8)
    q = np.ascontiguousarray(qual, dtype=np.uint8)
    al = np.ascontiguousarray(alphas, dtype=np.float64)
    pg = np.zeros(len(al) * 9)
    L.orc_demux_pair_pg(_p(a), _p(q), len(a), len(al), _p(al), _p(pg))
    return pg.reshape(len(al), 3, 3)


def fmx_pair_pileup(allele, qual, alpha=0.5):
    L = lib()
    a = np.ascontiguousarray(allele, dtype=np.uint8)
    q = np.ascontiguousarray(qual, dtype=np.uint8)
    gls = np.zeros(9)
    cnt = np.zeros(3, dtype=np.int32)
    ld = L.orc_fmx_pair_pileup(_p(a), _p(q), len(a), alpha, _p(gls), _p(cnt))
    return gls, cnt, ld


def fmx_merge(gls_a, cnt_a, ld_a, gls_b, cnt_b, ld_b):
    L = lib()
    g = np.array(gls_a, dtype=np.float64)
    c = np.array(cnt_a, dtype=np.int32)
    ld = np.array([ld_a], dtype=np.float64)
    gb = np.ascontiguousarray(gls_b, dtype=np.float64)
    cb = np.ascontiguousarray(cnt_b, dtype=np.int32)
    L.orc_fmx_merge(_p(g), _p(c), _p(ld), _p(gb), _p(cb), float(ld_b))
    return g, c, float(ld[0])


def fmx_opts(n_clusters, doublet_prior=0.5, geno_error=0.1, max_iter=10, early_stop=True, frac_init_clust=1.0,
             singlet_score_thres=-1e300, mode_old=False, randomize_singlet_score=False, seed=0, bf_thres=5.41, iter_init=0,
             keep_init_missing=False):
    """iter_init > 0 (the reference's default is 10) together with mode_old runs freemuxlet-old's own vote seeding."""
    return OFmxOpts(n_clusters, doublet_prior, geno_error, max_iter, int(early_stop), frac_init_clust,
                    singlet_score_thres, int(mode_old), int(randomize_singlet_score), int(seed), bf_thres, int(iter_init),
                    int(keep_init_missing))


def fmx_run(plp, opts, init_clust=None, want_clusters=False, want_pair_gl=False, want_llk=False, n_threads=1):
    L = lib()
    out = np.zeros(plp.n_cells, dtype=FMX_CELL_DTYPE)
    res = OFmxResult()
    ic = np.ascontiguousarray(init_clust, dtype=np.int32) if init_clust is not None else None
    nS = opts.n_clusters
    gl = np.empty((plp.n_snps, nS, 9)) if want_clusters else None
    cnt = np.empty((plp.n_snps, nS, 3), dtype=np.int32) if want_clusters else None
    pgl = np.empty((plp.n_pairs, 9)) if want_pair_gl else None
    llk = np.empty((plp.n_cells, nS * (nS + 1) // 2)) if want_llk else None
    cs = plp.c_struct(OPileup)
    rc = L.orc_fmx_run(C.byref(cs), C.byref(opts), _p(ic), _p(out), _p(gl), _p(cnt), C.byref(res), _p(pgl), _p(llk), n_threads)
    assert rc == 0, rc
    return dict(cells=out, res=res, clust_gl=gl, clust_cnt=cnt, pair_gl=pgl, llk=llk)


def fmx_run_aux(plp, opts, init_clust=None, n_threads=1):
    """orc_fmx_run_aux: the run + the cluster pileups of the initial assignment (--aux-files)."""
    L = lib()
    out = np.zeros(plp.n_cells, dtype=FMX_CELL_DTYPE)
    res = OFmxResult()
    ic = np.ascontiguousarray(init_clust, dtype=np.int32) if init_clust is not None else None
    nS = opts.n_clusters
    gl, gl0 = np.empty((plp.n_snps, nS, 9)), np.empty((plp.n_snps, nS, 9))
    cnt, cnt0 = np.empty((plp.n_snps, nS, 3), dtype=np.int32), np.empty((plp.n_snps, nS, 3), dtype=np.int32)
    cs = plp.c_struct(OPileup)
    rc = L.orc_fmx_run_aux(C.byref(cs), C.byref(opts), _p(ic), _p(out), _p(gl), _p(cnt), C.byref(res), n_threads, _p(gl0), _p(cnt0))
    assert rc == 0, rc
    return out, res, gl, cnt, gl0, cnt0


def fmx_estep(plp, pair_gl, clust_gl, nS, geno_error, cell_begin=0, cell_end=None, n_threads=1):
    L = lib()
    ce = plp.n_cells if cell_end is None else cell_end
    llk = np.empty((ce - cell_begin, nS * (nS + 1) // 2))
    cs = plp.c_struct(OPileup)
    pg = np.ascontiguousarray(pair_gl)
    cg = np.ascontiguousarray(clust_gl)
    L.orc_fmx_estep(C.byref(cs), _p(pg), _p(cg), nS, geno_error, cell_begin, ce, _p(llk), n_threads)
    return llk
