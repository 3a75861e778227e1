"""ctypes binding of libpopscle_b200.so (include/popscle_b200.h).

Thin host-side mirror used by the tests, bench.py and the multi-GPU drivers; the C++ CLI host
(popscle_b200/host/) links the same library directly.  There is NO CPU fallback here: if the
shared library is missing, or no sm_100 device is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PSCL_LIB_PATH") or os.path.join(_HERE, "libpopscle_b200.so")  # override: A/B builds of the same sources

PSCL_OK = 0
PSCL_SNG, PSCL_DBL, PSCL_AMB = 0, 1, 2
TYPE_NAMES = {0: "SNG", 1: "DBL", 2: "AMB"}


class PsclError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"popscle_b200 error {code}: {msg}")
        self.code = code


class CPileup(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("n_snps", C.c_int32), ("n_pairs", C.c_int64), ("n_reads", C.c_int64),
                ("cell_ptr", C.c_void_p), ("pair_snp", C.c_void_p), ("pair_read_ptr", C.c_void_p),
                ("read_allele", C.c_void_p), ("read_qual", C.c_void_p), ("snp_af", C.c_void_p),
                ("pair_read_ptr32", C.c_void_p), ("read_aq", C.c_void_p),
                ("cell_first_snp", C.c_void_p), ("pair_snp_delta16", C.c_void_p), ("pair_nreads8", C.c_void_p),
                ("pair_snp_delta8", C.c_void_p), ("snp_gap_big", C.c_void_p), ("cell_gap_big_ptr", C.c_void_p),
                ("pair_nreads2", C.c_void_p), ("nreads_big", C.c_void_p), ("nreads_big_ptr", C.c_void_p),
                ("n_gap_big", C.c_int64), ("n_nreads_big", C.c_int64),
                ("read_packed", C.c_void_p), ("read_palette", C.c_void_p), ("read_bits", C.c_int32), ("reserved_", C.c_int32),
                ("cell_read_ptr", C.c_void_p)]


class CGeno(C.Structure):
    _fields_ = [("n_samples", C.c_int32), ("gp", C.c_void_p), ("has_gp", C.c_void_p),
                ("gp_f32", C.c_void_p), ("gt8", C.c_void_p), ("geno_err_snp", C.c_void_p), ("geno_err", C.c_double)]


@dataclasses.dataclass
class RawGeno:
    """ABI 4 genotype input: the reader's posteriors before the geno-error mixing (sc_drop_seq.cpp:287-315), mixed on the
    device.  Give `gt8` (uint8 [V][nv] hard calls 0/1/2) or `gp_f32` (float32 [V][nv][3]); `err` is the error rate
    (--geno-error-offset), `err_snp` a per-SNP override (--geno-error-coeff with R2)."""
    gt8: np.ndarray | None = None
    gp_f32: np.ndarray | None = None
    err: float = 0.1
    err_snp: np.ndarray | None = None


class CDemuxOpts(C.Structure):
    _fields_ = [("n_alpha", C.c_int32), ("alphas", C.c_void_p), ("doublet_prior", C.c_double)]


class CFmxOpts(C.Structure):
    _fields_ = [("n_clusters", C.c_int32), ("doublet_prior", C.c_double), ("geno_error", C.c_double),
                ("max_iter", C.c_int32), ("early_stop", C.c_int32), ("frac_init_clust", C.c_double),
                ("singlet_score_thres", C.c_double), ("mode_old", C.c_int32),
                ("randomize_singlet_score", C.c_int32), ("seed", C.c_int32),
                ("bf_thres", C.c_double), ("iter_init", C.c_int32), ("keep_init_missing", C.c_int32)]


class CMultiTiming(C.Structure):
    _fields_ = [("n_gpus", C.c_int32), ("iters", C.c_int32), ("total_ms", C.c_double), ("seed_ms", C.c_double),
                ("allreduce_ms", C.c_double), ("allreduce_bytes", C.c_int64),
                ("upload_ms", C.c_double * 16), ("setup_ms", C.c_double * 16), ("compute_ms", C.c_double * 16),
                ("kernel_ms", C.c_double * 16), ("units", C.c_int64 * 16)]


class CFmxResult(C.Structure):
    _fields_ = [("n_iter", C.c_int32), ("n_changed", C.c_int32), ("n_singlet", C.c_int32),
                ("n_doublet", C.c_int32), ("n_ambiguous", C.c_int32)]


# numpy views of the per-cell records (layout == the C structs; checked against sizeof in tests)
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
assert DEMUX_CELL_DTYPE.itemsize == 160

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
assert FMX_CELL_DTYPE.itemsize == 160


@dataclass
class Pileup:
    """Host-side flat pileup (the content of sc_dropseq_lib_t after load_from_plp,
    reference sc_drop_seq.cpp:103-384), cell-major CSR."""
    n_cells: int
    n_snps: int
    cell_ptr: np.ndarray       # int64 [C+1]
    pair_snp: np.ndarray       # int32 [P]
    pair_read_ptr: np.ndarray  # int64 [P+1]
    read_allele: np.ndarray    # uint8 [N]
    read_qual: np.ndarray      # uint8 [N]
    snp_af: np.ndarray | None = None  # float64 [V]

    def __post_init__(self):
        self.cell_ptr = np.ascontiguousarray(self.cell_ptr, dtype=np.int64)
        self.pair_snp = np.ascontiguousarray(self.pair_snp, dtype=np.int32)
        self.pair_read_ptr = np.ascontiguousarray(self.pair_read_ptr, dtype=np.int64)
        self.read_allele = np.ascontiguousarray(self.read_allele, dtype=np.uint8)
        self.read_qual = np.ascontiguousarray(self.read_qual, dtype=np.uint8)
        if self.snp_af is not None:
            self.snp_af = np.ascontiguousarray(self.snp_af, dtype=np.float64)

    @property
    def n_pairs(self) -> int:
        return int(self.pair_snp.shape[0])

    @property
    def n_reads(self) -> int:
        return int(self.read_allele.shape[0])

    def compact(self):
        """(pair_read_ptr32, read_aq): the compact arrays of ABI 2 (uint32 offsets, allele<<6|qual), cached."""
        c = getattr(self, "_compact", None)
        if c is None:
            if self.n_reads >= 1 << 32 or (self.read_qual > 63).any() or (self.read_allele > 2).any():
                raise PsclError(-1, "compact pileup needs n_reads < 2^32, qual <= 63, allele <= 2")
            c = (np.ascontiguousarray(self.pair_read_ptr, dtype=np.uint32),
                 np.ascontiguousarray((self.read_allele << 6) | self.read_qual, dtype=np.uint8))
            self._compact = c
        return c

    def compact3(self):
        """(cell_first_snp, pair_snp_delta16, pair_nreads8): the delta-coded pair arrays of ABI 3, or None when a
        SNP gap >= 65536 or a pair with >= 256 base-calls rules them out."""
        c = getattr(self, "_compact3", 0)
        if c == 0:
            c = None
            nrd = np.diff(self.pair_read_ptr)
            P = self.n_pairs
            first = np.zeros(self.n_cells, dtype=np.int32)
            starts = self.cell_ptr[:-1]
            nonempty = self.cell_ptr[1:] > starts
            delta = np.zeros(P, dtype=np.int64)
            if P:
                delta[1:] = np.diff(self.pair_snp.astype(np.int64))
                delta[starts[nonempty]] = 0
                first[nonempty] = self.pair_snp[starts[nonempty]]
            if P == 0 or (delta.min() >= 0 and delta.max() < 65536 and nrd.max() < 256):
                c = (first, delta.astype(np.uint16), nrd.astype(np.uint8))
            self._compact3 = c
        return c

    def cell_read_ptr(self):
        """ABI 7: first base-call of every cell, [C+1] int64 (cached; a caller may pin its own copy in `_cell_read_ptr`)."""
        c = getattr(self, "_cell_read_ptr", None)
        if c is None:
            c = np.ascontiguousarray(np.asarray(self.pair_read_ptr, dtype=np.int64)[np.asarray(self.cell_ptr, dtype=np.int64)])
            self._cell_read_ptr = c
        return c

    def compact4(self):
        """The ABI-6 pair arrays (1.25 B per pair): (cell_first_snp, pair_snp_delta8, snp_gap_big, cell_gap_big_ptr,
        pair_nreads2, nreads_big, nreads_big_ptr), or None when a pair has no or >= 256 base-calls."""
        c = getattr(self, "_compact4", 0)
        if c == 0:
            c = None
            P, Cn = self.n_pairs, self.n_cells
            nrd = np.diff(self.pair_read_ptr)
            starts = self.cell_ptr[:-1]
            nonempty = self.cell_ptr[1:] > starts
            first = np.zeros(Cn, dtype=np.int32)
            gap = np.zeros(P, dtype=np.int64)
            if P:
                gap[1:] = np.diff(self.pair_snp.astype(np.int64))
                gap[starts[nonempty]] = 0
                first[nonempty] = self.pair_snp[starts[nonempty]]
            if P == 0 or (gap.min() >= 0 and gap.max() < (1 << 32) and nrd.min() >= 1 and nrd.max() < 256):
                big = gap >= 255
                d8 = np.where(big, 255, gap).astype(np.uint8)
                gap_big = gap[big].astype(np.uint32)
                cum_big = np.concatenate([[0], np.cumsum(big)]).astype(np.int64)
                cell_big_ptr = np.ascontiguousarray(cum_big[self.cell_ptr])
                esc = nrd >= 4
                f2 = np.where(esc, 0, nrd).astype(np.uint8)
                pad = (-P) % 4
                f2p = np.concatenate([f2, np.zeros(pad, np.uint8)]).reshape(-1, 4)
                n2 = (f2p[:, 0] | (f2p[:, 1] << 2) | (f2p[:, 2] << 4) | (f2p[:, 3] << 6)).astype(np.uint8)
                nbig = nrd[esc].astype(np.uint8)
                cum_esc = np.concatenate([[0], np.cumsum(esc)]).astype(np.int64)
                blocks = np.minimum(np.arange(P // 1024 + 2, dtype=np.int64) * 1024, P)
                nbig_ptr = np.ascontiguousarray(cum_esc[blocks])
                c = (first, d8, gap_big, cell_big_ptr, np.ascontiguousarray(n2), nbig, nbig_ptr)
            self._compact4 = c
        return c

    def packed_reads(self):
        """(read_packed, read_palette, bits): the base-calls as 4/5/6-bit indices into the palette of distinct
        allele<<6|qual bytes (ABI 6), or None when there are more than 64 distinct values."""
        c = getattr(self, "_packed_reads", 0)
        if c == 0:
            c = None
            _, aq = self.compact()
            pal, idx = np.unique(aq, return_inverse=True)
            if len(pal) <= 64:
                bits = 4 if len(pal) <= 16 else 5 if len(pal) <= 32 else 6
                n = len(aq)
                b = ((idx.astype(np.uint8)[:, None] >> np.arange(bits, dtype=np.uint8)) & 1).astype(np.uint8).ravel()  # little-endian bit string
                b = np.concatenate([b, np.zeros((-len(b)) % 8 + 8, np.uint8)])
                packed = np.packbits(b, bitorder="little")
                palette = np.zeros(1 << bits, dtype=np.uint8)
                palette[:len(pal)] = pal
                assert len(packed) >= (n * bits + 7) // 8 + 1
                c = (np.ascontiguousarray(packed), palette, bits)
            self._packed_reads = c
        return c

    def c_struct(self, cls=CPileup, compact=False):
        """compact: False = wide arrays; True / 2 = ABI 2 (32-bit offsets, packed reads); 3 = ABI 3 (16-bit SNP gaps, 8-bit
        counts); 4 = ABI 6 (8-bit gaps, 2-bit counts, exceptions on the side) — each falls back to the previous form when
        the pileup does not fit it."""
        s = cls()
        s.n_cells, s.n_snps, s.n_pairs, s.n_reads = self.n_cells, self.n_snps, self.n_pairs, self.n_reads
        s.cell_ptr = self.cell_ptr.ctypes.data
        s.pair_snp = self.pair_snp.ctypes.data
        s.pair_read_ptr = self.pair_read_ptr.ctypes.data
        s.read_allele = self.read_allele.ctypes.data
        s.read_qual = self.read_qual.ctypes.data
        s.snp_af = self.snp_af.ctypes.data if self.snp_af is not None else None
        if compact:  # only the compact arrays cross the ABI
            p32, aq = self.compact()
            s.pair_read_ptr32, s.read_aq = p32.ctypes.data, aq.ctypes.data
            s.pair_read_ptr = s.read_allele = s.read_qual = None
            c4 = self.compact4() if compact == 4 else None
            c3 = self.compact3() if compact == 3 or (compact == 4 and c4 is None) else None
            if c4 is not None:
                first, d8, gbig, cbp, n2, nbig, nbp = c4
                s.cell_first_snp, s.pair_snp_delta8, s.cell_gap_big_ptr = first.ctypes.data, d8.ctypes.data, cbp.ctypes.data
                s.snp_gap_big = gbig.ctypes.data if len(gbig) else None
                s.pair_nreads2, s.nreads_big_ptr = n2.ctypes.data, nbp.ctypes.data
                s.nreads_big = nbig.ctypes.data if len(nbig) else None
                s.n_gap_big, s.n_nreads_big = len(gbig), len(nbig)
                s.pair_snp = s.pair_read_ptr32 = None
                crp = self.cell_read_ptr()  # ABI 7: lets the run slice counts and base-calls with the gaps
                s.cell_read_ptr = crp.ctypes.data
                pr = self.packed_reads()
                if pr is not None:  # 4-6 bits per base-call instead of 8
                    s.read_packed, s.read_palette, s.read_bits = pr[0].ctypes.data, pr[1].ctypes.data, pr[2]
                    s.read_aq = None
            elif c3 is not None:  # ABI 3: only deltas and counts for the pair arrays
                s.cell_first_snp, s.pair_snp_delta16, s.pair_nreads8 = (x.ctypes.data for x in c3)
                s.pair_snp = s.pair_read_ptr32 = None
        return s

    def slice_cells(self, c0: int, c1: int) -> "Pileup":
        """Barcode shard [c0, c1) as an independent pileup (demuxlet multi-GPU sharding)."""
        p0, p1 = int(self.cell_ptr[c0]), int(self.cell_ptr[c1])
        r0, r1 = int(self.pair_read_ptr[p0]), int(self.pair_read_ptr[p1])
        return Pileup(c1 - c0, self.n_snps, self.cell_ptr[c0:c1 + 1] - p0, self.pair_snp[p0:p1],
                      self.pair_read_ptr[p0:p1 + 1] - r0, self.read_allele[r0:r1], self.read_qual[r0:r1], self.snp_af)

    def slice_snps(self, v0: int, v1: int) -> "Pileup":
        """SNP shard [v0, v1): same cells (global ids), only the pairs of those SNPs; SNP ids stay
        global so the AF table and cluster table index the same way (freemuxlet SNP sharding)."""
        keep = (self.pair_snp >= v0) & (self.pair_snp < v1)
        pair_cell = np.repeat(np.arange(self.n_cells, dtype=np.int64), np.diff(self.cell_ptr))
        cnt = np.bincount(pair_cell[keep], minlength=self.n_cells)
        cell_ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        nrd = np.diff(self.pair_read_ptr)
        rkeep = np.repeat(keep, nrd)
        prp = np.concatenate([[0], np.cumsum(nrd[keep])]).astype(np.int64)
        return Pileup(self.n_cells, self.n_snps, cell_ptr, self.pair_snp[keep], prp, self.read_allele[rkeep],
                      self.read_qual[rkeep], self.snp_af)


_lib = None


def load_library() -> C.CDLL:
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PsclError(-2, f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    sigs = {
        "pscl_abi_version": (C.c_int, []),
        "pscl_create": (C.c_int, [C.c_int, C.POINTER(vp), C.c_char_p, C.c_size_t]),
        "pscl_destroy": (None, [vp]),
        "pscl_last_error": (C.c_char_p, [vp]),
        "pscl_stream": (vp, [vp]),
        "pscl_set_stream": (C.c_int, [vp, vp]),
        "pscl_sync": (C.c_int, [vp]),
        "pscl_launch_count": (i64, [vp]),
        "pscl_set_partial_budget": (C.c_int, [vp, C.c_size_t]),
        "pscl_debug_fail_alloc": (C.c_int, [vp, C.c_int]),
        "pscl_plp_upload": (C.c_int, [vp, C.POINTER(CPileup), C.POINTER(vp)]),
        "pscl_plp_free": (None, [vp, vp]),
        "pscl_demux_set_geno": (C.c_int, [vp, C.POINTER(CGeno), i32]),
        "pscl_demux_score": (C.c_int, [vp, vp, C.POINTER(CDemuxOpts), i32, i32]),
        "pscl_demux_fetch": (C.c_int, [vp, vp, vp]),
        "pscl_demux_keep_grid": (C.c_int, [vp, C.c_int]),
        "pscl_demux_force_general": (C.c_int, [vp, C.c_int]),
        "pscl_demux_select_kernel": (C.c_int, [vp, C.c_int]),
        "pscl_demux_run": (C.c_int, [vp, C.POINTER(CPileup), C.POINTER(CGeno), C.POINTER(CDemuxOpts), vp, vp]),
        "pscl_demux_last_kernel_ms": (C.c_int, [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "pscl_demux_last_kernel": (C.c_int, [vp]),
        "pscl_fmx_run": (C.c_int, [vp, C.POINTER(CPileup), C.POINTER(CFmxOpts), vp, vp, vp, vp, C.POINTER(CFmxResult)]),
        "pscl_fmx_run_aux": (C.c_int, [vp, C.POINTER(CPileup), C.POINTER(CFmxOpts), vp, vp, vp, vp, C.POINTER(CFmxResult), vp, vp]),
        "pscl_fmx_init": (C.c_int, [vp, vp, C.POINTER(CFmxOpts)]),
        "pscl_fmx_stage1": (C.c_int, [vp, vp]),
        "pscl_fmx_seed": (C.c_int, [vp, vp, vp, vp]),
        "pscl_fmx_mstep": (C.c_int, [vp, vp]),
        "pscl_fmx_estep": (C.c_int, [vp, i32, vp]),
        "pscl_fmx_classify": (C.c_int, [vp, vp, vp, C.POINTER(CFmxResult)]),
        "pscl_fmx_fetch": (C.c_int, [vp, vp, vp, vp]),
        "pscl_fmx_last_kernel_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "pscl_multi_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(vp), C.c_char_p, C.c_size_t]),
        "pscl_multi_destroy": (None, [vp]),
        "pscl_multi_last_error": (C.c_char_p, [vp]),
        "pscl_multi_size": (C.c_int, [vp]),
        "pscl_multi_ctx": (vp, [vp, C.c_int]),
        "pscl_multi_demux_run": (C.c_int, [vp, C.POINTER(CPileup), C.POINTER(CGeno), C.POINTER(CDemuxOpts), vp, vp]),
        "pscl_multi_fmx_run": (C.c_int, [vp, C.POINTER(CPileup), C.POINTER(CFmxOpts), vp, vp, vp, vp, C.POINTER(CFmxResult)]),
        "pscl_multi_last_timing": (C.c_int, [vp, C.POINTER(CMultiTiming)]),
        "pscl_bind_thread_to_device": (C.c_int, [C.c_int]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError = the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "pscl_abi_version", "pscl_create", "pscl_destroy", "pscl_last_error", "pscl_stream", "pscl_set_stream",
    "pscl_sync", "pscl_launch_count", "pscl_set_partial_budget", "pscl_debug_fail_alloc", "pscl_plp_upload", "pscl_plp_free",
    "pscl_demux_set_geno", "pscl_demux_score", "pscl_demux_fetch", "pscl_demux_keep_grid",
    "pscl_demux_force_general", "pscl_demux_select_kernel", "pscl_demux_run", "pscl_demux_last_kernel_ms", "pscl_demux_last_kernel", "pscl_fmx_run", "pscl_fmx_run_aux", "pscl_fmx_init",
    "pscl_fmx_stage1", "pscl_fmx_seed", "pscl_fmx_mstep", "pscl_fmx_estep", "pscl_fmx_classify",
    "pscl_fmx_fetch", "pscl_fmx_last_kernel_ms",
    "pscl_multi_create", "pscl_multi_destroy", "pscl_multi_last_error", "pscl_multi_size", "pscl_multi_ctx",
    "pscl_multi_demux_run", "pscl_multi_fmx_run", "pscl_multi_last_timing", "pscl_bind_thread_to_device",
]


class Context:
    """One pscl_ctx (one GPU)."""
    accepts_compact = True  # demux_run takes RawGeno genotypes and compact=2/3 pileup forms (cli.py asks)

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = self.lib.pscl_create(device, C.byref(h), err, len(err))
        if rc != PSCL_OK:
            raise PsclError(rc, err.value.decode())
        self.h = h
        self.device = device
        if stream is not None:
            self._chk(self.lib.pscl_set_stream(self.h, C.c_void_p(stream)))

    def _chk(self, rc: int):
        if rc != PSCL_OK:
            raise PsclError(rc, self.lib.pscl_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.pscl_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return int(self.lib.pscl_stream(self.h) or 0)

    def sync(self):
        self._chk(self.lib.pscl_sync(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.pscl_launch_count(self.h))

    def debug_fail_alloc(self, nth: int):
        """test knob: the nth device allocation from now fails (PSCL_ENOMEM)"""
        self._chk(self.lib.pscl_debug_fail_alloc(self.h, int(nth)))

    def set_partial_budget(self, nbytes: int):
        self._chk(self.lib.pscl_set_partial_budget(self.h, nbytes))

    # ---- pileup ----
    def upload(self, plp: Pileup, compact: bool = False) -> "DevicePileup":
        cs = plp.c_struct(compact=compact)
        out = C.c_void_p()
        self._chk(self.lib.pscl_plp_upload(self.h, C.byref(cs), C.byref(out)))
        return DevicePileup(self, out, plp.n_cells, plp.n_snps, plp.n_pairs, plp.n_reads)

    # ---- demuxlet ----
    @staticmethod
    def _geno(gp, has_gp):
        """CGeno + the arrays it points into.  `gp`: float64 [V][nv][3] (the mixed table), or a RawGeno (ABI 4: the
        reader's float32 posteriors or uint8 hard calls plus the genotype error, mixed on the device)."""
        g = CGeno()
        keep = []
        if isinstance(gp, RawGeno):
            if gp.gt8 is not None:
                a = np.ascontiguousarray(gp.gt8, dtype=np.uint8)
                assert a.ndim == 2
                g.gt8 = a.ctypes.data
            else:
                a = np.ascontiguousarray(gp.gp_f32, dtype=np.float32)
                assert a.ndim == 3 and a.shape[2] == 3, a.shape
                g.gp_f32 = a.ctypes.data
            keep.append(a)
            if gp.err_snp is not None:
                e = np.ascontiguousarray(gp.err_snp, dtype=np.float64)
                assert e.shape == (a.shape[0],)
                g.geno_err_snp = e.ctypes.data
                keep.append(e)
            g.geno_err = float(gp.err)
        else:
            a = np.ascontiguousarray(gp, dtype=np.float64)
            assert a.ndim == 3 and a.shape[2] == 3, a.shape
            g.gp = a.ctypes.data
            keep.append(a)
        g.n_samples = a.shape[1]
        if has_gp is not None:
            hg = np.ascontiguousarray(has_gp, dtype=np.uint8)
            g.has_gp = hg.ctypes.data
            keep.append(hg)
        return g, keep, a.shape[0], a.shape[1]

    def demux_set_geno(self, gp, has_gp: np.ndarray | None, n_snps: int):
        g, keep, V, nv = self._geno(gp, has_gp)
        assert V == n_snps, (V, n_snps)
        self._chk(self.lib.pscl_demux_set_geno(self.h, C.byref(g), n_snps))
        self.sync()  # the sources are pageable numpy buffers
        self._nv = nv

    def demux_keep_grid(self, enable: bool):
        self._chk(self.lib.pscl_demux_keep_grid(self.h, int(enable)))
        self._keep = enable

    def demux_force_general(self, enable: bool):
        self._chk(self.lib.pscl_demux_force_general(self.h, int(enable)))

    def demux_select_kernel(self, which: int):
        """0 = auto, 1 = k_demux_default on genotype rows, 2 = k_demux_general, 3 = k_demux_cls, 4 = k_demux_poly,
        5 = k_demux_ab, 6 = k_demux_default on dictionary-coded genotypes when the table allows (see popscle_b200.h)."""
        self._chk(self.lib.pscl_demux_select_kernel(self.h, int(which)))

    def demux_last_kernel(self) -> int:
        """which accumulation kernel the last demux_score launched (demux_select_kernel's numbering)"""
        return int(self.lib.pscl_demux_last_kernel(self.h))

    def demux_score(self, dplp: "DevicePileup", alphas, doublet_prior: float = 0.5, cell_begin: int = 0,
                    cell_end: int | None = None):
        al = np.ascontiguousarray(alphas, dtype=np.float64)
        o = CDemuxOpts(len(al), al.ctypes.data, doublet_prior)
        ce = dplp.n_cells if cell_end is None else cell_end
        self._chk(self.lib.pscl_demux_score(self.h, dplp.h, C.byref(o), cell_begin, ce))
        self._dm_shape = (ce - cell_begin, self._nv, self._nv, len(al))

    def demux_fetch(self, want_grid: bool = False):
        n = self._dm_shape[0]
        out = np.zeros(n, dtype=DEMUX_CELL_DTYPE)
        grid = np.empty(self._dm_shape, dtype=np.float64) if want_grid else None
        self._chk(self.lib.pscl_demux_fetch(self.h, out.ctypes.data, grid.ctypes.data if want_grid else None))
        return (out, grid) if want_grid else out

    def demux_last_kernel_ms(self):
        a, b = C.c_float(), C.c_float()
        self._chk(self.lib.pscl_demux_last_kernel_ms(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def demux_run(self, plp: Pileup, gp: np.ndarray, has_gp, alphas, doublet_prior: float = 0.5,
                  want_grid: bool = False, compact: bool = False, out: np.ndarray | None = None):
        """The one-call path of the CLI host: host buffers in, per-cell records out (into `out` when given: a caller that
        repeats the call keeps one — pinned — record array instead of a fresh pageable one per call)."""
        al = np.ascontiguousarray(alphas, dtype=np.float64)
        cs = plp.c_struct(compact=compact)
        g, keep, _, nv = self._geno(gp, has_gp)
        o = CDemuxOpts(len(al), al.ctypes.data, doublet_prior)
        if out is None:
            out = np.zeros(plp.n_cells, dtype=DEMUX_CELL_DTYPE)
        elif out.dtype != DEMUX_CELL_DTYPE or out.shape != (plp.n_cells,) or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous array of n_cells DEMUX_CELL_DTYPE records")
        grid = np.empty((plp.n_cells, nv, nv, len(al))) if want_grid else None
        self._chk(self.lib.pscl_demux_run(self.h, C.byref(cs), C.byref(g), C.byref(o), out.ctypes.data,
                                          grid.ctypes.data if want_grid else None))
        return (out, grid) if want_grid else out

    # ---- freemuxlet ----
    @staticmethod
    def fmx_opts(n_clusters: int, doublet_prior=0.5, geno_error=0.1, max_iter=10, early_stop=True,
                 frac_init_clust=1.0, singlet_score_thres=-1e300, mode_old=False, randomize_singlet_score=False, seed=0,
                 bf_thres=5.41, iter_init=0, keep_init_missing=False) -> CFmxOpts:
        """iter_init > 0 (the reference's default is 10) or a missing init_clust, together with mode_old, runs
        freemuxlet-old's own pairwise / vote seeding."""
        return CFmxOpts(n_clusters, doublet_prior, geno_error, max_iter, int(early_stop), frac_init_clust,
                        singlet_score_thres, int(mode_old), int(randomize_singlet_score), int(seed), bf_thres, int(iter_init),
                        int(keep_init_missing))

    def fmx_run(self, plp: Pileup, opts: CFmxOpts, init_clust: np.ndarray | None = None, want_clusters=False,
                compact: bool = False):
        cs = plp.c_struct(compact=compact)
        out = np.zeros(plp.n_cells, dtype=FMX_CELL_DTYPE)
        res = CFmxResult()
        ic = None
        if init_clust is not None:
            ic = np.ascontiguousarray(init_clust, dtype=np.int32)
        gl = cnt = None
        if want_clusters:
            gl = np.empty((plp.n_snps, opts.n_clusters, 9), dtype=np.float64)
            cnt = np.empty((plp.n_snps, opts.n_clusters, 3), dtype=np.int32)
        self._chk(self.lib.pscl_fmx_run(self.h, C.byref(cs), C.byref(opts), ic.ctypes.data if ic is not None else None,
                                        out.ctypes.data, gl.ctypes.data if gl is not None else None,
                                        cnt.ctypes.data if cnt is not None else None, C.byref(res)))
        return out, res, gl, cnt

    def fmx_run_aux(self, plp: Pileup, opts: CFmxOpts, init_clust: np.ndarray | None = None, compact: bool = False):
        """pscl_fmx_run_aux: the whole run + the cluster pileups of the initial assignment (--aux-files):
        (cells, result, clust_gl, clust_cnt, clust_gl0, clust_cnt0)."""
        cs = plp.c_struct(compact=compact)
        out = np.zeros(plp.n_cells, dtype=FMX_CELL_DTYPE)
        res = CFmxResult()
        ic = np.ascontiguousarray(init_clust, dtype=np.int32) if init_clust is not None else None
        shape = (plp.n_snps, opts.n_clusters)
        gl, gl0 = np.empty(shape + (9,), dtype=np.float64), np.empty(shape + (9,), dtype=np.float64)
        cnt, cnt0 = np.empty(shape + (3,), dtype=np.int32), np.empty(shape + (3,), dtype=np.int32)
        self._chk(self.lib.pscl_fmx_run_aux(self.h, C.byref(cs), C.byref(opts), ic.ctypes.data if ic is not None else None, out.ctypes.data,
                                            gl.ctypes.data, cnt.ctypes.data, C.byref(res), gl0.ctypes.data, cnt0.ctypes.data))
        return out, res, gl, cnt, gl0, cnt0

    # step-level freemuxlet API (SNP-sharded EM): *_dev arguments are raw device pointers (ints)
    def fmx_init(self, dplp: "DevicePileup", opts: CFmxOpts):
        self._chk(self.lib.pscl_fmx_init(self.h, dplp.h, C.byref(opts)))
        self._fmx_shape = (dplp.n_cells, dplp.n_snps, opts.n_clusters)

    def fmx_stage1(self, stage1_dev: int):
        self._chk(self.lib.pscl_fmx_stage1(self.h, C.c_void_p(stage1_dev)))

    def fmx_seed(self, stage1_dev: int, init_clust_dev: int | None, clust_dev: int):
        self._chk(self.lib.pscl_fmx_seed(self.h, C.c_void_p(stage1_dev),
                                         C.c_void_p(init_clust_dev) if init_clust_dev else None, C.c_void_p(clust_dev)))

    def fmx_mstep(self, clust_dev: int | None):
        self._chk(self.lib.pscl_fmx_mstep(self.h, C.c_void_p(clust_dev) if clust_dev else None))

    def fmx_estep(self, it: int, llk_dev: int):
        self._chk(self.lib.pscl_fmx_estep(self.h, it, C.c_void_p(llk_dev)))

    def fmx_classify(self, llk_dev: int, clust_dev: int) -> CFmxResult:
        res = CFmxResult()
        self._chk(self.lib.pscl_fmx_classify(self.h, C.c_void_p(llk_dev), C.c_void_p(clust_dev), C.byref(res)))
        return res

    def fmx_fetch(self, want_clusters: bool = False):
        nc, nv, ns = self._fmx_shape
        out = np.zeros(nc, dtype=FMX_CELL_DTYPE)
        gl = np.empty((nv, ns, 9), dtype=np.float64) if want_clusters else None
        cnt = np.empty((nv, ns, 3), dtype=np.int32) if want_clusters else None
        self._chk(self.lib.pscl_fmx_fetch(self.h, out.ctypes.data, gl.ctypes.data if want_clusters else None,
                                          cnt.ctypes.data if want_clusters else None))
        return out, gl, cnt

    def fmx_last_kernel_ms(self) -> float:
        a = C.c_float()
        self._chk(self.lib.pscl_fmx_last_kernel_ms(self.h, C.byref(a)))
        return a.value


class Multi:
    """pscl_multi: several GPUs driven from this process (barcode-sharded demuxlet, SNP-sharded freemuxlet)."""
    accepts_compact = True

    def __init__(self, gpu_ids=None, n_gpu: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        if gpu_ids is not None:
            ids = (C.c_int * len(gpu_ids))(*gpu_ids)
            rc = self.lib.pscl_multi_create(ids, len(gpu_ids), C.byref(h), err, len(err))
        else:
            rc = self.lib.pscl_multi_create(None, n_gpu, C.byref(h), err, len(err))
        if rc != PSCL_OK:
            raise PsclError(rc, err.value.decode())
        self.h = h

    def _chk(self, rc: int):
        if rc != PSCL_OK:
            raise PsclError(rc, self.lib.pscl_multi_last_error(self.h).decode())

    @property
    def size(self) -> int:
        return int(self.lib.pscl_multi_size(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.pscl_multi_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def select_demux_kernel(self, which: int):
        for i in range(self.size):
            rc = self.lib.pscl_demux_select_kernel(C.c_void_p(self.lib.pscl_multi_ctx(self.h, i)), int(which))
            if rc != PSCL_OK:
                raise PsclError(rc, "pscl_demux_select_kernel")

    def timing(self) -> dict:
        t = CMultiTiming()
        self._chk(self.lib.pscl_multi_last_timing(self.h, C.byref(t)))
        n = t.n_gpus
        return {"n_gpus": n, "iters": t.iters, "total_ms": t.total_ms, "seed_ms": t.seed_ms, "allreduce_ms": t.allreduce_ms,
                "allreduce_bytes": t.allreduce_bytes, "upload_ms": list(t.upload_ms)[:n], "setup_ms": list(t.setup_ms)[:n],
                "compute_ms": list(t.compute_ms)[:n], "kernel_ms": list(t.kernel_ms)[:n], "units": list(t.units)[:n]}

    def demux_run(self, plp: Pileup, gp, has_gp, alphas, doublet_prior: float = 0.5, want_grid: bool = False, compact=False):
        al = np.ascontiguousarray(alphas, dtype=np.float64)
        cs = plp.c_struct(compact=compact)
        g, keep, _, nv = Context._geno(gp, has_gp)
        o = CDemuxOpts(len(al), al.ctypes.data, doublet_prior)
        out = np.zeros(plp.n_cells, dtype=DEMUX_CELL_DTYPE)
        grid = np.empty((plp.n_cells, nv, nv, len(al))) if want_grid else None
        self._chk(self.lib.pscl_multi_demux_run(self.h, C.byref(cs), C.byref(g), C.byref(o), out.ctypes.data,
                                                grid.ctypes.data if want_grid else None))
        return (out, grid) if want_grid else out

    fmx_opts = staticmethod(Context.fmx_opts)

    def fmx_run(self, plp: Pileup, opts: CFmxOpts, init_clust=None, want_clusters=False, compact=False):
        cs = plp.c_struct(compact=compact)
        out = np.zeros(plp.n_cells, dtype=FMX_CELL_DTYPE)
        res = CFmxResult()
        ic = np.ascontiguousarray(init_clust, dtype=np.int32) if init_clust is not None else None
        gl = cnt = None
        if want_clusters:
            gl = np.empty((plp.n_snps, opts.n_clusters, 9), dtype=np.float64)
            cnt = np.empty((plp.n_snps, opts.n_clusters, 3), dtype=np.int32)
        self._chk(self.lib.pscl_multi_fmx_run(self.h, C.byref(cs), C.byref(opts), ic.ctypes.data if ic is not None else None,
                                              out.ctypes.data, gl.ctypes.data if gl is not None else None,
                                              cnt.ctypes.data if cnt is not None else None, C.byref(res)))
        return out, res, gl, cnt


def bind_to_device(device: int) -> bool:
    """Runs the calling thread on the CPUs of the GPU's NUMA node (pinned buffers allocated afterwards are node-local)."""
    return load_library().pscl_bind_thread_to_device(int(device)) == PSCL_OK


class DevicePileup:
    def __init__(self, ctx: Context, h, n_cells, n_snps, n_pairs, n_reads):
        self.ctx, self.h = ctx, h
        self.n_cells, self.n_snps, self.n_pairs, self.n_reads = n_cells, n_snps, n_pairs, n_reads

    def free(self):
        if self.h and self.ctx.h:
            self.ctx.lib.pscl_plp_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
