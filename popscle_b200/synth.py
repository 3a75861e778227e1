"""Seeded synthetic pileups of the shapes BASELINE.json names.

Produces the flat arrays `sc_dropseq_lib_t::load_from_plp` would hold after reading dsc-pileup
CEL/VAR/PLP files (reference sc_drop_seq.cpp:103-384): reads already min-BQ filtered and capped.
Everything is vectorised numpy so config 2 (2e7 pairs) is generated in seconds.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .capi import Pileup

# (C cells, nv samples, V SNPs, mean SNPs per cell, alpha grid) of BASELINE.json configs[0..4]
CONFIGS = {
    1: dict(C=500, nv=4, V=3000, kbar=300, alphas=(0.0, 0.5)),
    2: dict(C=10_000, nv=8, V=100_000, kbar=2000, alphas=(0.0, 0.5)),
    3: dict(C=10_000, nv=8, V=100_000, kbar=2000, alphas=(0.0, 0.5)),
    4: dict(C=50_000, nv=64, V=1_000_000, kbar=5000, alphas=tuple(0.025 * i for i in range(21))),
    5: dict(C=100_000, nv=16, V=500_000, kbar=4000, alphas=(0.0, 0.5)),
}


@dataclass
class Synth:
    plp: Pileup
    geno: np.ndarray      # int8 [nv][V] donor genotypes (0/1/2 ALT copies)
    af: np.ndarray        # float64 [V]
    truth_d1: np.ndarray  # int32 [C]
    truth_d2: np.ndarray  # int32 [C], == d1 for singlets
    seed: int


def make_pileup(C: int, nv: int, V: int, kbar: float, seed: int, cap_bq: int = 20, min_bq: int = 13,
                doublet_frac: float = 0.1, cell_scale: float = 1.0, snp_seed: int | None = None,
                cell_seed: int | None = None) -> Synth:
    """SURVEY.md §8(d) generator.  cap_bq/min_bq are applied the way the loader does
    (sc_drop_seq.cpp:361-369), so `read_qual` is what add_read receives.
    cell_seed: independent cell-level draws against the SAME SNP-level truth (AF, donor genotypes): block `cell_seed` of a
    larger problem's cells, so that a barcode-sharded run can generate each block where it is scored."""
    rng = np.random.Generator(np.random.PCG64(seed))
    # snp_seed: independent SNP-level draws (AF, genotypes) with the SAME cell-level truth, used to give
    # every rank of an SNP-sharded run its own SNP range of the same cells
    rng_snp = rng if snp_seed is None else np.random.Generator(np.random.PCG64([seed, snp_seed]))
    af = np.round(rng_snp.uniform(0.05, 0.5, V), 5)
    geno = rng_snp.binomial(2, np.broadcast_to(af, (nv, V))).astype(np.int8)
    if snp_seed is not None:
        rng.uniform(0.05, 0.5, 1)  # keep the cell-level stream independent of V
    if cell_seed is not None:
        rng = np.random.Generator(np.random.PCG64([seed, 7919, cell_seed]))
    is_dbl = rng.random(C) < doublet_frac
    d1 = rng.integers(0, nv, C).astype(np.int32)
    d2 = ((d1 + rng.integers(1, max(nv, 2), C)) % nv).astype(np.int32)
    d2 = np.where(is_dbl, d2, d1).astype(np.int32)
    sigma = 0.5
    mu = np.log(kbar * cell_scale) - 0.5 * sigma * sigma
    K = np.clip(np.rint(rng.lognormal(mu, sigma, C)), min(50, max(V // 4, 1)), max(V // 4, 1)).astype(np.int64)
    cell = np.repeat(np.arange(C, dtype=np.int64), K)
    snp = rng.integers(0, V, cell.shape[0], dtype=np.int64)
    key = np.sort(cell * V + snp)  # cell-major, SNP ascending ...
    if key.shape[0]:
        key = key[np.concatenate(([True], key[1:] != key[:-1]))]  # ... duplicates dropped (np.unique, 5x faster)
    cell = key // V
    snp = (key - cell * V).astype(np.int32)
    P = key.shape[0]
    cell_ptr = np.zeros(C + 1, dtype=np.int64)
    np.cumsum(np.bincount(cell, minlength=C), out=cell_ptr[1:])
    nrd = 1 + np.minimum(rng.poisson(0.3, P), 7)
    prp = np.zeros(P + 1, dtype=np.int64)
    np.cumsum(nrd, out=prp[1:])
    N = int(prp[-1])
    rpair = np.repeat(np.arange(P, dtype=np.int64), nrd)
    rcell = cell[rpair]
    use2 = rng.random(N) < 0.5
    donor = np.where(use2, d2[rcell], d1[rcell])
    g = geno[donor, snp[rpair]]
    true_alt = rng.random(N) < (g * 0.5)
    q = rng.integers(13, 41, N)
    is_err = rng.random(N) < np.power(10.0, -q / 10.0)
    u = rng.random(N)
    allele = np.where(true_alt, 1, 0).astype(np.uint8)
    flipped = (1 - allele).astype(np.uint8)
    allele = np.where(is_err, np.where(u < 1.0 / 3.0, flipped, 2), allele).astype(np.uint8)
    keep = q >= min_bq
    assert keep.all()
    qual = np.minimum(q, cap_bq).astype(np.uint8)
    plp = Pileup(C, V, cell_ptr, snp, prp, allele, qual, af.astype(np.float64))
    return Synth(plp, geno, af, d1, d2, seed)


def make_config(cfg: int, seed: int | None = None, cell_scale: float = 1.0, cells: int | None = None) -> Synth:
    c = CONFIGS[cfg]
    return make_pileup(cells or c["C"], c["nv"], c["V"], c["kbar"], (20260101 + cfg) if seed is None else seed,
                       cell_scale=cell_scale)


def gt_to_gp(geno: np.ndarray, geno_error_offset: float = 0.1) -> np.ndarray:
    """`--field GT` genotype table the way load_from_plp builds it (sc_drop_seq.cpp:285-315):
    one-hot float32 posteriors (bcf_filtered_reader.cpp:385-409, no missing GT), per-SNP average
    started at 1e-10, then gps = (1-err)*gp + err*avg.  Returns float64 [V][nv][3]."""
    nv, V = geno.shape
    gp = np.zeros((V, nv, 3), dtype=np.float64)
    gp[np.arange(V)[:, None], np.arange(nv)[None, :], geno.T.astype(np.int64)] = 1.0
    gp = gp.astype(np.float32).astype(np.float64)
    # avgGPs[i%3] accumulates in sample order starting from 1e-10 (:288-292)
    avg = np.full((V, 3), 1e-10)
    for j in range(nv):
        avg = avg + gp[:, j, :]
    s = avg[:, 0] + avg[:, 1] + avg[:, 2]
    avg = avg / s[:, None]
    err = min(max(geno_error_offset, 0.0), 0.999)
    if err > 0:
        gp = (1 - err) * gp + err * avg[:, None, :]
    return np.ascontiguousarray(gp)
