"""Output writers of `popscle demuxlet` / `popscle freemuxlet`, byte-compatible with the reference
given equal doubles: `.best` (cmd_cram_demuxlet.cpp:629, :993-1013), `.lmix`
(cmd_cram_freemux2.cpp:111, :161), `.clust1.samples.gz` (:660-665) and `.clust1.vcf.gz` (:608-658).

Python's % operator follows C printf for %d / %.2f / %.5f / %.2g / %.3g, so the rows are formatted
with the reference's own format strings.
"""
from __future__ import annotations

import gzip
import math
import time

import numpy as np

from .capi import TYPE_NAMES

BEST_HEADER = ("INT_ID\tBARCODE\tNUM.SNPS\tNUM.READS\tDROPLET.TYPE\tBEST.GUESS\tBEST.LLK\tNEXT.GUESS\tNEXT.LLK\t"
               "DIFF.LLK.BEST.NEXT\tBEST.POSTERIOR\tSNG.POSTERIOR\tSNG.BEST.GUESS\tSNG.BEST.LLK\tSNG.NEXT.GUESS\t"
               "SNG.NEXT.LLK\tSNG.ONLY.POSTERIOR\tDBL.BEST.GUESS\tDBL.BEST.LLK\tDIFF.LLK.SNG.DBL\n")
BEST_ROW = "%d\t%s\t%u\t%d\t%s\t%s,%s,%.2f\t%.2f\t%s,%s,%.2f\t%.2f\t%.2f\t%.2g\t%.2g\t%s\t%.2f\t%s\t%.2f\t%.5f\t%s,%s,%.2f\t%.2f\t%.2f\n"


def best_rows(cells, barcodes, uniq_reads, samples, alphas, min_total=0, min_umi=0, min_snp=0, totl_reads=None):
    """Rows in lexicographic barcode order; INT_ID is the rank in that order and also counts skipped
    droplets (the for-loop increments ncells before `continue`, cmd_cram_demuxlet.cpp:636-653)."""
    order = sorted(range(len(barcodes)), key=lambda i: barcodes[i].encode())  # std::map<std::string> = byte order
    for rank, i in enumerate(order):
        r = cells[i]
        tot = uniq_reads[i] if totl_reads is None else totl_reads[i]
        if tot < min_total or uniq_reads[i] < min_umi or r["n_snps"] < min_snp:
            continue
        if r["n_snps"] == 0:
            continue  # :653
        sm = lambda j: samples[int(j)]
        yield BEST_ROW % (rank, barcodes[i], int(r["n_snps"]), int(uniq_reads[i]), TYPE_NAMES[int(r["type"])],
                          sm(r["best_j"]), sm(r["best_k"]), alphas[int(r["best_a"])], r["best_llk"],
                          sm(r["next_j"]), sm(r["next_k"]), alphas[int(r["next_a"])], r["next_llk"],
                          r["best_llk"] - r["next_llk"], r["best_pp"], r["sng_pp"],
                          sm(r["sng_best"]), r["sng_best_llk"], sm(r["sng_next"]), r["sng_next_llk"], r["sng_only_pp"],
                          sm(r["dbl_best_j"]), sm(r["dbl_best_k"]), alphas[int(r["dbl_best_a"])], r["dbl_best_llk"],
                          r["sng_best_llk"] - r["dbl_best_llk"])


def write_best(path, cells, barcodes, uniq_reads, samples, alphas, **kw):
    with open(path, "w") as f:  # plain text: hts_open(..., "w") (cmd_cram_demuxlet.cpp:578)
        f.write(BEST_HEADER)
        for row in best_rows(cells, barcodes, uniq_reads, samples, alphas, **kw):
            f.write(row)


LMIX_HEADER = "INT_ID\tBARCODE\tNSNPs\tNREADs\tDBL.LLK\tSNG.LLK\tBF.SINGLET\tBF.SINGLET.PER.SNP\n"


LMIX_HEADER_OLD = "INT_ID\tBARCODE\tNSNPs\tNREADs\tDBL.LLK\tSNG.LLK\tLOG.BF\tBFpSNP\n"


def write_lmix(path, cells, barcodes, old=False):
    """old = freemuxlet-old: prints llk0-llk2 under the LOG.BF / BFpSNP headers (cmd_cram_freemuxlet.cpp:111,:163)"""
    with open(path, "w") as f:
        f.write(LMIX_HEADER_OLD if old else LMIX_HEADER)
        for i, r in enumerate(cells):
            d = (r["llk0"] - r["llk2"]) if old else (r["llk2"] - r["llk0"])
            per = d / r["n_snps"] if r["n_snps"] else (math.nan if d == 0 else math.copysign(math.inf, d))
            f.write("%d\t%s\t%d\t%d\t%.2f\t%.2f\t%.2f\t%s\n" % (i, barcodes[i], r["n_snps"], r["n_reads"], r["llk0"], r["llk2"], d,
                                                                _c_float(per, 4)))


def _c_float(x, prec):
    """%.{prec}lf with glibc's spelling of non-finite values"""
    if math.isnan(x):
        return "-nan" if math.copysign(1.0, x) < 0 else "nan"
    if math.isinf(x):
        return "-inf" if x < 0 else "inf"
    return "%.*f" % (prec, x)


SAMPLES_HEADER = BEST_HEADER
SAMPLES_ROW = "%d\t%s\t%d\t%d\t%s\t%d,%d\t%.2f\t%d,%d\t%.2f\t%.2f\t%.5f\t%.2g\t%d\t%.2f\t%d\t%.2f\t%.5f\t%d,%d\t%.2f\t%.2f\n"


def samples_rows(cells, barcodes):
    for i, r in enumerate(cells):  # cell-id order (cmd_cram_freemux2.cpp:662)
        yield SAMPLES_ROW % (i, barcodes[i], r["n_snps"], r["n_reads"], "AMB" if r["type"] == 2 else ("SNG" if r["type"] == 0 else "DBL"),
                             r["best_j"], r["best_k"], r["best_llk"], r["next_j"], r["next_k"], r["next_llk"],
                             r["best_llk"] - r["next_llk"], r["best_pp"], r["sng_pp"], r["sng_best"], r["sng_best_llk"],
                             r["sng_next"], r["sng_next_llk"], r["sng_only_pp"], r["dbl_best_j"], r["dbl_best_k"],
                             r["dbl_best_llk"], r["sng_best_llk"] - r["dbl_best_llk"])


def write_clust_samples(path, cells, barcodes):
    with gzip.open(path, "wt") as f:
        f.write(SAMPLES_HEADER)
        for row in samples_rows(cells, barcodes):
            f.write(row)


def clust_vcf_rows(sites, clust_gl, clust_cnt, observed, initial=False):
    """Per-cluster GT:GQ:DP:AD:PL:GP from the diagonal cluster GLs (cmd_cram_freemux2.cpp:623-656).  initial = the rows of
    --aux-files' .clust0.vcf.gz (:316-344), which differ in two details: the posterior is (prior * GL) / maxGL rather than
    prior * (GL / maxGL), and GQ is -0.1 * log10 (sic) rather than -10 * log10."""
    V, nS = clust_gl.shape[0], clust_gl.shape[1]
    for v in range(V):
        if not observed[v]:
            continue
        af = float(sites.af[v])
        gps = ((1. - af) * (1. - af), 2. * af * (1. - af), af * af)
        out = ["%s\t%d\t.\t%s\t%s\t.\tPASS\tAF=%.5f\tGT:GQ:DP:AD:PL:GP" % (sites.chrom[v], sites.pos[v], sites.ref[v][0], sites.alt[v][0], af)]
        for i in range(nS):
            g = (float(clust_gl[v, i, 0]), float(clust_gl[v, i, 4]), float(clust_gl[v, i, 8]))
            mx = max(g)
            pls = [int(-10.0 * math.log10(x / mx)) for x in g]
            pps = [(gps[k] * g[k] / mx if initial else gps[k] * (g[k] / mx)) + 1e-100 for k in range(3)]
            s = pps[0] + pps[1] + pps[2]
            pps = [x / s for x in pps]
            best = (0 if pps[0] > pps[2] else 2) if pps[0] > pps[1] else (1 if pps[1] > pps[2] else 2)
            gq = min(int((-0.1 if initial else -10) * math.log10(1.0 - pps[best] + 1e-100)), 255)
            n = clust_cnt[v, i]
            out.append("\t%d/%d:%d:%d:%d,%d:%d,%d,%d:%.3g,%.3g,%.3g" % (1 if best == 2 else 0, 1 if best > 0 else 0, gq, n[0], n[1], n[2],
                                                                        pls[0], pls[1], pls[2], pps[0], pps[1], pps[2]))
        yield "".join(out) + "\n"


def write_clust_vcf(path, sites, rid2chr, clust_gl, clust_cnt, observed, now=None, initial=False):
    nS = clust_gl.shape[1]
    ltm = time.localtime(now)
    with gzip.open(path, "wt") as f:
        f.write("##fileformat=VCFv4.2\n")
        # the reference prints 1970+tm_year (sic: tm_year counts from 1900), :610
        f.write("##fileDate=%04d%02d%02d\n" % (1970 + ltm.tm_year - 1900, ltm.tm_mon, ltm.tm_mday))
        f.write("##source=cramore-freemuxlet\n")
        for c in rid2chr:
            f.write("##contig=<ID=%s>\n" % c)
        f.write('##INFO=<ID=AF,Number=A,Type=Float,Description="Allele Frequency">\n')
        f.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        f.write('##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Phred-scale Genotype Quality">\n')
        f.write('##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Read Depth">\n')
        f.write('##FORMAT=<ID=AD,Number=R,Type=Integer,Description="Allelic Read Depth">\n')
        f.write('##FORMAT=<ID=PL,Number=G,Type=Integer,Description="Phred-scale genotype likelihood">\n')
        f.write('##FORMAT=<ID=GP,Number=G,Type=Float,Description="Posterior probability using pooled allele frequencies">\n')
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT" + "".join("\tCLUST%d" % i for i in range(nS)) + "\n")
        for row in clust_vcf_rows(sites, clust_gl, clust_cnt, observed, initial=initial):
            f.write(row)


def write_clust0_samples(path, cells, barcodes):
    """--aux-files: the initial cluster of every droplet, -1 = none (cmd_cram_freemux2.cpp:265-274)"""
    with gzip.open(path, "wt") as f:
        f.write("INT_ID\tBARCODE\tCLUST0\n")
        for i, b in enumerate(barcodes):
            f.write("%d\t%s\t%d\n" % (i, b, int(cells["init_clust"][i])))


def observed_snps(plp) -> np.ndarray:
    """snps_observed (cmd_cram_freemux2.cpp:279-287): SNPs covered by at least one kept droplet"""
    obs = np.zeros(plp.n_snps, dtype=bool)
    obs[plp.pair_snp] = True
    return obs
