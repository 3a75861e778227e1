// writers.cpp — see writers.h
#include "writers.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <ctime>
#include <numeric>

#include "tsv.h"

namespace pscl_host {
namespace {

// the role of hprintf (hts_utils.cpp:1013-1034): printf into a plain or gzip file
struct Out {
  FILE* f = nullptr;
  gzFile gz = nullptr;
  std::string path;
  Out(const std::string& p, bool gzip) : path(p) {
    if (gzip) gz = gzopen(p.c_str(), "wb"); else f = fopen(p.c_str(), "w");
    if (!f && !gz) throw host_error("Cannot open file " + p + " for writing");
  }
  ~Out() { if (f) fclose(f); if (gz) gzclose(gz); }
  void printf(const char* fmt, ...) __attribute__((format(printf, 2, 3))) {
    char stackbuf[4096];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(stackbuf, sizeof stackbuf, fmt, ap);
    va_end(ap);
    std::vector<char> big;
    const char* p = stackbuf;
    if (n >= (int)sizeof stackbuf) {
      big.resize((size_t)n + 1);
      va_start(ap, fmt);
      vsnprintf(big.data(), big.size(), fmt, ap);
      va_end(ap);
      p = big.data();
    }
    const bool ok = gz ? gzwrite(gz, p, (unsigned)n) == n : fwrite(p, 1, (size_t)n, f) == (size_t)n;
    if (!ok && n > 0) throw host_error("hprintf failed. Aborting.. (" + path + ")");
  }
};

const char* type_name(int t) { return t == PSCL_AMB ? "AMB" : (t == PSCL_SNG ? "SNG" : "DBL"); }

}  // namespace

void write_best(const std::string& path, const Loaded& L, const std::vector<pscl_demux_cell>& cells,
                const std::vector<double>& alphas, int min_total, int min_umi, int min_snp) {
  Out w(path, false);  // hts_open(..., "w"): plain text (cmd_cram_demuxlet.cpp:578)
  w.printf("INT_ID\tBARCODE\tNUM.SNPS\tNUM.READS\tDROPLET.TYPE\tBEST.GUESS\tBEST.LLK\tNEXT.GUESS\tNEXT.LLK\tDIFF.LLK.BEST.NEXT\t"
           "BEST.POSTERIOR\tSNG.POSTERIOR\tSNG.BEST.GUESS\tSNG.BEST.LLK\tSNG.NEXT.GUESS\tSNG.NEXT.LLK\tSNG.ONLY.POSTERIOR\t"
           "DBL.BEST.GUESS\tDBL.BEST.LLK\tDIFF.LLK.SNG.DBL\n");
  // rows in bc_map (std::map<std::string,...>) order; INT_ID counts skipped droplets too (:636-653)
  std::vector<int32_t> order(L.n_cells);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return L.barcodes[a] < L.barcodes[b]; });
  auto sm = [&](int32_t j) { return (j >= 0 && j < (int32_t)L.samples.size()) ? L.samples[j].c_str() : "."; };
  for (int32_t rank = 0; rank < L.n_cells; ++rank) {
    const int32_t i = order[rank];
    const pscl_demux_cell& r = cells[i];
    if (L.cell_totl_reads[i] < min_total || L.cell_uniq_reads[i] < min_umi || r.n_snps < min_snp) continue;
    if (r.n_snps == 0) continue;
    w.printf("%d\t%s\t%u\t%d\t%s\t%s,%s,%.2lf\t%.2lf\t%s,%s,%.2lf\t%.2lf\t%.2lf\t%.2lg\t%.2lg\t%s\t%.2lf\t%s\t%.2lf\t%.5lf\t%s,%s,%.2lf\t%.2lf\t%.2lf\n",
             rank, L.barcodes[i].c_str(), (unsigned)r.n_snps, (int)L.cell_uniq_reads[i], type_name(r.type),
             sm(r.best_j), sm(r.best_k), alphas[r.best_a], r.best_llk, sm(r.next_j), sm(r.next_k), alphas[r.next_a], r.next_llk,
             r.best_llk - r.next_llk, r.best_pp, r.sng_pp, sm(r.sng_best), r.sng_best_llk, sm(r.sng_next), r.sng_next_llk,
             r.sng_only_pp, sm(r.dbl_best_j), sm(r.dbl_best_k), alphas[r.dbl_best_a < 0 ? 0 : r.dbl_best_a], r.dbl_best_llk,
             r.sng_best_llk - r.dbl_best_llk);
  }
}

void write_lmix(const std::string& path, const Loaded& L, const std::vector<pscl_fmx_cell>& cells, bool old_mode) {
  Out w(path, false);
  w.printf(old_mode ? "INT_ID\tBARCODE\tNSNPs\tNREADs\tDBL.LLK\tSNG.LLK\tLOG.BF\tBFpSNP\n"
                    : "INT_ID\tBARCODE\tNSNPs\tNREADs\tDBL.LLK\tSNG.LLK\tBF.SINGLET\tBF.SINGLET.PER.SNP\n");
  for (int32_t i = 0; i < L.n_cells; ++i) {
    const pscl_fmx_cell& r = cells[i];
    const double d = old_mode ? r.llk0 - r.llk2 : r.llk2 - r.llk0;  // cmd_cram_freemuxlet.cpp:163 vs cmd_cram_freemux2.cpp:161
    w.printf("%d\t%s\t%d\t%d\t%.2lf\t%.2lf\t%.2lf\t%.4lf\n", i, L.barcodes[i].c_str(), r.n_snps, r.n_reads, r.llk0, r.llk2, d, d / r.n_snps);
  }
}

void write_clust_samples(const std::string& path, const Loaded& L, const std::vector<pscl_fmx_cell>& cells) {
  Out w(path, true);
  w.printf("INT_ID\tBARCODE\tNUM.SNPS\tNUM.READS\tDROPLET.TYPE\tBEST.GUESS\tBEST.LLK\tNEXT.GUESS\tNEXT.LLK\tDIFF.LLK.BEST.NEXT\t"
           "BEST.POSTERIOR\tSNG.POSTERIOR\tSNG.BEST.GUESS\tSNG.BEST.LLK\tSNG.NEXT.GUESS\tSNG.NEXT.LLK\tSNG.ONLY.POSTERIOR\t"
           "DBL.BEST.GUESS\tDBL.BEST.LLK\tDIFF.LLK.SNG.DBL\n");
  for (int32_t i = 0; i < L.n_cells; ++i) {
    const pscl_fmx_cell& r = cells[i];
    w.printf("%d\t%s\t%d\t%d\t%s\t%d,%d\t%.2lf\t%d,%d\t%.2lf\t%.2lf\t%.5lf\t%.2lg\t%d\t%.2lf\t%d\t%.2lf\t%.5lf\t%d,%d\t%.2lf\t%.2lf\n",
             i, L.barcodes[i].c_str(), r.n_snps, r.n_reads, type_name(r.type), r.best_j, r.best_k, r.best_llk, r.next_j, r.next_k,
             r.next_llk, r.best_llk - r.next_llk, r.best_pp, r.sng_pp, r.sng_best, r.sng_best_llk, r.sng_next, r.sng_next_llk,
             r.sng_only_pp, r.dbl_best_j, r.dbl_best_k, r.dbl_best_llk, r.sng_best_llk - r.dbl_best_llk);
  }
}

void write_clust0_samples(const std::string& path, const Loaded& L, const std::vector<pscl_fmx_cell>& cells) {
  Out w(path, true);
  w.printf("INT_ID\tBARCODE\tCLUST0\n");
  for (int32_t i = 0; i < L.n_cells; ++i) w.printf("%d\t%s\t%d\n", i, L.barcodes[i].c_str(), cells[i].init_clust);
}

void write_clust_vcf(const std::string& path, const Loaded& L, int nS, const std::vector<double>& clust_gl,
                     const std::vector<int32_t>& clust_cnt, bool initial) {
  Out w(path, true);
  time_t now = std::time(NULL);
  tm* ltm = localtime(&now);
  w.printf("##fileformat=VCFv4.2\n");
  w.printf("##fileDate=%04d%02d%02d\n", 1970 + ltm->tm_year, 1 + ltm->tm_mon, ltm->tm_mday);  // (sic) cmd_cram_freemux2.cpp:610
  w.printf("##source=cramore-freemuxlet\n");
  for (const auto& c : L.rid2chr) w.printf("##contig=<ID=%s>\n", c.c_str());
  w.printf("##INFO=<ID=AF,Number=A,Type=Float,Description=\"Allele Frequency\">\n");
  w.printf("##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n");
  w.printf("##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Phred-scale Genotype Quality\">\n");
  w.printf("##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"Read Depth\">\n");
  w.printf("##FORMAT=<ID=AD,Number=R,Type=Integer,Description=\"Allelic Read Depth\">\n");
  w.printf("##FORMAT=<ID=PL,Number=G,Type=Integer,Description=\"Phred-scale genotype likelihood\">\n");
  w.printf("##FORMAT=<ID=GP,Number=G,Type=Float,Description=\"Posterior probability using pooled allele frequencies\">\n");
  w.printf("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT");
  for (int i = 0; i < nS; ++i) w.printf("\tCLUST%d", i);
  w.printf("\n");
  std::vector<uint8_t> observed(L.n_snps, 0);  // snps_observed (:279-287)
  for (int32_t s : L.pair_snp) observed[s] = 1;
  for (int32_t v = 0; v < L.n_snps; ++v) {
    if (!observed[v]) continue;
    const double af = L.af[v];
    w.printf("%s\t%d\t.\t%c\t%c\t.\tPASS\tAF=%.5lf\tGT:GQ:DP:AD:PL:GP", L.chrom[v].c_str(), L.pos[v], L.ref[v], L.alt[v], af);
    const double gps[3] = {(1. - af) * (1. - af), 2. * af * (1. - af), af * af};
    for (int i = 0; i < nS; ++i) {  // :633-655
      const double* gl = &clust_gl[((size_t)v * nS + i) * 9];
      const int32_t* n = &clust_cnt[((size_t)v * nS + i) * 3];
      double maxGL = gl[0];
      if (maxGL < gl[4]) maxGL = gl[4];
      if (maxGL < gl[8]) maxGL = gl[8];
      int32_t pls[3] = {(int32_t)(-10.0 * log10(gl[0] / maxGL)), (int32_t)(-10.0 * log10(gl[4] / maxGL)), (int32_t)(-10.0 * log10(gl[8] / maxGL))};
      double pps[3] = {gps[0] * (gl[0] / maxGL) + 1e-100, gps[1] * (gl[4] / maxGL) + 1e-100, gps[2] * (gl[8] / maxGL) + 1e-100};
      if (initial) { pps[0] = gps[0] * gl[0] / maxGL + 1e-100; pps[1] = gps[1] * gl[4] / maxGL + 1e-100; pps[2] = gps[2] * gl[8] / maxGL + 1e-100; }  // :327-329
      const double sumPP = pps[0] + pps[1] + pps[2];
      pps[0] /= sumPP; pps[1] /= sumPP; pps[2] /= sumPP;
      const int bestG = (pps[0] > pps[1]) ? (pps[0] > pps[2] ? 0 : 2) : (pps[1] > pps[2] ? 1 : 2);
      int gq = initial ? (int)(-0.1 * log10(1 - pps[bestG] + 1e-100)) : (int32_t)(-10 * log10(1.0 - pps[bestG] + 1e-100));  // :341 (sic) / :652
      if (gq > 255) gq = 255;
      w.printf("\t%d/%d:%d:%d:%d,%d:%d,%d,%d:%.3lg,%.3lg,%.3lg", bestG == 2 ? 1 : 0, bestG > 0 ? 1 : 0, gq, n[0], n[1], n[2], pls[0], pls[1],
               pls[2], pps[0], pps[1], pps[2]);
    }
    w.printf("\n");
  }
}

}  // namespace pscl_host
