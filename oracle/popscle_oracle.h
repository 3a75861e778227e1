/*
 * popscle_oracle.h — CPU restatement of the popscle demuxlet / freemuxlet likelihood path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under popscle_b200/ (the product) may include, link or call
 * this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * do, and only as the checker / CPU baseline.
 *
 * Pinning status: the reference tree holds NO golden vectors, tests or fixtures for this path
 * (SURVEY.md §4, §8c).  The restatement is pinned against the reference's own translation
 * units compiled from /root/reference by oracle/build_ref.sh into oracle/_ref/ (the real
 * cmdCramDemuxlet / cmdCramFreemux2 / sc_drop_seq.cpp / PhredHelper.cpp code, linked against a
 * small htslib stand-in, oracle/htslib_standin/): tests/golden/ holds the files that binary wrote and
 * tests/test_golden.py requires this restatement to reproduce them byte for byte.
 *
 * Every function cites the reference file:line it follows.  All arithmetic is FP64 in the
 * reference's operation order (compile WITHOUT -ffast-math and with -ffp-contract=off).
 */
#ifndef POPSCLE_ORACLE_H
#define POPSCLE_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Flat pileup (same meaning as sc_dropseq_lib_t::cell_umis, sc_drop_seq.h:165): cell-major CSR,
 * pairs of a cell in ascending SNP id, reads of a pair in the reference's iteration order. */
typedef struct orc_pileup {
  int32_t n_cells, n_snps;
  int64_t n_pairs, n_reads;
  const int64_t* cell_ptr;      /* [C+1] */
  const int32_t* pair_snp;      /* [P]   */
  const int64_t* pair_read_ptr; /* [P+1] */
  const uint8_t* read_allele;   /* [N] 0 ref 1 alt 2 other */
  const uint8_t* read_qual;     /* [N] phred */
  const double* snp_af;         /* [V] */
} orc_pileup;

/* layout-identical to pscl_demux_cell (include/popscle_b200.h) so tests can memcmp ids */
typedef struct orc_demux_cell {
  int32_t n_snps, type;
  int32_t best_j, best_k, best_a, next_j, next_k, next_a;
  int32_t sng_best, sng_next;
  int32_t dbl_best_j, dbl_best_k, dbl_best_a, dbl_next_j, dbl_next_k, dbl_next_a;
  double best_llk, next_llk, best_pp, sng_pp, sng_best_llk, sng_next_llk, sng_only_pp;
  double dbl_best_llk, dbl_next_llk, sum_llk, sng_llk, reserved_;
} orc_demux_cell;

typedef struct orc_fmx_opts {
  int32_t n_clusters;
  double doublet_prior, geno_error;
  int32_t max_iter, early_stop;
  double frac_init_clust, singlet_score_thres;
  int32_t mode_old;
  int32_t randomize_singlet_score, seed; /* cmd_cram_freemux2.cpp:164-181 */
  /* freemuxlet-old's own seeding (cmd_cram_freemuxlet.cpp:184-346), used when mode_old and (no init_clust or iter_init > 0) */
  double bf_thres;            /* --bf-thres, default 5.41 (:22) */
  int32_t iter_init;          /* --iter-init, default 10: > 0 runs the ten vote-refinement sweeps (:300) */
  int32_t keep_init_missing;  /* --keep-init-missing (:336) */
} orc_fmx_opts;

/* layout-identical to pscl_fmx_cell */
typedef struct orc_fmx_cell {
  int32_t n_snps, n_reads, type, clust;
  int32_t best_j, best_k, next_j, next_k, sng_best, sng_next;
  int32_t dbl_best_j, dbl_best_k, dbl_next_j, dbl_next_k, init_clust, reserved_;
  double best_llk, next_llk, best_pp, sng_pp, sng_only_pp;
  double sng_best_llk, sng_next_llk, dbl_best_llk, dbl_next_llk, sum_llk, llk0, llk2;
} orc_fmx_cell;

typedef struct orc_fmx_result {
  int32_t n_iter, n_changed, n_singlet, n_doublet, n_ambiguous;
} orc_fmx_result;

/* PhredHelper.cpp:29-32 */
double orc_phred2err(int q);
double orc_phred2mat(int q);
/* sc_drop_seq.cpp:5-8 */
double orc_log_add(double la, double lb);

/* cmd_cram_demuxlet.cpp:655-725: pG[nalpha*9] of one (cell,SNP) pair */
void orc_demux_pair_pg(const uint8_t* allele, const uint8_t* qual, int64_t n_reads, int n_alpha,
                       const double* alphas, double* pG);

/* cmd_cram_demuxlet.cpp:636-991 for cells [cell_begin, cell_end).  gp = [V][nv][3] (after the
 * geno-error mixing of sc_drop_seq.cpp:285-315), has_gp NULL = all present.  llk_grid (nullable)
 * receives the FULL llksAB of every cell, [cell-cell_begin][j][k][n].  n_threads > 1 runs the
 * (independent) cells on an OpenMP team. */
int orc_demux(const orc_pileup* plp, int nv, const double* gp, const uint8_t* has_gp, int n_alpha,
              const double* alphas, double doublet_prior, int cell_begin, int cell_end,
              orc_demux_cell* out, double* llk_grid, int n_threads);

/* sc_drop_seq.cpp:452-509: gls[9], logdenom, counts[3]={nreads,nref,nalt} of one pair */
double orc_fmx_pair_pileup(const uint8_t* allele, const uint8_t* qual, int64_t n_reads, double alpha,
                           double* gls, int32_t* counts);
/* sc_drop_seq.h:77-101: a.merge(b); gls 9 doubles, counts 3 ints, logdenom in/out */
void orc_fmx_merge(double* gls_a, int32_t* cnt_a, double* logdenom_a, const double* gls_b,
                   const int32_t* cnt_b, double logdenom_b);

/* cmd_cram_freemux2.cpp:117-605 (mode_old: EM rules of cmd_cram_freemuxlet.cpp:456-653).
 * init_clust NULL = greedy seeding (:217-261).  Outputs nullable except `out`.
 * pair_gl (nullable) receives the stage-1 [P][9] GLs; llk_last (nullable) the last E-step's
 * [C][npairs] LLKs.  n_threads parallelises the per-cell E-step only. */
int orc_fmx_run(const orc_pileup* plp, const orc_fmx_opts* opts, const int32_t* init_clust,
                orc_fmx_cell* out, double* clust_gl, int32_t* clust_cnt, orc_fmx_result* res,
                double* pair_gl, double* llk_last, int n_threads);
/* the same run + the cluster pileups of the initial assignment (--aux-files, cmd_cram_freemux2.cpp:277-347) */
int orc_fmx_run_aux(const orc_pileup* plp, const orc_fmx_opts* opts, const int32_t* init_clust,
                    orc_fmx_cell* out, double* clust_gl, int32_t* clust_cnt, orc_fmx_result* res, int n_threads,
                    double* clust_gl0, int32_t* clust_cnt0);

/* one E-step over cells [cell_begin,cell_end) given a dense cluster table [V][nS][9]
 * (cmd_cram_freemux2.cpp:383-456) — used as the timed CPU baseline unit for freemuxlet */
int orc_fmx_estep(const orc_pileup* plp, const double* pair_gl, const double* clust_gl, int nS,
                  double geno_error, int cell_begin, int cell_end, double* llk, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
