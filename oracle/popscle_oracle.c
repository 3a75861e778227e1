/*
 * popscle_oracle.c — CPU restatement of the reference likelihood path (see popscle_oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into, imported by or called from the product.
 *
 * The loops keep the reference's structure (one log() per (j,k,n) term, same summation order,
 * FP64), only the containers differ: flat CSR arrays instead of nested std::map.  Compile with
 * -O2 -ffp-contract=off (no FMA fusion, like the reference's x86-64 -O3 build).
 */
#include "popscle_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MIN_NORM_GL 1e-6 /* sc_drop_seq.h:14 */

/* ---- PhredHelper.cpp:24-40 ---------------------------------------------------------------- */
static double g_err[256], g_mat[256];
static int g_phred_init = 0;
static void phred_init(void) {
  if (g_phred_init) return;
  for (int i = 0; i <= 255; i++) {
    g_err[i] = (i > 1) ? pow(0.1, i * 0.1) : 0.75; /* :30 */
    g_mat[i] = 1. - g_err[i];                      /* :32 */
  }
  g_phred_init = 1;
}
double orc_phred2err(int q) { phred_init(); return g_err[q & 255]; }
double orc_phred2mat(int q) { phred_init(); return g_mat[q & 255]; }

/* ---- sc_drop_seq.cpp:5-8 ------------------------------------------------------------------ */
double orc_log_add(double la, double lb) {
  if (la > lb) return la + log(1.0 + exp(lb - la));
  else return lb + log(1.0 + exp(la - lb));
}

/* ---- cmd_cram_demuxlet.cpp:655-725 --------------------------------------------------------- */
void orc_demux_pair_pg(const uint8_t* allele, const uint8_t* qual, int64_t n_reads, int nAlpha,
                       const double* gridAlpha, double* pGs) {
  phred_init();
  for (int i = 0; i < nAlpha * 9; ++i) pGs[i] = 1.0; /* :657 */
  for (int64_t r = 0; r < n_reads; ++r) {            /* :660 */
    uint8_t al = allele[r], bq = qual[r];
    if (al == 2) continue;                                 /* :664 */
    double pR = (al == 0) ? g_mat[bq] : g_err[bq] / 3.0;   /* :666 */
    double pA = (al == 1) ? g_mat[bq] : g_err[bq] / 3.0;   /* :667 */
    double maxpG = 0;
    for (int k = 0; k < nAlpha; ++k)
      for (int l = 0; l < 3; ++l)
        for (int m = 0; m < 3; ++m) {
          double p = 0.5 * l + (m - l) * 0.5 * gridAlpha[k]; /* :673 */
          double* pG = &pGs[k * 9 + l * 3 + m];
          *pG *= (pR * (1.0 - p) + pA * p);                  /* :685 */
          if (maxpG < *pG) maxpG = *pG;
        }
    for (int i = 0; i < nAlpha * 9; ++i) pGs[i] /= maxpG;    /* :692-699 */
  }
  double maxpG = 0;
  for (int i = 0; i < nAlpha * 9; ++i) { /* :704-715 */
    pGs[i] += 1e-10;
    if (maxpG < pGs[i]) maxpG = pGs[i];
  }
  for (int i = 0; i < nAlpha * 9; ++i) pGs[i] /= maxpG; /* :718-725 */
}

/* ---- cmd_cram_demuxlet.cpp:636-991, one cell ------------------------------------------------ */
static void demux_one_cell(const orc_pileup* plp, int32_t c, int nv, const double* gp,
                           const uint8_t* has_gp, int nAlpha, const double* gridAlpha,
                           double doublet_prior, double* llksAB, double* pGs, double* sumPs,
                           orc_demux_cell* o) {
  int j, k, l, m, n;
  memset(llksAB, 0, sizeof(double) * nv * nv * nAlpha); /* :643 */
  memset(o, 0, sizeof(*o));
  int64_t pb = plp->cell_ptr[c], pe = plp->cell_ptr[c + 1];
  o->n_snps = (int32_t)(pe - pb); /* cell_umis[i].size(), :996 */
  for (int64_t p = pb; p < pe; ++p) { /* :656 */
    int64_t rb = plp->pair_read_ptr[p], re = plp->pair_read_ptr[p + 1];
    orc_demux_pair_pg(plp->read_allele + rb, plp->read_qual + rb, re - rb, nAlpha, gridAlpha, pGs);
    int32_t isnp = plp->pair_snp[p];
    if (has_gp && !has_gp[isnp]) continue; /* :733 */
    const double* gps = gp + (int64_t)isnp * nv * 3;
    for (j = 0; j < nv; ++j) {   /* :734 */
      for (k = 0; k < nv; ++k) { /* :736 */
        for (n = 0; n < nAlpha; ++n) sumPs[n] = 0;
        for (l = 0; l < 3; ++l)
          for (m = 0; m < 3; ++m) {
            double p = gps[j * 3 + l] * gps[k * 3 + m]; /* :740 */
            for (n = 0; n < nAlpha; ++n) sumPs[n] += (p * pGs[n * 9 + l * 3 + m]); /* :742 */
          }
        for (n = 0; n < nAlpha; ++n) llksAB[j * nv * nAlpha + k * nAlpha + n] += log(sumPs[n]); /* :746 */
      }
    }
    /* llksA0 / llks00 (:749-774) have no live reader — not restated (SURVEY §8a D3') */
  }

  int32_t sBest = -1, sNext = -1, dBest1 = -1, dBest2 = -1, dNext1 = -1, dNext2 = -1,
          dblBestAlpha = -1, dblNextAlpha = -1; /* :788 */
  double sngBestLLK = -1e300, sngNextLLK = -1e300, dblBestLLK = -1e300, dblNextLLK = -1e300;
  double sumLLK = -1e-300, sngLLK = -1e-300; /* :791 (sic) */
  double bestPP = -1e300, sngPP, sngOnlyPP;
  double log_single_prior = log((1.0 - doublet_prior) / nv);                         /* :793 */
  double log_doublet_prior1 = log(doublet_prior / nv / (nv - 1.) / (nAlpha - 1.));   /* :794 */
  double log_doublet_prior2 = log(doublet_prior / nv / (nv - 1.) / (nAlpha - 1.) * 2); /* :795 */

  for (j = 0; j < nv; ++j) { /* :804-821 */
    sumLLK = orc_log_add(sumLLK, llksAB[j * nv * nAlpha] + log_single_prior);
    sngLLK = orc_log_add(sngLLK, llksAB[j * nv * nAlpha] + log_single_prior);
    for (k = 0; k < nv; ++k) {
      if (j == k) continue;
      for (n = 1; n < nAlpha; ++n) {
        if (gridAlpha[n] == 0.5) {
          if (k > j) continue;
          sumLLK = orc_log_add(sumLLK, llksAB[j * nv * nAlpha + k * nAlpha + n] + log_doublet_prior2);
        } else
          sumLLK = orc_log_add(sumLLK, llksAB[j * nv * nAlpha + k * nAlpha + n] + log_doublet_prior1);
      }
    }
  }
  for (j = 0; j < nv; ++j) { /* :827-837 */
    double x = llksAB[j * nv * nAlpha];
    if (sngBestLLK < x) { sngNextLLK = sngBestLLK; sNext = sBest; sBest = j; sngBestLLK = x; }
    else if (sngNextLLK < x) { sNext = j; sngNextLLK = x; }
  }
  for (j = 0; j < nv; ++j) /* :883-906 */
    for (k = 0; k < nv; ++k) {
      if (j == k) continue;
      for (n = 1; n < nAlpha; ++n) {
        double x = llksAB[j * nv * nAlpha + k * nAlpha + n];
        if (dblBestLLK < x) {
          dNext1 = dBest1; dNext2 = dBest2; dblNextAlpha = dblBestAlpha; dblNextLLK = dblBestLLK;
          dBest1 = j; dBest2 = k; dblBestAlpha = n; dblBestLLK = x;
        } else if (dblNextLLK < x) { dNext1 = j; dNext2 = k; dblNextAlpha = n; dblNextLLK = x; }
      }
    }

  int32_t jBest, kBest, jNext, kNext, alphaBest, alphaNext, type;
  double bestLLK, nextLLK;
  if (dblBestLLK > sngBestLLK + 2) { /* :925-946 */
    type = 1;
    bestPP = exp(dblBestLLK + ((gridAlpha[dblBestAlpha] == 0.5) ? log_doublet_prior2 : log_doublet_prior1) - sumLLK);
    jBest = dBest1; kBest = dBest2; bestLLK = dblBestLLK; alphaBest = dblBestAlpha;
    if (dblNextLLK > sngBestLLK + 2) { jNext = dNext1; kNext = dNext2; nextLLK = dblNextLLK; alphaNext = dblNextAlpha; }
    else { jNext = kNext = sBest; nextLLK = sngBestLLK; alphaNext = 0; }
  } else if (sngBestLLK > sngNextLLK + 2) { /* :947-967 */
    type = 0;
    bestPP = sngBestLLK + log_single_prior - sumLLK; /* no exp, :949 */
    jBest = kBest = sBest; bestLLK = sngBestLLK; alphaBest = 0;
    if (dblBestLLK > sngNextLLK + 2) { jNext = dBest1; kNext = dBest2; nextLLK = dblBestLLK; alphaNext = dblBestAlpha; }
    else { jNext = kNext = sNext; nextLLK = sngNextLLK; alphaNext = 0; }
  } else { /* :968-988 */
    type = 2;
    bestPP = sngBestLLK + log_single_prior - sumLLK;
    jBest = kBest = sBest; bestLLK = sngBestLLK; alphaBest = 0;
    if (dblBestLLK > sngNextLLK + 2) { jNext = dBest1; kNext = dBest2; nextLLK = dblBestLLK; alphaNext = dblBestAlpha; }
    else { jNext = kNext = sNext; nextLLK = sngNextLLK; alphaNext = 0; }
  }
  sngPP = exp(sngLLK - sumLLK);                             /* :990 */
  sngOnlyPP = exp(sngBestLLK + log_single_prior - sngLLK);  /* :991 */

  o->type = type;
  o->best_j = jBest; o->best_k = kBest; o->best_a = alphaBest;
  o->next_j = jNext; o->next_k = kNext; o->next_a = alphaNext;
  o->sng_best = sBest; o->sng_next = sNext;
  o->dbl_best_j = dBest1; o->dbl_best_k = dBest2; o->dbl_best_a = dblBestAlpha;
  o->dbl_next_j = dNext1; o->dbl_next_k = dNext2; o->dbl_next_a = dblNextAlpha;
  o->best_llk = bestLLK; o->next_llk = nextLLK; o->best_pp = bestPP; o->sng_pp = sngPP;
  o->sng_best_llk = sngBestLLK; o->sng_next_llk = sngNextLLK; o->sng_only_pp = sngOnlyPP;
  o->dbl_best_llk = dblBestLLK; o->dbl_next_llk = dblNextLLK; o->sum_llk = sumLLK; o->sng_llk = sngLLK;
}

int orc_demux(const orc_pileup* plp, int nv, const double* gp, const uint8_t* has_gp, int nAlpha,
              const double* gridAlpha, double doublet_prior, int cell_begin, int cell_end,
              orc_demux_cell* out, double* llk_grid, int n_threads) {
  if (!plp || nv < 1 || nAlpha < 1 || cell_begin < 0 || cell_end > plp->n_cells) return -1;
  phred_init();
  size_t gsz = (size_t)nv * nv * nAlpha;
  if (n_threads < 1) n_threads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
  {
    double* llksAB = (double*)malloc(sizeof(double) * gsz);
    double* pGs = (double*)malloc(sizeof(double) * nAlpha * 9);
    double* sumPs = (double*)malloc(sizeof(double) * nAlpha);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
    for (int c = cell_begin; c < cell_end; ++c) {
      demux_one_cell(plp, c, nv, gp, has_gp, nAlpha, gridAlpha, doublet_prior, llksAB, pGs, sumPs,
                     &out[c - cell_begin]);
      if (llk_grid) memcpy(llk_grid + (size_t)(c - cell_begin) * gsz, llksAB, sizeof(double) * gsz);
    }
    free(llksAB); free(pGs); free(sumPs);
  }
  return 0;
}

/* ---- sc_drop_seq.cpp:452-509 ---------------------------------------------------------------- */
double orc_fmx_pair_pileup(const uint8_t* allele, const uint8_t* qual, int64_t n_reads, double alpha,
                           double* gls, int32_t* counts) {
  phred_init();
  double tmp, logdenom = 0;
  for (int i = 0; i < 9; ++i) gls[i] = 1.0;
  counts[0] = counts[1] = counts[2] = 0;
  for (int64_t r = 0; r < n_reads; ++r) {
    uint8_t al = allele[r], bq = qual[r];
    ++counts[0];              /* :465 */
    if (al > 1) continue;     /* :467 */
    if (al == 0) ++counts[1]; else ++counts[2];
    double M = g_mat[bq], E4 = g_err[bq] / 4.;
    int ref = (al == 0);
    gls[0] *= (M * (ref ? 1.0 : 0.0) + E4);                                   /* :482 */
    gls[1] *= (M * (ref ? 1. - alpha / 2. : alpha / 2.) + E4);
    gls[2] *= (M * (ref ? 1.0 - alpha : alpha) + E4);
    gls[3] *= (M * (ref ? (1. + alpha) / 2. : (1. - alpha) / 2.) + E4);
    gls[4] *= (M * (ref ? .5 : .5) + E4);
    gls[5] *= (M * (ref ? (1. - alpha) / 2. : (1. + alpha) / 2.) + E4);
    gls[6] *= (M * (ref ? alpha : 1. - alpha) + E4);
    gls[7] *= (M * (ref ? alpha / 2. : 1. - alpha / 2.) + E4);
    gls[8] *= (M * (ref ? 0.0 : 1.0) + E4);                                   /* :490 */
    tmp = 0;
    for (int i = 0; i < 9; ++i) tmp += gls[i];
    for (int i = 0; i < 9; ++i) gls[i] /= tmp;
    logdenom += log(tmp); /* :495 */
  }
  for (int i = 0; i < 9; ++i)
    if (gls[i] < MIN_NORM_GL) gls[i] = MIN_NORM_GL; /* :498-501 */
  tmp = 0;
  for (int i = 0; i < 9; ++i) tmp += gls[i];
  for (int i = 0; i < 9; ++i) gls[i] /= tmp;
  logdenom += log(tmp);
  return logdenom;
}

/* ---- sc_drop_seq.h:77-101 ------------------------------------------------------------------- */
void orc_fmx_merge(double* gls, int32_t* cnt, double* logdenom, const double* ogls, const int32_t* ocnt,
                   double ologdenom) {
  cnt[0] += ocnt[0]; cnt[1] += ocnt[1]; cnt[2] += ocnt[2];
  *logdenom += ologdenom;
  for (int i = 0; i < 9; ++i) gls[i] *= ogls[i];
  double tmp = 0;
  for (int i = 0; i < 9; ++i) tmp += gls[i];
  *logdenom += log(tmp);
  for (int i = 0; i < 9; ++i) gls[i] /= tmp;
  for (int i = 0; i < 9; ++i)
    if (gls[i] < MIN_NORM_GL) gls[i] = MIN_NORM_GL;
  tmp = 0;
  for (int i = 0; i < 9; ++i) tmp += gls[i];
  *logdenom += log(tmp);
  for (int i = 0; i < 9; ++i) gls[i] /= tmp;
}

/* dense stand-in for std::map<int32_t,snp_droplet_pileup> per cluster: [V][nS] entries with a
 * presence flag (map membership matters only for the seeding distance, sc_drop_seq.cpp:551-552) */
typedef struct clust_tab {
  int V, nS;
  double* gls;     /* [V][nS][9] */
  int32_t* cnt;    /* [V][nS][3] */
  double* logden;  /* [V][nS]    */
  uint8_t* present;/* [V][nS]    */
} clust_tab;

static void tab_clear(clust_tab* t) {
  size_t n = (size_t)t->V * t->nS;
  for (size_t i = 0; i < n * 9; ++i) t->gls[i] = 1.0; /* default ctor, sc_drop_seq.h:72-75 */
  memset(t->cnt, 0, n * 3 * sizeof(int32_t));
  memset(t->logden, 0, n * sizeof(double));
  memset(t->present, 0, n);
}
static int tab_alloc(clust_tab* t, int V, int nS) {
  size_t n = (size_t)V * nS;
  t->V = V; t->nS = nS;
  t->gls = (double*)malloc(n * 9 * sizeof(double));
  t->cnt = (int32_t*)malloc(n * 3 * sizeof(int32_t));
  t->logden = (double*)malloc(n * sizeof(double));
  t->present = (uint8_t*)malloc(n);
  if (!t->gls || !t->cnt || !t->logden || !t->present) return -1;
  tab_clear(t);
  return 0;
}
static void tab_free(clust_tab* t) { free(t->gls); free(t->cnt); free(t->logden); free(t->present); }

static void tab_merge_cell(clust_tab* t, int j, const orc_pileup* plp, int32_t c, const double* pair_gl,
                           const int32_t* pair_cnt, const double* pair_ld) {
  for (int64_t p = plp->cell_ptr[c]; p < plp->cell_ptr[c + 1]; ++p) {
    size_t e = (size_t)plp->pair_snp[p] * t->nS + j;
    orc_fmx_merge(t->gls + e * 9, t->cnt + e * 3, t->logden + e, pair_gl + p * 9, pair_cnt + p * 3, pair_ld[p]);
    t->present[e] = 1;
  }
}

/* ---- cmd_cram_freemux2.cpp:383-456, cells [cb,ce) -------------------------------------------- */
static void estep_cell(const orc_pileup* plp, const double* pair_gl, const double* clust_gl, int nS,
                       double geno_error, int apply_err, int32_t c, double* llks, double* lks) {
  int npairs = nS * (nS + 1) / 2;
  double gp1s[3], gp2s[3], gp0s[3], sum1, sum2;
  for (int i = 0; i < npairs; ++i) llks[i] = 0;
  for (int64_t p = plp->cell_ptr[c]; p < plp->cell_ptr[c + 1]; ++p) {
    int32_t s = plp->pair_snp[p];
    double af = plp->snp_af[s];
    gp0s[0] = (1.0 - af) * (1.0 - af); gp0s[1] = 2 * af * (1.0 - af); gp0s[2] = af * af; /* :388-390 */
    const double* glis = pair_gl + p * 9;
    double lk;
    for (int j = 0; j < nS; ++j) {
      const double* g1 = clust_gl + ((size_t)s * nS + j) * 9;
      gp1s[0] = (1.0 - af) * (1.0 - af) * g1[0]; gp1s[1] = 2 * af * (1.0 - af) * g1[4]; gp1s[2] = af * af * g1[8]; /* :402-404 */
      sum1 = gp1s[0] + gp1s[1] + gp1s[2];
      gp1s[0] /= sum1; gp1s[1] /= sum1; gp1s[2] /= sum1;
      if (apply_err) { /* :410-415 */
        gp1s[0] = (1 - geno_error) * gp1s[0] + geno_error * gp0s[0];
        gp1s[1] = (1 - geno_error) * gp1s[1] + geno_error * gp0s[1];
        gp1s[2] = (1 - geno_error) * gp1s[2] + geno_error * gp0s[2];
      }
      for (int k = 0; k < j; ++k) {
        const double* g2 = clust_gl + ((size_t)s * nS + k) * 9;
        gp2s[0] = (1.0 - af) * (1.0 - af) * g2[0]; gp2s[1] = 2 * af * (1.0 - af) * g2[4]; gp2s[2] = af * af * g2[8];
        sum2 = gp2s[0] + gp2s[1] + gp2s[2];
        gp2s[0] /= sum2; gp2s[1] /= sum2; gp2s[2] /= sum2;
        if (apply_err) {
          gp2s[0] = (1 - geno_error) * gp2s[0] + geno_error * gp0s[0];
          gp2s[1] = (1 - geno_error) * gp2s[1] + geno_error * gp0s[1];
          gp2s[2] = (1 - geno_error) * gp2s[2] + geno_error * gp0s[2];
        }
        lk = 0;
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) lk += (glis[a * 3 + b] * gp1s[a] * gp2s[b]); /* :443 */
        lks[j * (j + 1) / 2 + k] = lk;
      }
      lk = 0;
      for (int a = 0; a < 3; ++a) lk += (glis[a * 3 + a] * gp1s[a]); /* :450 */
      lks[j * (j + 1) / 2 + j] = lk;
    }
    for (int i = 0; i < npairs; ++i) llks[i] += log(lks[i]); /* :454-455 */
  }
}

int orc_fmx_estep(const orc_pileup* plp, const double* pair_gl, const double* clust_gl, int nS,
                  double geno_error, int cell_begin, int cell_end, double* llk, int n_threads) {
  int npairs = nS * (nS + 1) / 2;
  if (n_threads < 1) n_threads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
  {
    double* lks = (double*)malloc(sizeof(double) * npairs);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8)
#endif
    for (int c = cell_begin; c < cell_end; ++c)
      estep_cell(plp, pair_gl, clust_gl, nS, geno_error, geno_error > 0, c,
                 llk + (size_t)(c - cell_begin) * npairs, lks);
    free(lks);
  }
  return 0;
}

typedef struct { double score; int32_t id; } score_id;
static const double* g_sort_scores;
/* sc_drop_comp_t (sc_drop_seq.h:190-198): score desc, ties larger id first */
static int cmp_drop(const void* a, const void* b) {
  int32_t lhs = *(const int32_t*)a, rhs = *(const int32_t*)b;
  double cmp = g_sort_scores[lhs] - g_sort_scores[rhs];
  if (cmp != 0) return cmp > 0 ? -1 : 1;
  return lhs > rhs ? -1 : (lhs < rhs ? 1 : 0);
}

static int fmx_run_impl(const orc_pileup* plp, const orc_fmx_opts* o, const int32_t* init_clust, orc_fmx_cell* out,
                        double* clust_gl_out, int32_t* clust_cnt_out, orc_fmx_result* res, double* pair_gl_out,
                        double* llk_last, int n_threads, double* clust_gl0, int32_t* clust_cnt0);
int orc_fmx_run(const orc_pileup* plp, const orc_fmx_opts* o, const int32_t* init_clust, orc_fmx_cell* out,
                double* clust_gl_out, int32_t* clust_cnt_out, orc_fmx_result* res, double* pair_gl_out,
                double* llk_last, int n_threads) {
  return fmx_run_impl(plp, o, init_clust, out, clust_gl_out, clust_cnt_out, res, pair_gl_out, llk_last, n_threads, NULL, NULL);
}
/* the same run; clust_gl0 / clust_cnt0 (nullable) receive the cluster pileups of the initial assignment — what --aux-files
 * writes to <out>.clust0.vcf.gz (cmd_cram_freemux2.cpp:277-347) */
int orc_fmx_run_aux(const orc_pileup* plp, const orc_fmx_opts* o, const int32_t* init_clust, orc_fmx_cell* out,
                    double* clust_gl_out, int32_t* clust_cnt_out, orc_fmx_result* res, int n_threads,
                    double* clust_gl0, int32_t* clust_cnt0) {
  return fmx_run_impl(plp, o, init_clust, out, clust_gl_out, clust_cnt_out, res, NULL, NULL, n_threads, clust_gl0, clust_cnt0);
}
static int fmx_run_impl(const orc_pileup* plp, const orc_fmx_opts* o, const int32_t* init_clust, orc_fmx_cell* out,
                        double* clust_gl_out, int32_t* clust_cnt_out, orc_fmx_result* res, double* pair_gl_out,
                        double* llk_last, int n_threads, double* clust_gl0, int32_t* clust_cnt0) {
  if (!plp || !o || !out || o->n_clusters < 1) return -1;
  phred_init();
  const int C = plp->n_cells, V = plp->n_snps, nS = o->n_clusters;
  const int64_t P = plp->n_pairs;
  const int npairs = nS * (nS + 1) / 2;
  double* pair_gl = (double*)malloc(sizeof(double) * 9 * (P > 0 ? P : 1));
  int32_t* pair_cnt = (int32_t*)malloc(sizeof(int32_t) * 3 * (P > 0 ? P : 1));
  double* pair_ld = (double*)malloc(sizeof(double) * (P > 0 ? P : 1));
  double* scores = (double*)malloc(sizeof(double) * (C > 0 ? C : 1));
  int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (C > 0 ? C : 1));
  int32_t* clusts = (int32_t*)malloc(sizeof(int32_t) * (C > 0 ? C : 1));
  int32_t* types = (int32_t*)malloc(sizeof(int32_t) * (C > 0 ? C : 1));
  double* llk = (double*)malloc(sizeof(double) * (size_t)(C > 0 ? C : 1) * npairs);
  clust_tab tab;
  if (tab_alloc(&tab, V, nS) != 0) return -2;
  memset(out, 0, sizeof(orc_fmx_cell) * C);

  /* stage 1 (cmd_cram_freemux2.cpp:117-163) */
  for (int32_t c = 0; c < C; ++c) {
    double llk0 = 0, llk2 = 0;
    int32_t nSNPs = 0, nReads = 0;
    for (int64_t p = plp->cell_ptr[c]; p < plp->cell_ptr[c + 1]; ++p) {
      int64_t rb = plp->pair_read_ptr[p], re = plp->pair_read_ptr[p + 1];
      double af = plp->snp_af[plp->pair_snp[p]];
      double* gls = pair_gl + p * 9;
      pair_ld[p] = orc_fmx_pair_pileup(plp->read_allele + rb, plp->read_qual + rb, re - rb, 0.5, gls, pair_cnt + p * 3);
      double lk0 = 0, lk2 = 0, gps[3];
      gps[0] = (1.0 - af) * (1.0 - af); gps[1] = 2.0 * af * (1.0 - af); gps[2] = af * af;
      for (int gi = 0; gi < 3; ++gi) {
        lk2 += (gls[gi * 3 + gi] * gps[gi]);
        for (int gj = 0; gj < 3; ++gj) lk0 += (gls[gi * 3 + gj] * gps[gi] * gps[gj]);
      }
      nReads += (int32_t)(re - rb); /* :150 */
      ++nSNPs;
      llk0 += log(lk0); llk2 += log(lk2);
    }
    scores[c] = llk2 - llk0; /* :159 */
    out[c].n_snps = nSNPs; out[c].n_reads = nReads; out[c].llk0 = llk0; out[c].llk2 = llk2;
  }
  if (pair_gl_out) memcpy(pair_gl_out, pair_gl, sizeof(double) * 9 * P);

  if (o->randomize_singlet_score) { /* :164-181: srand(seed or time), Fisher-Yates with libc rand() */
    srand(o->seed == 0 ? (unsigned)time(NULL) : (unsigned)o->seed);
    for (int32_t i = 0; i < C - 1; ++i) {
      int32_t j = i + rand() % (C - i);
      if (i < j) { double tmp = scores[j]; scores[j] = scores[i]; scores[i] = tmp; }
    }
  }
  for (int32_t c = 0; c < C; ++c) { clusts[c] = -1; types[c] = -1; order[c] = c; }
  if (o->mode_old && (!init_clust || o->iter_init > 0)) {
    /* freemuxlet-old: pairwise Bayes factors + votes (cmd_cram_freemuxlet.cpp:165-346).  The reference never seeds
     * rand(): glibc's default stream (seed 1), consumed nS values per visited droplet (:259, :312) and n-1 per
     * std::random_shuffle (:304; libstdc++: j = rand() % (i + 1) for i = 1..n-1, swap if i != j). */
    srand(1); /* what a fresh process starts with: every reference run draws the same numbers */
    g_sort_scores = scores;
    qsort(order, C, sizeof(int32_t), cmp_drop); /* :166-171 */
    /* SNP-major lists (snp_cell_plps, :113: std::map keyed by cell id -> ascending cell id) */
    int64_t* sptr = (int64_t*)calloc((size_t)V + 1, sizeof(int64_t));
    int64_t* spair = (int64_t*)malloc(sizeof(int64_t) * (P > 0 ? P : 1));
    int32_t* pcell = (int32_t*)malloc(sizeof(int32_t) * (P > 0 ? P : 1));
    for (int64_t p = 0; p < P; ++p) sptr[plp->pair_snp[p] + 1]++;
    for (int32_t v = 0; v < V; ++v) sptr[v + 1] += sptr[v];
    {
      int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * ((size_t)V + 1));
      memcpy(fill, sptr, sizeof(int64_t) * ((size_t)V + 1));
      for (int32_t c = 0; c < C; ++c)
        for (int64_t p = plp->cell_ptr[c]; p < plp->cell_ptr[c + 1]; ++p) { pcell[p] = c; spair[fill[plp->pair_snp[p]]++] = p; }
      free(fill);
    }
    /* dropDs[i][j], j < i (:174, :187-189): llk0 / llk2 only (nsnps / nread1 / nread2 feed --aux-files alone) */
    double* d0 = (double*)calloc((size_t)C * C, sizeof(double));
    double* d2 = (double*)calloc((size_t)C * C, sizeof(double));
    for (int32_t v = 0; v < V; ++v) { /* :190-222 */
      double af = plp->snp_af[v], gps[3];
      gps[0] = (1.0 - af) * (1.0 - af); gps[1] = 2.0 * af * (1.0 - af); gps[2] = af * af;
      for (int64_t a = sptr[v]; a < sptr[v + 1]; ++a) {
        const double* glis = pair_gl + spair[a] * 9;
        for (int64_t b = sptr[v]; b < a; ++b) {
          const double* gljs = pair_gl + spair[b] * 9;
          double lk0 = 0, lk2 = 0;
          for (int gi = 0; gi < 3; ++gi) {
            lk2 += (glis[gi * 3 + gi] * gljs[gi * 3 + gi] * gps[gi]);
            for (int gj = 0; gj < 3; ++gj) lk0 += (glis[gi * 3 + gi] * gljs[gj * 3 + gj] * gps[gi] * gps[gj]);
          }
          size_t e = (size_t)pcell[spair[a]] * C + pcell[spair[b]];
          d2[e] += log(lk2); d0[e] += log(lk0);
        }
      }
    }
#define DD(i, j) ((i) > (j) ? (size_t)(i) * C + (j) : (size_t)(j) * C + (i))
    double* votes = (double*)malloc(sizeof(double) * nS);
    if (init_clust) { /* :227-243 */
      for (int32_t c = 0; c < C; ++c) clusts[c] = init_clust[c] >= 0 ? init_clust[c] : -1;
    } else { /* :245-296 */
      for (int32_t i = 0; i < C; ++i) {
        int32_t si = order[i];
        if (i > C * o->frac_init_clust) continue; /* :248 */
        for (int j = 0; j < nS; ++j) votes[j] = rand() / (RAND_MAX + 1.) / 1000.; /* :258-260 */
        for (int32_t j = 0; j < i; ++j) {
          int32_t sj = order[j];
          size_t e = DD(si, sj);
          if (d0[e] - d2[e] > o->bf_thres) votes[clusts[sj]] -= 1.0;      /* :275-277 */
          else if (d2[e] - d0[e] > o->bf_thres) votes[clusts[sj]] += 1.0; /* :278-280 */
        }
        int elected = 0;
        double maxvote = votes[0];
        for (int j = 1; j < nS; ++j)
          if (maxvote < votes[j]) { elected = j; maxvote = votes[j]; } /* :282-289 */
        clusts[si] = elected;
      }
    }
    if (o->iter_init > 0) { /* :300-346: always ten sweeps */
      int32_t* orand = (int32_t*)malloc(sizeof(int32_t) * (C > 0 ? C : 1));
      for (int sweep = 0; sweep < 10; ++sweep) {
        for (int32_t i = 0; i < C; ++i) orand[i] = i;
        for (int32_t i = 1; i < C; ++i) { /* std::random_shuffle (libstdc++) */
          int32_t j = rand() % (i + 1);
          if (i != j) { int32_t t = orand[i]; orand[i] = orand[j]; orand[j] = t; }
        }
        for (int32_t i = 0; i < C; ++i) {
          int32_t si = orand[i];
          for (int j = 0; j < nS; ++j) votes[j] = rand() / (RAND_MAX + 1.) / 1000.;
          for (int32_t j = 0; j < C; ++j) {
            if (si != j) {
              size_t e = DD(si, j);
              double bf = d2[e] - d0[e];
              if (clusts[j] >= 0) {
                if (bf > o->bf_thres) ++votes[clusts[j]];
                else if (bf < 0 - o->bf_thres) --votes[clusts[j]];
              }
            }
          }
          int elected = 0;
          double maxvote = votes[0];
          for (int j = 1; j < nS; ++j)
            if (maxvote < votes[j]) { elected = j; maxvote = votes[j]; }
          if ((clusts[si] >= 0) || (o->keep_init_missing == 0)) clusts[si] = elected; /* :336-340 */
        }
      }
      free(orand);
    }
#undef DD
    free(votes); free(d0); free(d2); free(sptr); free(spair); free(pcell);
  } else if (init_clust) { /* :198-216 */
    for (int32_t c = 0; c < C; ++c)
      if (init_clust[c] >= 0) { clusts[c] = init_clust[c]; types[c] = 0; }
  } else { /* :184-189, :217-261 */
    g_sort_scores = scores;
    qsort(order, C, sizeof(int32_t), cmp_drop);
    for (int32_t i = 0; i < C; ++i) {
      int32_t si = order[i];
      if (i > C * o->frac_init_clust) continue;          /* :225 */
      if (scores[si] < o->singlet_score_thres) continue;  /* :226 */
      int maxClust = 0;
      double maxScore = 0;
      for (int j = 0; j < nS; ++j) { /* sc_drop_seq.cpp:544-578 */
        double d0 = 0, d2 = 0;
        for (int64_t p = plp->cell_ptr[si]; p < plp->cell_ptr[si + 1]; ++p) {
          int32_t s = plp->pair_snp[p];
          size_t e = (size_t)s * nS + j;
          if (!tab.present[e]) continue;
          double af = plp->snp_af[s], lk0 = 0, lk2 = 0, gps[3];
          gps[0] = (1.0 - af) * (1.0 - af); gps[1] = 2.0 * af * (1.0 - af); gps[2] = af * af;
          const double* glis = pair_gl + p * 9;
          const double* gljs = tab.gls + e * 9;
          for (int gi = 0; gi < 3; ++gi) {
            lk2 += (glis[gi * 3 + gi] * gljs[gi * 3 + gi] * gps[gi]);
            for (int gj = 0; gj < 3; ++gj) lk0 += (glis[gi * 3 + gi] * gljs[gj * 3 + gj] * gps[gi] * gps[gj]);
          }
          d2 += log(lk2); d0 += log(lk0);
        }
        double sc = d2 - d0;
        if (j == 0) { maxScore = sc; maxClust = 0; }       /* :235-236 */
        else if (sc > maxScore) { maxClust = j; maxScore = sc; } /* :238-241 */
      }
      clusts[si] = maxClust; types[si] = 0;
      tab_merge_cell(&tab, maxClust, plp, si, pair_gl, pair_cnt, pair_ld); /* :248-251 */
    }
  }
  for (int32_t c = 0; c < C; ++c) out[c].init_clust = clusts[c];

  /* cluster pileups (:277-288) */
  tab_clear(&tab);
  for (int32_t c = 0; c < C; ++c)
    if (clusts[c] >= 0) tab_merge_cell(&tab, clusts[c], plp, c, pair_gl, pair_cnt, pair_ld);
  if (clust_gl0) memcpy(clust_gl0, tab.gls, sizeof(double) * (size_t)V * nS * 9);   /* --aux-files, :291-347 */
  if (clust_cnt0) memcpy(clust_cnt0, tab.cnt, sizeof(int32_t) * (size_t)V * nS * 3);

  for (int32_t c = 0; c < C; ++c) { /* :350-370 */
    orc_fmx_cell* r = &out[c];
    r->best_j = r->best_k = r->next_j = r->next_k = -1;
    r->sng_best = r->sng_next = r->dbl_best_j = r->dbl_best_k = r->dbl_next_j = r->dbl_next_k = -1;
    r->best_llk = r->next_llk = r->sng_best_llk = r->sng_next_llk = r->dbl_best_llk = r->dbl_next_llk = -1e300;
    r->best_pp = r->sng_pp = r->sng_only_pp = r->sum_llk = -1e300;
  }

  double log_single_prior = log((1.0 - o->doublet_prior) / nS);            /* :379 */
  double log_double_prior = log(o->doublet_prior / nS / (nS - 1) * 2.0);   /* :380 */
  int iter, nsingle = 0, namb = 0, nchanged = 0;
  if (n_threads < 1) n_threads = 1;
  for (iter = 0; iter < o->max_iter; ++iter) {
    int apply_err = o->mode_old ? (o->geno_error > 0 && iter + 1 == o->max_iter) /* freemuxlet.cpp:485 */
                                : (o->geno_error > 0);                            /* freemux2.cpp:410 */
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
    {
      double* lks = (double*)malloc(sizeof(double) * npairs);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8)
#endif
      for (int32_t c = 0; c < C; ++c) {
        double* llks = llk + (size_t)c * npairs;
        estep_cell(plp, pair_gl, tab.gls, nS, o->geno_error, apply_err, c, llks, lks);
        int32_t sBest = -1, sNext = -1, dBest1 = -1, dBest2 = -1, dNext1 = -1, dNext2 = -1; /* :459-498 */
        double sngBestLLK = -1e300, sngNextLLK = -1e300, dblBestLLK = -1e300, dblNextLLK = -1e300;
        double sumLLK = -1e300, sngLLK = -1e300, tmpLLK;
        for (int j = 0; j < nS; ++j) {
          for (int k = 0; k < j; ++k) {
            tmpLLK = llks[j * (j + 1) / 2 + k];
            if (tmpLLK > dblBestLLK) { dNext1 = dBest1; dNext2 = dBest2; dblNextLLK = dblBestLLK; dBest1 = j; dBest2 = k; dblBestLLK = tmpLLK; }
            else if (tmpLLK > dblNextLLK) { dNext1 = j; dNext2 = k; dblNextLLK = tmpLLK; }
            sumLLK = orc_log_add(sumLLK, tmpLLK + log_double_prior);
          }
          tmpLLK = llks[j * (j + 1) / 2 + j];
          if (tmpLLK > sngBestLLK) { sNext = sBest; sngNextLLK = sngBestLLK; sBest = j; sngBestLLK = tmpLLK; }
          else if (tmpLLK > sngNextLLK) { sNext = j; sngNextLLK = tmpLLK; }
          sumLLK = orc_log_add(sumLLK, tmpLLK + log_single_prior);
          sngLLK = orc_log_add(sngLLK, tmpLLK + log_single_prior);
        }
        orc_fmx_cell* r = &out[c];
        r->sng_best = sBest; r->sng_best_llk = sngBestLLK; r->sng_next = sNext; r->sng_next_llk = sngNextLLK;
        r->dbl_best_j = dBest1; r->dbl_best_k = dBest2; r->dbl_best_llk = dblBestLLK;
        r->dbl_next_j = dNext1; r->dbl_next_k = dNext2; r->dbl_next_llk = dblNextLLK;
        r->sng_pp = exp(sngLLK - sumLLK);                               /* :510 */
        r->sng_only_pp = exp(sngBestLLK + log_single_prior - sngLLK);   /* :511 */
        r->sum_llk = sumLLK;
      }
      free(lks);
    }

    /* classify + M-step (:516-597) */
    tab_clear(&tab);
    nsingle = namb = nchanged = 0;
    for (int32_t c = 0; c < C; ++c) {
      orc_fmx_cell* r = &out[c];
      if (!o->mode_old) clusts[c] = -1; /* :520 (absent in freemuxlet.cpp) */
      if (r->dbl_best_llk > r->sng_best_llk + 2) {
        if (types[c] != 1) ++nchanged;
        types[c] = 1;
        r->best_pp = (r->dbl_best_llk + log_double_prior - r->sum_llk);
        r->best_j = r->dbl_best_j; r->best_k = r->dbl_best_k; r->best_llk = r->dbl_best_llk;
        if (r->dbl_next_llk > r->sng_best_llk + 2) { r->next_j = r->dbl_next_j; r->next_k = r->dbl_next_k; r->next_llk = r->dbl_next_llk; }
        else { r->next_j = r->next_k = r->sng_best; r->next_llk = r->sng_best_llk; }
      } else if (r->sng_best_llk > r->sng_next_llk + 2) {
        if ((types[c] != 0) || (r->best_j != r->sng_best) || (r->best_k != r->sng_best)) ++nchanged;
        types[c] = 0; ++nsingle;
        r->best_pp = (r->sng_best_llk + log_single_prior - r->sum_llk);
        r->best_j = r->best_k = r->sng_best; r->best_llk = r->sng_best_llk;
        if (!o->mode_old) clusts[c] = r->best_j; /* :553 */
        if (r->dbl_best_llk > r->sng_next_llk + 2) { r->next_j = r->dbl_best_j; r->next_k = r->dbl_best_k; r->next_llk = r->dbl_best_llk; }
        else { r->next_j = r->next_k = r->sng_next; r->next_llk = r->sng_next_llk; }
      } else {
        if (types[c] != 2) ++nchanged;
        types[c] = 2; ++namb;
        r->best_pp = (r->sng_best_llk + log_single_prior - r->sum_llk);
        r->best_j = r->best_k = r->sng_best; r->best_llk = r->sng_best_llk;
        if (r->dbl_best_llk > r->sng_next_llk + 2) { r->next_j = r->dbl_best_j; r->next_k = r->dbl_best_k; r->next_llk = r->dbl_next_llk; /* sic :578 */ }
        else { r->next_j = r->next_k = r->sng_next; r->next_llk = r->sng_next_llk; }
      }
      if ((r->best_j == r->best_k) && (types[c] == 0)) /* :592-594 */
        tab_merge_cell(&tab, r->best_j, plp, c, pair_gl, pair_cnt, pair_ld);
    }
    if (!o->mode_old && o->early_stop && nchanged == 0) { ++iter; break; } /* :601-604 */
  }
  for (int32_t c = 0; c < C; ++c) { out[c].type = types[c]; out[c].clust = clusts[c]; }
  if (res) {
    res->n_iter = iter; res->n_changed = nchanged; res->n_singlet = nsingle;
    res->n_doublet = C - nsingle - namb; res->n_ambiguous = namb;
  }
  if (clust_gl_out) memcpy(clust_gl_out, tab.gls, sizeof(double) * (size_t)V * nS * 9);
  if (clust_cnt_out) memcpy(clust_cnt_out, tab.cnt, sizeof(int32_t) * (size_t)V * nS * 3);
  if (llk_last) memcpy(llk_last, llk, sizeof(double) * (size_t)C * npairs);
  free(pair_gl); free(pair_cnt); free(pair_ld); free(scores); free(order); free(clusts); free(types); free(llk);
  tab_free(&tab);
  return 0;
}
