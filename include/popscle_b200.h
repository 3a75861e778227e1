/*
 * popscle_b200.h — C ABI of the B200-native demuxlet / freemuxlet genotype-likelihood engine.
 *
 * The reference (statgen/popscle) has no plugin / FFI layer: each command is one function
 * `int32_t cmdXxx(int32_t argc, char** argv)` (reference commands.h:33-38, cramore.cpp:43-45).
 * The seam this library replaces is the block of each command body that sits AFTER
 * `sc_dropseq_lib_t::load_from_plp` has returned (cmd_cram_demuxlet.cpp:122,
 * cmd_cram_freemux2.cpp:87) and BEFORE the `hprintf` row writers
 * (cmd_cram_demuxlet.cpp:993, cmd_cram_freemux2.cpp:608-665).  Each entry point below cites the
 * reference lines it stands in for.  The state that crosses the seam in the reference is
 * `sc_dropseq_lib_t` (sc_drop_seq.h:130-184) — nested std::maps — which is flattened here to a
 * cell-major CSR / SoA pileup image.
 *
 * Conventions
 *  - plain C, no C++/torch types; all buffers are caller-owned HOST memory unless a name ends
 *    in `_dev` (then it is a CUDA device pointer on the context's GPU);
 *  - every function returns PSCL_OK (0) or a negative pscl_status; the message for the last
 *    failure of a context is available from pscl_last_error().  Nothing throws across the ABI
 *    (the reference's error() prints "FATAL ERROR" and throws, Error.cpp:29-42; the host wrapper
 *    re-creates that behaviour from the status code);
 *  - there is NO CPU fallback: without a CUDA device pscl_create() fails with PSCL_ENODEV.
 */
#ifndef POPSCLE_B200_H
#define POPSCLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSCL_ABI_VERSION 7

typedef enum pscl_status {
  PSCL_OK = 0,
  PSCL_EINVAL = -1,  /* bad argument (message says which)                           */
  PSCL_ENODEV = -2,  /* no usable CUDA device / wrong architecture                  */
  PSCL_ECUDA = -3,   /* CUDA runtime error (message carries cudaGetErrorString)     */
  PSCL_ENOMEM = -4,  /* device or host allocation failed                            */
  PSCL_ESTATE = -5   /* call sequence error (e.g. EM step before init)              */
} pscl_status;

/* droplet types, as printed in DROPLET.TYPE (cmd_cram_demuxlet.cpp:925-988,
 * cmd_cram_freemux2.cpp:663: types 0=SNG 1=DBL 2=AMB) */
enum { PSCL_SNG = 0, PSCL_DBL = 1, PSCL_AMB = 2 };

/* ------------------------------------------------------------------------------------------
 * Pileup image (replaces sc_dropseq_lib_t::cell_umis / snp_umis / snps, sc_drop_seq.h:158-173).
 * Cell-major CSR.  A "pair" is one (cell, SNP) with >= 1 base-call, i.e. one entry of
 * cell_umis[cell]; pairs of a cell are sorted by ascending SNP id (std::map order,
 * cmd_cram_demuxlet.cpp:656).  A "read" is one base-call that passed --min-BQ, already
 * capped at --cap-BQ (sc_drop_seq.cpp:361-369).
 * ------------------------------------------------------------------------------------------ */
typedef struct pscl_pileup {
  int32_t n_cells;              /* C  = scl.nbcs                                               */
  int32_t n_snps;               /* V  = scl.nsnps                                              */
  int64_t n_pairs;              /* P                                                           */
  int64_t n_reads;              /* N                                                           */
  const int64_t* cell_ptr;      /* [C+1] pair range of each cell                               */
  const int32_t* pair_snp;      /* [P]   SNP id of the pair                                    */
  const int64_t* pair_read_ptr; /* [P+1] read range of each pair                               */
  const uint8_t* read_allele;   /* [N]   0 = REF, 1 = ALT, 2 = other (sc_drop_seq.cpp:58)      */
  const uint8_t* read_qual;     /* [N]   phred base quality, 0..93                             */
  const double* snp_af;         /* [V]   AF column of .var.gz (freemuxlet prior); may be NULL
                                         for demuxlet                                          */
  /* Compact alternatives (ABI 2; NULL = not given).  A host that builds them halves the bytes that
   * cross PCIe per pileup (13 -> 9 B per pair + 2 -> 1 B per base-call); when one is given the
   * wide array it replaces may be NULL. */
  const uint32_t* pair_read_ptr32; /* [P+1] the same offsets as pair_read_ptr (n_reads < 2^32)  */
  const uint8_t* read_aq;          /* [N]   allele << 6 | qual, qual <= 63                      */
  /* ABI 3: the pair arrays as small deltas (decoded on the device by a per-cell scan and a global
   * scan): 3 B per pair instead of 8.  pair_snp_delta16[p] = pair_snp[p] - pair_snp[p-1] inside a
   * cell (0 for the first pair of a cell, whose SNP id is cell_first_snp[cell]); every delta must be
   * < 65536 and every pair must have < 256 base-calls, else the host keeps the wider arrays. */
  const int32_t* cell_first_snp;     /* [C]   SNP id of the first pair of each cell (any value for empty cells) */
  const uint16_t* pair_snp_delta16;  /* [P]                                                      */
  const uint8_t* pair_nreads8;       /* [P]   base-calls of each pair                            */
  /* ABI 6: the same two pair arrays in 1.25 B per pair (5 B per pair over PCIe became 1.3).  SNP gaps are small
   * (a cell covers ~2 % of the SNPs in ascending order) and so are the counts (1.3 base-calls per pair on average), so both
   * are stored short with the rare large values on the side, in pair order:
   *   pair_snp_delta8[p]  = the gap to the previous pair's SNP id if it is < 255, else 255 and the gap itself is the next
   *                         unread entry of snp_gap_big; 0 (ignored) at a cell's first pair;
   *   cell_gap_big_ptr[c] = how many entries of snp_gap_big belong to the cells before c;
   *   pair_nreads2        = two bits per pair, pair p in bits 2*(p%4).. of byte p/4: 1..3 base-calls, or 0 and the count
   *                         (4..255) is the next unread entry of nreads_big;
   *   nreads_big_ptr[k]   = how many entries of nreads_big belong to the pairs before 1024*k.
   * Given together with cell_first_snp (ABI 3); pair_snp_delta8 + pair_nreads2 replace pair_snp_delta16 + pair_nreads8. */
  const uint8_t* pair_snp_delta8;    /* [P]                                                      */
  const uint32_t* snp_gap_big;       /* [n_gap_big]                                              */
  const int64_t* cell_gap_big_ptr;   /* [C+1]                                                    */
  const uint8_t* pair_nreads2;       /* [(P+3)/4]                                                */
  const uint8_t* nreads_big;         /* [n_nreads_big]                                           */
  const int64_t* nreads_big_ptr;     /* [P/1024 + 2]                                             */
  int64_t n_gap_big, n_nreads_big;
  /* ABI 6: the base-calls as indices into a palette of the distinct allele<<6|qual bytes — after --min-BQ / --cap-BQ there
   * are few of them (24 with the commands' defaults 13 / 20: 5 bits per base-call instead of 8).  read_bits in {4, 5, 6};
   * base-call r occupies bits [r*read_bits, (r+1)*read_bits) of the little-endian bit string read_packed; replaces read_aq. */
  const uint8_t* read_packed;        /* [(N*read_bits + 7)/8 + 1] (one byte of slack)            */
  const uint8_t* read_palette;       /* [1 << read_bits] allele<<6|qual of each index            */
  int32_t read_bits;                 /* 0 = not given                                            */
  int32_t reserved_;
  /* ABI 7 (optional, NULL = not given): the first base-call of every cell, cell_read_ptr[c] = pair_read_ptr[cell_ptr[c]]
   * (cell_read_ptr[C] = n_reads).  A host that sends base-call COUNTS (ABI 3 / 6) instead of offsets loses nothing by adding
   * these 8 bytes per cell, and they let pscl_demux_run cut the counts and base-calls at the same cells as the SNP gaps:
   * every slice of the pileup then lands whole and is decoded and scored while the next one crosses PCIe. */
  const int64_t* cell_read_ptr;      /* [C+1]                                                    */
} pscl_pileup;

/* Genotype table (replaces sc_snp_t::gps, sc_drop_seq.h:29-37, filled at
 * sc_drop_seq.cpp:287-315 — i.e. AFTER the geno-error mixing). */
typedef struct pscl_geno {
  int32_t n_samples;     /* nv = vr.get_nsamples()                                              */
  const double* gp;      /* [V * nv * 3] P(genotype = 0/1/2) per SNP, sample; rows of SNPs with
                            has_gp == 0 are ignored.  NULL when one of the raw forms below is given */
  const uint8_t* has_gp; /* [V] 1 if snps[v].gps != NULL (cmd_cram_demuxlet.cpp:733); NULL = all 1 */
  /* ABI 4, optional: the reader's posteriors BEFORE the geno-error mixing, which the library then
   * does on the device exactly as sc_drop_seq.cpp:287-315 does (per-SNP average started at 1e-10 and
   * accumulated in sample order, gps = (1-err)*gp + err*avg in double, err clamped to [0, 0.999]).
   * gp_f32: [V * nv * 3] floats as vr.get_posterior_probability() returns them (any --field), half
   * the bytes of `gp`.  gt8: [V * nv] hard calls 0/1/2 (--field GT without missing calls: the
   * one-hot rows of bcf_filtered_reader.cpp:385-409), 1/24 of the bytes.  The error rate is
   * `geno_err_snp[v]` when that is given (--geno-error-coeff with an R2 INFO field), else `geno_err`. */
  const float* gp_f32;
  const uint8_t* gt8;
  const double* geno_err_snp;
  double geno_err;
} pscl_geno;

/* ------------------------------------------------------------------------------------------
 * demuxlet — replaces cmd_cram_demuxlet.cpp:636-991 (per-barcode loop: pG fold :655-725,
 * pair grid :733-747, priors / logAdd sums :788-821, best/next scans :827-906,
 * SNG/DBL/AMB decision :925-991).  One record per cell = the numeric content of one `.best`
 * row (:993-1013).  Indices are sample indices (host maps them to vr.get_sample_id_at) and
 * alpha-grid indices (host prints gridAlpha[idx]).
 * ------------------------------------------------------------------------------------------ */
typedef struct pscl_demux_opts {
  int32_t n_alpha;        /* |gridAlpha|; must be >= 2 (reference divides by nAlpha-1, :794)    */
  const double* alphas;   /* gridAlpha; alphas[0] is taken as the singlet plane (:806,:828)     */
  double doublet_prior;   /* --doublet-prior, default 0.5 (:32)                                 */
} pscl_demux_opts;

typedef struct pscl_demux_cell {
  int32_t n_snps;                               /* cell_umis[i].size() (:996)                   */
  int32_t type;                                 /* PSCL_SNG / PSCL_DBL / PSCL_AMB               */
  int32_t best_j, best_k, best_a;               /* BEST.GUESS  (:999)                           */
  int32_t next_j, next_k, next_a;               /* NEXT.GUESS  (:1001)                          */
  int32_t sng_best, sng_next;                   /* SNG.BEST.GUESS / SNG.NEXT.GUESS              */
  int32_t dbl_best_j, dbl_best_k, dbl_best_a;   /* DBL.BEST.GUESS (:1011)                       */
  int32_t dbl_next_j, dbl_next_k, dbl_next_a;   /* second-best doublet (used by :933-939)       */
  double best_llk, next_llk;                    /* BEST.LLK, NEXT.LLK                           */
  double best_pp;                               /* BEST.POSTERIOR — exp() only in the DBL branch
                                                   (:927 vs :949,:970), reproduced as is        */
  double sng_pp;                                /* SNG.POSTERIOR  = exp(sngLLK - sumLLK) (:990) */
  double sng_best_llk, sng_next_llk;
  double sng_only_pp;                           /* SNG.ONLY.POSTERIOR (:991)                    */
  double dbl_best_llk, dbl_next_llk;
  double sum_llk, sng_llk;                      /* the two logAdd sums (:791, init -1e-300 sic) */
  double reserved_;                             /* pad to 160 bytes                             */
} pscl_demux_cell;

/* ------------------------------------------------------------------------------------------
 * freemuxlet — replaces cmd_cram_freemux2.cpp:117-163 (stage 1), :184-261 (sort + greedy
 * seeding), :277-288 (cluster pileup build), :373-605 (EM) ; `mode_old` selects the EM variant
 * of cmd_cram_freemuxlet.cpp:456-653 (geno_error only on the last iteration, no early stop,
 * cluster ids not reset) and its pairwise/vote seeding (:165-346; see bf_thres / iter_init below).
 * ------------------------------------------------------------------------------------------ */
typedef struct pscl_fmx_opts {
  int32_t n_clusters;         /* --nsample                                                      */
  double doublet_prior;       /* default 0.5  (cmd_cram_freemux2.cpp:19)                        */
  double geno_error;          /* default 0.1  (:20)                                             */
  int32_t max_iter;           /* reference hard-codes 10 (:373)                                 */
  int32_t early_stop;         /* 1 = break when nchanged == 0 (:601)                            */
  double frac_init_clust;     /* default 1.0 (:28)                                              */
  double singlet_score_thres; /* default -1e300 (:26)                                           */
  int32_t mode_old;           /* 0 = freemux2 (popscle freemuxlet), 1 = freemuxlet-old EM rules */
  /* ABI 5: --randomize-singlet-score (cmd_cram_freemux2.cpp:164-181): the singlet scores are shuffled
   * with libc rand() (Fisher-Yates, j = i + rand() % (n - i)) after srand(seed ? seed : time(0)),
   * before the droplets are sorted for the greedy seeding.  0 = off.                             */
  int32_t randomize_singlet_score;
  int32_t seed;               /* --seed                                                         */
  /* ABI 6: freemuxlet-old's own seeding (cmd_cram_freemuxlet.cpp:165-346), used when mode_old is set and there is no
   * init_clust or iter_init > 0: the pairwise Bayes-factor matrix over the shared SNPs of every two droplets (on the
   * device), votes of the already clustered droplets (:245-296) and, when iter_init > 0, ten refinement sweeps in
   * std::random_shuffle order (:300-346) — on the un-seeded libc rand() stream, as the reference.                   */
  double bf_thres;            /* --bf-thres, default 5.41 (cmd_cram_freemuxlet.cpp:22)          */
  int32_t iter_init;          /* --iter-init, default 10; 0 skips the sweeps                    */
  int32_t keep_init_missing;  /* --keep-init-missing (:336)                                     */
} pscl_fmx_opts;

typedef struct pscl_fmx_cell {
  int32_t n_snps, n_reads;              /* NUM.SNPS, NUM.READS (incl. allele==2, :150)          */
  int32_t type;                         /* 0 SNG 1 DBL 2 AMB                                     */
  int32_t clust;                        /* clusts[i] after the last classify (-1 unless SNG)     */
  int32_t best_j, best_k, next_j, next_k;
  int32_t sng_best, sng_next;
  int32_t dbl_best_j, dbl_best_k, dbl_next_j, dbl_next_k;
  int32_t init_clust;                   /* cluster after seeding (clust0)                        */
  int32_t reserved_;
  double best_llk, next_llk;
  double best_pp;                       /* stored as a LOG in all branches (:526,:549,:571)      */
  double sng_pp, sng_only_pp;
  double sng_best_llk, sng_next_llk, dbl_best_llk, dbl_next_llk;
  double sum_llk;
  double llk0, llk2;                    /* stage-1 DBL.LLK / SNG.LLK of .lmix (:161)             */
} pscl_fmx_cell;

typedef struct pscl_fmx_result {
  int32_t n_iter;      /* EM iterations executed                                                */
  int32_t n_changed;   /* nchanged of the last iteration                                        */
  int32_t n_singlet, n_doublet, n_ambiguous;
} pscl_fmx_result;

typedef struct pscl_ctx pscl_ctx;
typedef struct pscl_plp pscl_plp; /* device-resident pileup image */

/* ---- context ---- */
int pscl_abi_version(void);
/* Opens CUDA device `device` (ordinal).  Fails with PSCL_ENODEV when there is no sm_100 GPU. */
int pscl_create(int device, pscl_ctx** out, char* err, size_t errlen);
void pscl_destroy(pscl_ctx* ctx);
const char* pscl_last_error(const pscl_ctx* ctx);
/* cudaStream_t the context launches on (as void*), for callers that time with CUDA events
 * or order their own work (e.g. a torch.distributed all-reduce) after ours. */
void* pscl_stream(pscl_ctx* ctx);
int pscl_sync(pscl_ctx* ctx);

/* ---- pileup upload: host CSR -> packed device image (1 B/read, 4+4 B/pair) ----
 * Uses the compact arrays of pscl_pileup when present (no repacking kernels, fewer H2D bytes). */
int pscl_plp_upload(pscl_ctx* ctx, const pscl_pileup* host, pscl_plp** out);
void pscl_plp_free(pscl_ctx* ctx, pscl_plp* plp);

/* ---- demuxlet ---- */
/* Uploads the genotype table (replicated per GPU).  Replaces the previous one. */
int pscl_demux_set_geno(pscl_ctx* ctx, const pscl_geno* geno, int32_t n_snps);
/* Scores cells [cell_begin, cell_end) of the resident pileup.  Results stay on the device;
 * kernels are enqueued on pscl_stream() and the call returns without synchronising. */
int pscl_demux_score(pscl_ctx* ctx, const pscl_plp* plp, const pscl_demux_opts* opts,
                     int32_t cell_begin, int32_t cell_end);
/* Copies the per-cell records of the last pscl_demux_score() to `out[cell_end-cell_begin]`
 * (synchronises).  `llk_grid`, if not NULL, receives the live part of llksAB as
 * [cell][j][k][n] doubles (n_cells * nv * nv * n_alpha): singlets at [j][0][0], doublets at
 * [j][k][n>=1] (j != k); entries the reference never reads (:806,:883-906) are NaN.
 * Requires pscl_demux_score to have been called with the grid enabled (see below). */
int pscl_demux_fetch(pscl_ctx* ctx, pscl_demux_cell* out, double* llk_grid);
/* Debug switch: keep the per-cell LLK grid on the device so pscl_demux_fetch can return it. */
int pscl_demux_keep_grid(pscl_ctx* ctx, int enable);
/* One-call path used by the CLI hosts: genotype table + pileup upload + score + fetch; returns with the
 * records in `out` (and the grid in `llk_grid`, if not NULL).  When `host` carries the compact arrays (ABI 3 / 6)
 * and the shape is the default kernel's (alpha grid {0, 0.5}, <= 8 samples, no grid requested), a pileup of
 * >= 4 M pairs is run PIPELINED: every copy of the call is queued on a second stream at once (small arrays,
 * counts, base-calls, then the SNP gaps in slices of whole cells), the genotype tables are built and the
 * counts / base-calls decoded under the copies, every slice's gaps are decoded as they land and the cells
 * scored in groups of slices, and the records come back with the image's validity flag in one read (one host
 * drain per call).  Page-locked host arrays (cudaHostAlloc / cudaHostRegister), `out` included, make the
 * overlap real; pageable ones give the same records without it.  The records are the bytes of the one-shot
 * run whatever the slicing.  Environment (none needed): PSCL_SLICES=n / PSCL_GROUPS=g force and shape the
 * pipeline at any size (1 = off), PSCL_SLICE_READS=1 / PSCL_SLICE_FULL=1 slice the base-calls / counts and base-calls too (need ABI 7's
 * cell_read_ptr or the offsets), PSCL_STAGES=n selects the older form instead (one scoring launch whose warps
 * decode the gaps and wait on a flag word per slice), PSCL_TRACE=1 prints the wall-clock of every phase and
 * PSCL_TIMELINE=1 the device time stamps of the call on stderr. */
int pscl_demux_run(pscl_ctx* ctx, const pscl_pileup* host, const pscl_geno* geno,
                   const pscl_demux_opts* opts, pscl_demux_cell* out, double* llk_grid);
/* Device time (ms, CUDA events on pscl_stream) of the kernels of the last pscl_demux_score. */
int pscl_demux_last_kernel_ms(pscl_ctx* ctx, float* ms_main, float* ms_total);
/* Number of kernel launches issued by this context so far (bench.py's gpu_launches). */
int64_t pscl_launch_count(const pscl_ctx* ctx);

/* ---- freemuxlet ---- */
/* Whole run on one GPU: stage 1, seeding (or init_clust[C], -1 = unassigned; NULL = greedy
 * seeding), cluster build, EM.  `clust_gl` (nullable) receives [V][n_clusters][9] cluster
 * genotype likelihoods (absent entries = 1.0, as operator[] default-constructs them,
 * cmd_cram_freemux2.cpp:634); `clust_cnt` (nullable) [V][n_clusters][3] = nreads,nref,nalt. */
int pscl_fmx_run(pscl_ctx* ctx, const pscl_pileup* host, const pscl_fmx_opts* opts,
                 const int32_t* init_clust, pscl_fmx_cell* out, double* clust_gl,
                 int32_t* clust_cnt, pscl_fmx_result* res);
/* ABI 7: the same run, which also hands back the cluster pileups of the INITIAL assignment (after the seeding, before the
 * first E-step) in `clust_gl0` / `clust_cnt0` (nullable, shaped like clust_gl / clust_cnt): what `--aux-files` writes to
 * <out>.clust0.vcf.gz (cmd_cram_freemux2.cpp:277-347); the initial clusters of <out>.clust0.samples.gz (:265-274) are
 * pscl_fmx_cell.init_clust. */
int pscl_fmx_run_aux(pscl_ctx* ctx, const pscl_pileup* host, const pscl_fmx_opts* opts,
                     const int32_t* init_clust, pscl_fmx_cell* out, double* clust_gl,
                     int32_t* clust_cnt, pscl_fmx_result* res, double* clust_gl0, int32_t* clust_cnt0);

/* Step-level API for SNP-sharded multi-GPU EM: each rank uploads the pairs of its
 * SNP range (cells keep global ids), and the caller all-reduces the per-cell partial sums between
 * the steps.  All `_dev` pointers are device memory on the context's GPU; kernels are enqueued on
 * pscl_stream().  Sequence:
 *   init -> stage1 -> [all-reduce stage1_dev] -> seed -> mstep(clust_dev) ->
 *   repeat { estep -> [all-reduce llk_dev] -> classify -> mstep(NULL) } -> fetch            */
int pscl_fmx_init(pscl_ctx* ctx, const pscl_plp* plp, const pscl_fmx_opts* opts);
/* Stage 1 (cmd_cram_freemux2.cpp:117-159): per-pair 9-GL (kept on the device) and the per-cell
 * partial sums of the local pairs, written to stage1_dev[4*C] doubles as four planes
 * llk0 | llk2 | nsnps | nreads. */
int pscl_fmx_stage1(pscl_ctx* ctx, double* stage1_dev);
/* Consumes the (all-reduced) stage-1 planes: fills the per-cell records, then either takes the
 * initial clusters from init_clust_dev[C] (-1 = unassigned; --init-cluster, :198-216) or, when
 * it is NULL, runs the greedy seeding (:184-189, :223-260) on the local pairs — only meaningful
 * when this rank holds all pairs.  clust_dev[C] receives the initial cluster of every cell. */
int pscl_fmx_seed(pscl_ctx* ctx, const double* stage1_dev, const int32_t* init_clust_dev,
                  int32_t* clust_dev);
/* (Re)builds the cluster pileups of the local SNPs by ordered merges (:277-288, :590-596).
 * clust_dev[C] = cluster each cell is merged into (-1 = none); NULL = the singlet membership the
 * last pscl_fmx_classify produced. */
int pscl_fmx_mstep(pscl_ctx* ctx, const int32_t* clust_dev);
/* E-step partial LLKs of the local SNPs (:383-456) into llk_dev[C * npairs] (overwritten),
 * pair (j,k<=j) at index j(j+1)/2+k as in the reference. */
int pscl_fmx_estep(pscl_ctx* ctx, int32_t iter, double* llk_dev);
/* Per-cell epilogue + classification from the (all-reduced) llk_dev (:458-584); updates the
 * cell records, clust_dev and the membership of the next M-step; *res (nullable) receives
 * nchanged etc.  Synchronises. */
int pscl_fmx_classify(pscl_ctx* ctx, const double* llk_dev, int32_t* clust_dev,
                      pscl_fmx_result* res);
/* Per-cell records (nullable) and, when requested, the full cluster pileups of the current
 * membership (layout as in pscl_fmx_run).  Synchronises. */
int pscl_fmx_fetch(pscl_ctx* ctx, pscl_fmx_cell* out, double* clust_gl, int32_t* clust_cnt);
/* Device time (ms) of the last pscl_fmx_estep or pscl_fmx_mstep (CUDA events on pscl_stream). */
int pscl_fmx_last_kernel_ms(pscl_ctx* ctx, float* ms);

/* ------------------------------------------------------------------------------------------
 * Several GPUs in one process (ABI 6).  The reference is single-threaded and its only
 * parallelism advice is "--group-list ... for parallelized run" (cmd_cram_demuxlet.cpp:75): split the
 * barcodes by hand and run several processes.  pscl_multi does that split inside the library.
 *   demuxlet   barcodes sharded into contiguous ranges balanced by pair count, genotype table
 *              replicated, no collective (cells are independent: cmd_cram_demuxlet.cpp:636-1013);
 *   freemuxlet SNPs sharded into ranges balanced by pair count; stage 1 + greedy seeding
 *              (cmd_cram_freemux2.cpp:117-261) on the first GPU, then per EM iteration (:373-605) one
 *              all-reduce of the C x npairs partial LLKs over NVLink peer memory (the library's own
 *              kernel: fixed rank order, the same bits on every GPU), redundant classification and
 *              an SNP-local M-step.
 * One host thread per GPU, bound to the CPUs next to it.  The results equal the single-GPU ones
 * bit for bit for demuxlet, and up to the summation order of the SNP ranges for freemuxlet.
 * ------------------------------------------------------------------------------------------ */
typedef struct pscl_multi pscl_multi;
#define PSCL_MULTI_MAX_GPUS 16
typedef struct pscl_multi_timing {   /* wall-clock of the last pscl_multi_*_run, per GPU (host clocks) */
  int32_t n_gpus, iters;             /* EM iterations executed (freemuxlet)                              */
  double total_ms;                   /* whole call                                                        */
  double seed_ms;                    /* freemuxlet: whole-pileup stage 1 + greedy seeding on the first GPU */
  double allreduce_ms;               /* freemuxlet: mean per iteration, measured only with
                                        PSCL_MULTI_TIME_ALLREDUCE=1 (adds two stream synchronisations)  */
  int64_t allreduce_bytes;           /* C * npairs * 8                                                    */
  double upload_ms[PSCL_MULTI_MAX_GPUS];   /* H2D of the shard (freemuxlet: whole pileup + SNP filter)   */
  double setup_ms[PSCL_MULTI_MAX_GPUS];    /* freemuxlet: SNP-major view, stage 1, initial M-step         */
  double compute_ms[PSCL_MULTI_MAX_GPUS];  /* demuxlet: score + fetch; freemuxlet: the EM loop             */
  double kernel_ms[PSCL_MULTI_MAX_GPUS];   /* demuxlet: device time of the scoring kernels (CUDA events)  */
  int64_t units[PSCL_MULTI_MAX_GPUS];      /* (cell, SNP) pairs this GPU owned                            */
} pscl_multi_timing;

/* gpu_ids[n_gpu] = CUDA ordinals (NULL = 0..n_gpu-1; n_gpu <= 0 = every visible GPU).  A device may be
 * listed more than once (tests). */
int pscl_multi_create(const int* gpu_ids, int n_gpu, pscl_multi** out, char* err, size_t errlen);
void pscl_multi_destroy(pscl_multi* m);
const char* pscl_multi_last_error(const pscl_multi* m);
int pscl_multi_size(const pscl_multi* m);
pscl_ctx* pscl_multi_ctx(pscl_multi* m, int i);  /* the per-GPU context (kernel selection, timing) */
/* Same arguments and results as pscl_demux_run / pscl_fmx_run. */
int pscl_multi_demux_run(pscl_multi* m, const pscl_pileup* host, const pscl_geno* geno,
                         const pscl_demux_opts* opts, pscl_demux_cell* out, double* llk_grid);
int pscl_multi_fmx_run(pscl_multi* m, const pscl_pileup* host, const pscl_fmx_opts* opts,
                       const int32_t* init_clust, pscl_fmx_cell* out, double* clust_gl,
                       int32_t* clust_cnt, pscl_fmx_result* res);
int pscl_multi_last_timing(const pscl_multi* m, pscl_multi_timing* out);
/* Runs the calling thread on the CPUs of the NUMA node GPU `device` hangs off
 * (/sys/bus/pci/devices/<bus id>/local_cpulist), so that its pinned buffers and staging copies are
 * node-local.  Multi-process hosts (one rank per GPU) call it before allocating. */
int pscl_bind_thread_to_device(int device);

/* ---- knobs used by tests and the benchmark harness ---- */
/* Launch on a caller-owned stream (e.g. torch's current stream) instead of the context's own. */
int pscl_set_stream(pscl_ctx* ctx, void* cuda_stream);
/* Fault injection for the error-path tests: the nth device allocation from now on (nth >= 1) fails as an exhausted
 * device would (the call then returns PSCL_ENOMEM and the context stays usable); 0 = off. */
int pscl_debug_fail_alloc(pscl_ctx* ctx, int nth);
/* Upper bound of the per-batch partial-grid scratch of pscl_demux_score (default 1 GiB). */
int pscl_set_partial_budget(pscl_ctx* ctx, size_t bytes);
/* Route every alpha grid through the general demuxlet kernel (parity tests of that kernel). */
int pscl_demux_force_general(pscl_ctx* ctx, int enable);
/* Accumulation kernel: 0 = automatic (k_demux_default for the alpha grid {0, 0.5} with <= 8 samples,
 * k_demux_poly otherwise), 1 = k_demux_default on gathered genotype rows (lane per pair; needs that default shape),
 * 6 = k_demux_default on dictionary-coded genotypes (a table of <= 256 distinct triples, i.e. hard calls as --field GT
 * produces them, is read as 8-bit codes; any other table falls back to the rows),
 * 7 = the same on genotype classes (when, besides, no SNP holds more than three distinct triples — one per genotype —
 * a pair's factors are computed once per class and picked per accumulator; same expressions, bit-identical records; a
 * table with a fourth triple on some SNP falls back to 6; 0 chooses 7, then 6, then 1 — PSCL_NO_CLS=1 in the
 * environment skips 7),
 * 5 = k_demux_ab (k_demux_cls with two warps per batch; default shape),
 * 2 = k_demux_general (9-FMA baseline, any shape), 3 = k_demux_cls (class-split records, TMA packet
 * ring; default shape), 4 = k_demux_poly (polynomial in alpha, any shape).  All are parity-tested
 * against the same oracle. */
int pscl_demux_select_kernel(pscl_ctx* ctx, int which);
/* Which of the above the last pscl_demux_score actually launched (1..7), 0 before the first one. */
int pscl_demux_last_kernel(const pscl_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* POPSCLE_B200_H */
