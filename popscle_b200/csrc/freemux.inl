// freemux.inl — freemuxlet on the device (part of the single TU popscle_b200.cu)
//
// Replaces cmd_cram_freemux2.cpp:117-605 (`popscle freemuxlet`) and the EM rules of
// cmd_cram_freemuxlet.cpp:456-653 (`freemuxlet-old`, opts.mode_old):
//
//   k_fmx_stage1      per (cell,SNP) pair 9-genotype likelihoods (sc_drop_seq.cpp:452-509) and the
//                     per-cell singlet score sums (cmd_cram_freemux2.cpp:138-159)
//   k_fmx_seed        greedy sequential cluster seeding (:223-260, sc_drop_seq.cpp:544-578), one
//                     persistent CTA walking the cells in score order
//   k_fmx_mstep       cluster pileups rebuilt by ordered merges (:277-288, :590-596,
//                     sc_drop_seq.h:77-101); one thread per SNP replays the merges of the SNP-major
//                     pair list in ascending cell id (the clamp makes merge order matter), the nS
//                     pileups of the SNP in a shared-memory column of the thread
//   k_fmx_posterior   per (SNP, cluster) genotype posterior u = (1-e) norm(h * gl_diag) + e h (:402-415):
//                     it does not depend on the cell, so it is built once per iteration, not per pair
//   k_fmx_estep       cell x cluster-pair LLK partials (:383-456); one lane per (cell,SNP) pair, running
//                     products in registers, posterior rows gathered into shared memory (cp.async)
//   k_fmx_classify    best/next scans, logAdd sums, SNG/DBL/AMB decision, nchanged (:458-584)
//
// Data layout: stage 1 keeps the pair GLs twice — plane-major [9][P] in cell-major pair order (lane
// per pair kernels read 9 coalesced streams) and record-major [P][9] in SNP-major order (the M-step
// walks one SNP's cells as one contiguous block).

#define PSCL_FMX_MAX_CLUSTERS 32  /* the 32-bit cluster masks of the speculative seeding; the reference has no cap */
#define PSCL_MIN_NORM_GL 1e-6 /* sc_drop_seq.h:14 */

struct pscl_fmx_state {
  const pscl_plp* plp = nullptr;
  pscl_fmx_opts o{};
  int nS = 0, npairs = 0, US = 0, SD = 0;  // US: posterior row stride (doubles, even); SD: smem row stride
  int32_t C = 0, V = 0;
  int64_t P = 0;
  // SNP-major view of the pair list (ascending cell id inside one SNP)
  int64_t* snp_ptr = nullptr;    // [V+1]
  uint32_t* snp_pair = nullptr;  // [P] cell-major pair id of each SNP-major entry
  // warp-grouped SNP-major layout of the M-step: 32 consecutive SNPs form a group; entry t of SNP v sits in
  // slot grp_base[v/32] + t, lane v%32, so a warp reads entry t of its 32 SNPs as contiguous 32-element rows
  int64_t* grp_base = nullptr;   // [ceil(V/32)+1] first slot of each group (its length = the longest list in it)
  int32_t* cell_w = nullptr;     // [slots][32] cell of the entry, -1 = padding
  uint32_t* csc_pos = nullptr;   // [P] slot of each cell-major pair (stage 1 scatters through it)
  int64_t n_slots = 0;
  double* gl_soa = nullptr;      // [9][P]  cell-major
  double* gl_csc = nullptr;      // [slots][32][9]  pair GLs in the warp-grouped SNP-major layout
  double* clust_diag = nullptr;  // [V][nS][3]  diagonal GLs of the cluster pileups (all the E-step reads)
  double* u_tab = nullptr;       // [V][US]     per (SNP, cluster) posterior
  double* clust_gl = nullptr;    // [V][nS][9]  full cluster pileups (seeding, final output)
  int32_t* clust_cnt = nullptr;  // [V][nS][3]
  uint8_t* present = nullptr;    // [V][nS]     map membership during seeding (sc_drop_seq.cpp:551-552)
  pscl_fmx_cell* cells = nullptr;  // [C]
  int32_t* member = nullptr;       // [C] cluster the next M-step merges the cell into (-1 = none)
  double* item_s1 = nullptr;       // [n_items][2] stage-1 partial llk0 / llk2
  int32_t* item_nrd = nullptr;     // [n_items]    stage-1 partial read count
  double* item_llk = nullptr;      // [n_items][npairs]
  int32_t* counters = nullptr;     // [4] nchanged, nsingle, namb
  int32_t* order = nullptr;        // [C] seeding order
  int* work_counter = nullptr;
  double* own_stage1 = nullptr;    // buffers of the one-call path
  double* own_llk = nullptr;
  int32_t* own_clust = nullptr;
  bool stage1_done = false, begun = false, final_tab = false;
  int iters = 0;
  pscl_fmx_result last{};
  float ms_estep = 0.f, ms_mstep = 0.f, ms_classify = 0.f;
};

static void fmx_state_free(pscl_ctx* ctx) {
  pscl_fmx_state* s = ctx->fmx;
  if (!s) return;
  cudaFree(s->snp_ptr); cudaFree(s->snp_pair); cudaFree(s->grp_base); cudaFree(s->cell_w); cudaFree(s->csc_pos);
  cudaFree(s->gl_soa); cudaFree(s->gl_csc); cudaFree(s->clust_diag); cudaFree(s->u_tab);
  cudaFree(s->clust_gl); cudaFree(s->clust_cnt); cudaFree(s->present); cudaFree(s->cells);
  cudaFree(s->member); cudaFree(s->item_s1); cudaFree(s->item_nrd); cudaFree(s->item_llk);
  cudaFree(s->counters); cudaFree(s->order); cudaFree(s->work_counter);
  cudaFree(s->own_stage1); cudaFree(s->own_llk); cudaFree(s->own_clust);
  delete s;
  ctx->fmx = nullptr;
}

// ------------------------------------------------------------------------------------------------
// SNP-major view
// ------------------------------------------------------------------------------------------------
__global__ void k_fmx_iota(uint32_t* v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}

// after the stable sort by SNP: segment starts
__global__ void k_fmx_snp_ptr(const int32_t* __restrict__ key, int32_t V, int64_t P, int64_t* __restrict__ snp_ptr) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int32_t k = key[i];
  const int32_t prev = (i == 0) ? -1 : key[i - 1];
  for (int32_t v = prev + 1; v <= k; ++v) snp_ptr[v] = i;
  if (i == P - 1)
    for (int32_t v = k + 1; v <= V; ++v) snp_ptr[v] = P;
}
// longest list of each group of 32 SNPs (entry n_groups = 0 so that the exclusive scan ends with the total)
__global__ void k_fmx_group_len(const int64_t* __restrict__ snp_ptr, int32_t V, int32_t n_groups, int64_t* __restrict__ glen) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > n_groups) return;
  int64_t m = 0;
  if (g < n_groups)
    for (int v = g * 32; v < min(V, g * 32 + 32); ++v) m = max(m, snp_ptr[v + 1] - snp_ptr[v]);
  glen[g] = m;
}
// owning cell of each entry in the warp-grouped layout, slot of each cell-major pair
__global__ void k_fmx_csc_fill(const int32_t* __restrict__ key, const uint32_t* __restrict__ val,
                               const int64_t* __restrict__ cell_ptr, int32_t C, int64_t P, const int64_t* __restrict__ snp_ptr,
                               const int64_t* __restrict__ grp_base, int32_t* __restrict__ cell_w, uint32_t* __restrict__ csc_pos) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int32_t v = key[i];
  const uint32_t p = val[i];
  const int64_t slot = grp_base[v >> 5] + (i - snp_ptr[v]);
  csc_pos[p] = (uint32_t)slot;
  int lo = 0, hi = C;  // largest c with cell_ptr[c] <= p
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (cell_ptr[mid] <= (int64_t)p) lo = mid; else hi = mid;
  }
  cell_w[slot * 32 + (v & 31)] = lo;
}

__global__ void k_fmx_fill_i64(int64_t* v, int64_t n, int64_t x) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = x;
}

// ------------------------------------------------------------------------------------------------
// stage 1 (sc_drop_seq.cpp:452-509 + cmd_cram_freemux2.cpp:138-159)
// ------------------------------------------------------------------------------------------------
struct S1Args {
  const int32_t* pair_snp;
  const uint32_t* pair_rd;
  const uint8_t* rd_aq;
  const double* snp_af;
  const double* phred_err;
  const uint32_t* csc_pos;
  const int64_t* item_pbeg;
  const int64_t* item_pend;
  double* gl_soa;
  double* gl_csc;
  double* item_s1;
  int32_t* item_nrd;
  int64_t P;
  int32_t n_items;
};

__global__ void __launch_bounds__(256) k_fmx_stage1(S1Args a) {
  __shared__ double s_err[64];
  if (threadIdx.x < 64) s_err[threadIdx.x] = a.phred_err[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  // the 9 mixing weights of a REF base at alpha = 0.5 (:482-490); ALT is the mirror image
  const double wref[9] = {1.0, 0.75, 0.5, 0.75, 0.5, 0.25, 0.5, 0.25, 0.0};
  for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < a.n_items; item += warps) {
    const int64_t pb = a.item_pbeg[item], pe = a.item_pend[item];
    double l0 = 0.0, l2 = 0.0;
    int nrd_all = 0;
    for (int64_t p = pb + lane; p < pe; p += 32) {
      const uint32_t r0 = a.pair_rd[p], r1 = a.pair_rd[p + 1];
      double gl[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) gl[i] = 1.0;
      for (uint32_t r = r0; r < r1; ++r) {
        const uint32_t aq = a.rd_aq[r], al = aq >> 6;
        if (al > 1) continue;  // :467 (still counted in nreads, :465)
        const double err = s_err[aq & 63u];
        const double mat = 1.0 - err, e4 = err / 4.;  // PhredHelper.cpp:32
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {  // no FMA contraction: the reference multiplies, then adds
          gl[i] = __dmul_rn(gl[i], __dadd_rn(__dmul_rn(mat, al == 0 ? wref[i] : wref[8 - i]), e4));
          t = __dadd_rn(t, gl[i]);
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) gl[i] /= t;  // :493-494
      }
      nrd_all += (int)(r1 - r0);
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < 9; ++i) { gl[i] = (gl[i] < PSCL_MIN_NORM_GL) ? PSCL_MIN_NORM_GL : gl[i]; t = __dadd_rn(t, gl[i]); }  // :498-503
#pragma unroll
      for (int i = 0; i < 9; ++i) gl[i] /= t;
      const uint32_t q = a.csc_pos[p];
      const int32_t snp = a.pair_snp[p];
#pragma unroll
      for (int i = 0; i < 9; ++i) { a.gl_soa[(size_t)i * a.P + p] = gl[i]; a.gl_csc[((size_t)q * 32 + (snp & 31)) * 9 + i] = gl[i]; }
      const double af = a.snp_af[snp];
      double h[3];
      h[0] = __dmul_rn(1.0 - af, 1.0 - af); h[1] = __dmul_rn(__dmul_rn(2.0, af), 1.0 - af); h[2] = __dmul_rn(af, af);
      double lk0 = 0.0, lk2 = 0.0;
#pragma unroll
      for (int gi = 0; gi < 3; ++gi) {  // cmd_cram_freemux2.cpp:143-148
        lk2 = __dadd_rn(lk2, __dmul_rn(gl[gi * 3 + gi], h[gi]));
#pragma unroll
        for (int gj = 0; gj < 3; ++gj) lk0 = __dadd_rn(lk0, __dmul_rn(__dmul_rn(gl[gi * 3 + gj], h[gi]), h[gj]));
      }
      l0 += log(lk0);
      l2 += log(lk2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l0 += __shfl_xor_sync(0xffffffffu, l0, o);
      l2 += __shfl_xor_sync(0xffffffffu, l2, o);
      nrd_all += __shfl_xor_sync(0xffffffffu, nrd_all, o);
    }
    if (lane == 0) { a.item_s1[2 * item] = l0; a.item_s1[2 * item + 1] = l2; a.item_nrd[item] = nrd_all; }
  }
}

// per-cell sums of the item partials in item order -> planes llk0 | llk2 | nsnps | nreads of stage1[4*C]
__global__ void k_fmx_stage1_cells(const int32_t* __restrict__ cell_item_ptr, const int64_t* __restrict__ cell_ptr,
                                   const double* __restrict__ item_s1, const int32_t* __restrict__ item_nrd,
                                   int32_t C, double* __restrict__ stage1) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double l0 = 0.0, l2 = 0.0;
  int64_t nr = 0;
  for (int it = cell_item_ptr[c]; it < cell_item_ptr[c + 1]; ++it) { l0 += item_s1[2 * it]; l2 += item_s1[2 * it + 1]; nr += item_nrd[it]; }
  stage1[c] = l0;
  stage1[(size_t)C + c] = l2;
  stage1[(size_t)2 * C + c] = (double)(cell_ptr[c + 1] - cell_ptr[c]);
  stage1[(size_t)3 * C + c] = (double)nr;
}

// cell records from the (all-reduced) stage-1 planes; initial membership (cmd_cram_freemux2.cpp:192-216, :350-370)
__global__ void k_fmx_begin(const double* __restrict__ stage1, const int32_t* __restrict__ init_clust, int32_t C,
                            pscl_fmx_cell* __restrict__ cells, int32_t* __restrict__ clust, double* __restrict__ score) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  pscl_fmx_cell r;
  memset(&r, 0, sizeof(r));
  r.llk0 = stage1[c];
  r.llk2 = stage1[(size_t)C + c];
  r.n_snps = (int32_t)stage1[(size_t)2 * C + c];
  r.n_reads = (int32_t)stage1[(size_t)3 * C + c];
  score[c] = r.llk2 - r.llk0;  // :159
  const int ic = (init_clust && init_clust[c] >= 0) ? init_clust[c] : -1;
  r.type = (ic >= 0) ? 0 : -1;
  r.clust = r.init_clust = ic;
  r.best_j = r.best_k = r.next_j = r.next_k = -1;
  r.sng_best = r.sng_next = r.dbl_best_j = r.dbl_best_k = r.dbl_next_j = r.dbl_next_k = -1;
  r.best_llk = r.next_llk = r.sng_best_llk = r.sng_next_llk = r.dbl_best_llk = r.dbl_next_llk = -1e300;
  r.best_pp = r.sng_pp = r.sng_only_pp = r.sum_llk = -1e300;
  cells[c] = r;
  clust[c] = ic;
}

// ------------------------------------------------------------------------------------------------
// merge (sc_drop_seq.h:77-101); the reference divides by the sum, here one reciprocal per
// normalisation (<= 1 ulp apart; bit parity with glibc log() is out of reach anyway)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fmx_merge(double (&gl)[9], const double (&o)[9]) {
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) { gl[i] *= o[i]; t += gl[i]; }
  double r = 1.0 / t;
  t = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) { gl[i] *= r; gl[i] = (gl[i] < PSCL_MIN_NORM_GL) ? PSCL_MIN_NORM_GL : gl[i]; t += gl[i]; }
  r = 1.0 / t;
#pragma unroll
  for (int i = 0; i < 9; ++i) gl[i] *= r;
}

// ------------------------------------------------------------------------------------------------
// greedy seeding (cmd_cram_freemux2.cpp:223-260): one CTA, cells strictly in score order
// ------------------------------------------------------------------------------------------------
struct SeedArgs {
  const int32_t* order;
  const double* score;
  const int64_t* cell_ptr;
  const int32_t* pair_snp;
  const double* gl_soa;
  const double* snp_af;
  double* clust_gl;   // [V][nS][9], initialised to 1.0
  uint8_t* present;   // [V][nS], initialised to 0
  int32_t* clust;     // [C] out
  pscl_fmx_cell* cells;
  int64_t P;
  int32_t C, nS;
  double frac, thres;
};

__global__ void __launch_bounds__(1024, 1) k_fmx_seed(SeedArgs a) {
  __shared__ double s_d2[32], s_d0[32], s_sc[PSCL_FMX_MAX_CLUSTERS];
  __shared__ int s_choice;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nS = a.nS;
  const int wpc = 32 / nS;            // warps per cluster (nS <= 32)
  const int my_j = warp / wpc, my_w = warp % wpc;
  const bool active = my_j < nS;
  for (int i = 0; i < a.C; ++i) {
    const int si = a.order[i];
    if ((double)i > (double)a.C * a.frac) continue;  // :225
    if (a.score[si] < a.thres) continue;             // :226
    const int64_t pb = a.cell_ptr[si], pe = a.cell_ptr[si + 1];
    // ---- distance to every cluster over the SNPs present in both (sc_drop_seq.cpp:544-578) ----
    double d2 = 0.0, d0 = 0.0;
    if (active) {
      for (int64_t p = pb + my_w * 32 + lane; p < pe; p += (int64_t)wpc * 32) {
        const int32_t s = a.pair_snp[p];
        const size_t e = (size_t)s * nS + my_j;
        if (!a.present[e]) continue;
        const double af = a.snp_af[s];
        const double h0 = (1.0 - af) * (1.0 - af), h1 = 2.0 * af * (1.0 - af), h2 = af * af;
        const double ci0 = a.gl_soa[p], ci1 = a.gl_soa[(size_t)4 * a.P + p], ci2 = a.gl_soa[(size_t)8 * a.P + p];
        const double* cj = a.clust_gl + e * 9;
        const double cj0 = cj[0], cj1 = cj[4], cj2 = cj[8];
        const double lk2 = ci0 * cj0 * h0 + ci1 * cj1 * h1 + ci2 * cj2 * h2;
        const double a0 = ci0 * h0, a1 = ci1 * h1, a2 = ci2 * h2, b0 = cj0 * h0, b1 = cj1 * h1, b2 = cj2 * h2;
        const double lk0 = (a0 + a1 + a2) * (b0 + b1 + b2);  // sum_g sum_h ci[g] cj[h] h[g] h[h]
        d2 += log(lk2);
        d0 += log(lk0);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      d2 += __shfl_xor_sync(0xffffffffu, d2, o);
      d0 += __shfl_xor_sync(0xffffffffu, d0, o);
    }
    if (lane == 0) { s_d2[warp] = d2; s_d0[warp] = d0; }
    __syncthreads();
    if (tid < nS) {
      double t2 = 0.0, t0 = 0.0;
      for (int w = 0; w < wpc; ++w) { t2 += s_d2[tid * wpc + w]; t0 += s_d0[tid * wpc + w]; }
      s_sc[tid] = t2 - t0;  // dropDs[j].llk2 - dropDs[j].llk0
    }
    __syncthreads();
    if (tid == 0) {
      int best = 0;
      double bs = s_sc[0];
      for (int j = 1; j < nS; ++j)
        if (s_sc[j] > bs) { best = j; bs = s_sc[j]; }  // :238-241, first wins
      s_choice = best;
      a.clust[si] = best;
      a.cells[si].clust = a.cells[si].init_clust = best;
      a.cells[si].type = 0;
    }
    __syncthreads();
    const int jstar = s_choice;
    // ---- merge the cell into the chosen cluster (:248-251) ----
    for (int64_t p = pb + tid; p < pe; p += 1024) {
      const size_t e = (size_t)a.pair_snp[p] * nS + jstar;
      double gl[9], o[9];
      double* cg = a.clust_gl + e * 9;
#pragma unroll
      for (int g = 0; g < 9; ++g) { gl[g] = cg[g]; o[g] = a.gl_soa[(size_t)g * a.P + p]; }
      fmx_merge(gl, o);
#pragma unroll
      for (int g = 0; g < 9; ++g) cg[g] = gl[g];
      a.present[e] = 1;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// greedy seeding, batched: the same chain of decisions, but only what really depends on the previous cell is serial
// ------------------------------------------------------------------------------------------------
// The distance of cell i to cluster j is a sum over the SNPs the two share (sc_drop_seq.cpp:544-578), and the state of
// cluster j at SNP s changes only when a cell covering s is merged into j.  A batch of B cells (consecutive in seeding
// order) is therefore handled in three launches:
//   k_fmx_seed_dist    (all SMs) every (cell of the batch, SNP, cluster) contribution against the cluster table as it
//                      stands BEFORE the batch, kept per (pair, cluster) in `contrib`, plus their per-cell sums; and a
//                      count, per SNP, of the batch's cells that cover it (mark[s]).
//   k_fmx_seed_commit  (one CTA) walks the B cells in order, looking only at the pairs whose SNP another cell of the batch
//                      covers too (compacted into a dense list: cells share few SNPs, K^2/V = 40 of 2000 at configs[2]).
//                      A cell's true distance = snapshot sum + the corrections at the SNPs an EARLIER cell of the batch
//                      merged into (cluster bit mask in mark[s]): contribution against the CURRENT table minus the
//                      stored one.  Then argmax (first wins, :238-241) and the merge (:248-251) at those shared SNPs.
//   k_fmx_seed_merge   (all SMs) the merges at the SNPs only one cell of the batch covers: nobody else in the batch can
//                      see them, so they run in parallel once the clusters are decided.
// The decisions are those of the serial chain up to the summation order of the distance (a reordering of ~1e3 terms; ties
// within ~1e-13 relative could flip, as they could between any two compilers of the reference).
#define PSCL_SEED_SPLIT 4      /* CTAs per batch cell in k_fmx_seed_dist / k_fmx_seed_merge */
#define PSCL_SEED_LIST 8192    /* shared pairs of one cell the commit kernel lists at a time */
struct SeedBatchArgs {
  const int32_t* elig;        // [n_elig] cells that take part, in seeding order
  const int64_t* epair;       // [n_elig + 1] running pair count of those cells (offsets into contrib, relative to the batch)
  const int64_t* cell_ptr;
  const int32_t* pair_snp;
  const double* gl_soa;
  const double* snp_af;
  double* clust_gl;           // [V][nS][9]
  uint8_t* present;           // [V][nS]
  unsigned long long* mark;   // [V] (batch + 1) << 32 | cells of the batch covering the SNP (saturating) << 24 | mask of the
                              //     clusters merged into at this SNP by the commit kernel during that batch
  double* contrib;            // [pairs of the batch][nS]
  double* d0p;                // [B][PSCL_SEED_SPLIT][nS]
  int32_t* clust;
  pscl_fmx_cell* cells;
  int64_t P;
  int32_t nS, base, nb, batch;
};

// log lk2 - log lk0 of one (cell pair, cluster) (sc_drop_seq.cpp:556-571), diagonals only on both sides
__device__ __forceinline__ double fmx_seed_term(double ci0, double ci1, double ci2, double cj0, double cj1, double cj2, double h0, double h1, double h2) {
  const double lk2 = ci0 * cj0 * h0 + ci1 * cj1 * h1 + ci2 * cj2 * h2;
  const double lk0 = (ci0 * h0 + ci1 * h1 + ci2 * h2) * (cj0 * h0 + cj1 * h1 + cj2 * h2);  // sum_g sum_h ci[g] cj[h] h[g] h[h]
  return log(lk2) - log(lk0);
}
__device__ __forceinline__ void fmx_seed_merge_pair(const SeedBatchArgs& a, int64_t p, int32_t s, int j) {
  const size_t e = (size_t)s * a.nS + j;
  double gl[9], o[9];
  double* cg = a.clust_gl + e * 9;
#pragma unroll
  for (int g = 0; g < 9; ++g) { gl[g] = cg[g]; o[g] = a.gl_soa[(size_t)g * a.P + p]; }
  fmx_merge(gl, o);
#pragma unroll
  for (int g = 0; g < 9; ++g) cg[g] = gl[g];
  a.present[e] = 1;
}

template <int NSM>
__global__ void __launch_bounds__(256) k_fmx_seed_dist(SeedBatchArgs a) {
  __shared__ double s_w[8][NSM];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nS = a.nS;
  const int b = blockIdx.x, q = blockIdx.y;
  const int si = a.elig[a.base + b];
  const int64_t pb = a.cell_ptr[si], K = a.cell_ptr[si + 1] - pb;
  const int64_t t0 = K * q / PSCL_SEED_SPLIT, t1 = K * (q + 1) / PSCL_SEED_SPLIT;
  double* const out = a.contrib + (size_t)(a.epair[a.base + b] - a.epair[a.base]) * nS;
  const unsigned long long stamp = (unsigned long long)(a.batch + 1) << 32;
  double sum[NSM];
#pragma unroll
  for (int j = 0; j < NSM; ++j) sum[j] = 0.0;
  for (int64_t t = t0 + tid; t < t1; t += 256) {
    const int64_t p = pb + t;
    const int32_t s = a.pair_snp[p];
    {  // one more cell of this batch covers SNP s
      unsigned long long old = a.mark[s], want;
      do {
        const unsigned long long cur = old;
        want = ((cur >> 32) != (stamp >> 32)) ? (stamp | (1ull << 24)) : (((cur >> 24) & 0xffull) < 0xffull ? cur + (1ull << 24) : cur);
        old = atomicCAS(a.mark + s, cur, want);
        if (old == cur) break;
      } while (true);
    }
    const double af = a.snp_af[s];
    const double h0 = (1.0 - af) * (1.0 - af), h1 = 2.0 * af * (1.0 - af), h2 = af * af;
    const double ci0 = a.gl_soa[p], ci1 = a.gl_soa[(size_t)4 * a.P + p], ci2 = a.gl_soa[(size_t)8 * a.P + p];
#pragma unroll
    for (int j = 0; j < NSM; ++j) {
      if (j < nS) {
        const size_t e = (size_t)s * nS + j;
        double c = 0.0;
        if (a.present[e]) {
          const double* cj = a.clust_gl + e * 9;
          c = fmx_seed_term(ci0, ci1, ci2, cj[0], cj[4], cj[8], h0, h1, h2);
        }
        out[(size_t)t * nS + j] = c;
        sum[j] += c;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NSM; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum[j] += __shfl_xor_sync(0xffffffffu, sum[j], o);
    if (lane == 0) s_w[warp][j] = sum[j];
  }
  __syncthreads();
  if (tid < nS) {
    double x = 0.0;
    for (int w = 0; w < 8; ++w) x += s_w[w][tid];
    a.d0p[((size_t)b * PSCL_SEED_SPLIT + q) * nS + tid] = x;
  }
}

template <int NSM>
__global__ void __launch_bounds__(1024, 1) k_fmx_seed_commit(SeedBatchArgs a) {
  __shared__ double s_w[32][NSM], s_sc[NSM];
  __shared__ int s_choice, s_wcnt[32], s_n;
  __shared__ int s_list[PSCL_SEED_LIST];  // pair offsets (inside the cell) whose SNP another cell of the batch covers too
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nS = a.nS;
  const unsigned long long stamp = (unsigned long long)(a.batch + 1) << 32;
  for (int b = 0; b < a.nb; ++b) {
    const int si = a.elig[a.base + b];
    const int64_t pb = a.cell_ptr[si], K = a.cell_ptr[si + 1] - pb;
    const double* const stored = a.contrib + (size_t)(a.epair[a.base + b] - a.epair[a.base]) * nS;
    double delta[NSM];
#pragma unroll
    for (int j = 0; j < NSM; ++j) delta[j] = 0.0;
    int jstar = -1;
    // two passes over the cell's pairs, PSCL_SEED_LIST at a time: pass 0 lists the shared pairs and sums the corrections
    // (the decision follows it), pass 1 lists them again and merges
    for (int pass = 0; pass < 2; ++pass) {
      for (int64_t c0 = 0; c0 < K; c0 += PSCL_SEED_LIST) {
        const int64_t c1 = c0 + PSCL_SEED_LIST < K ? c0 + PSCL_SEED_LIST : K;
        // ---- deterministic compaction: position = (iteration, warp, lane) order ----
        // (a cell of <= PSCL_SEED_LIST pairs keeps pass 0's list for pass 1: the per-SNP cover counts do not change
        // inside the commit kernel)
        int n_list = (pass == 1 && K <= PSCL_SEED_LIST) ? s_n : 0;
        for (int64_t t = c0 + tid; t < c0 + (((c1 - c0) + 1023) & ~(int64_t)1023) && !(pass == 1 && K <= PSCL_SEED_LIST); t += 1024) {
          bool sh = false;
          if (t < c1) sh = ((a.mark[a.pair_snp[pb + t]] >> 24) & 0xffull) >= 2ull;
          const unsigned bal = __ballot_sync(0xffffffffu, sh);
          if (lane == 0) s_wcnt[warp] = __popc(bal);
          __syncthreads();
          int before = 0, total = 0;
          for (int w = 0; w < 32; ++w) { const int c = s_wcnt[w]; before += w < warp ? c : 0; total += c; }
          if (sh) s_list[n_list + before + __popc(bal & ((1u << lane) - 1u))] = (int)(t - c0);
          n_list += total;
          __syncthreads();
        }
        if (pass == 0 && tid == 0) s_n = n_list;
        // ---- the listed pairs, one per thread ----
        for (int q = tid; q < n_list; q += 1024) {
          const int64_t t = c0 + s_list[q];
          const int64_t p = pb + t;
          const int32_t s = a.pair_snp[p];
          if (pass == 0) {
            unsigned mask = (unsigned)a.mark[s] & 0xffffffu;  // clusters an earlier cell of this batch merged into at s
            if (mask) {
              const double af = a.snp_af[s];
              const double h0 = (1.0 - af) * (1.0 - af), h1 = 2.0 * af * (1.0 - af), h2 = af * af;
              const double ci0 = a.gl_soa[p], ci1 = a.gl_soa[(size_t)4 * a.P + p], ci2 = a.gl_soa[(size_t)8 * a.P + p];
              while (mask) {  // usually one bit: the lanes loop together whatever cluster each one corrects
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                const double* cj = a.clust_gl + ((size_t)s * nS + j) * 9;
                const double v = fmx_seed_term(ci0, ci1, ci2, cj[0], cj[4], cj[8], h0, h1, h2) - stored[(size_t)t * nS + j];
#pragma unroll
                for (int jj = 0; jj < NSM; ++jj) delta[jj] += (jj == j) ? v : 0.0;
              }
            }
          } else {
            fmx_seed_merge_pair(a, p, s, jstar);
            a.mark[s] |= 1ull << jstar;  // the stamp is this batch's (the SNP is covered by >= 2 of its cells)
          }
        }
        __syncthreads();
      }
      if (pass == 0) {
#pragma unroll
        for (int j = 0; j < NSM; ++j) {
          if (j < nS) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) delta[j] += __shfl_xor_sync(0xffffffffu, delta[j], o);
            if (lane == 0) s_w[warp][j] = delta[j];
          }
        }
        __syncthreads();
        if (tid < nS) {
          double x = 0.0;
          for (int q = 0; q < PSCL_SEED_SPLIT; ++q) x += a.d0p[((size_t)b * PSCL_SEED_SPLIT + q) * nS + tid];
          double d = 0.0;
          for (int w = 0; w < 32; ++w) d += s_w[w][tid];
          s_sc[tid] = x + d;
        }
        __syncthreads();
        if (tid == 0) {
          int best = 0;
          double bs = s_sc[0];
          for (int j = 1; j < nS; ++j)
            if (s_sc[j] > bs) { best = j; bs = s_sc[j]; }  // :238-241, first wins
          s_choice = best;
          a.clust[si] = best;
          a.cells[si].clust = a.cells[si].init_clust = best;
          a.cells[si].type = 0;
        }
        __syncthreads();
        jstar = s_choice;
      }
    }
    (void)stamp;
    __syncthreads();
  }
}

// the merges nobody else in the batch can see: SNPs covered by exactly one of its cells
__global__ void __launch_bounds__(256) k_fmx_seed_merge(SeedBatchArgs a) {
  const int b = blockIdx.x, q = blockIdx.y;
  const int si = a.elig[a.base + b];
  const int64_t pb = a.cell_ptr[si], K = a.cell_ptr[si + 1] - pb;
  const int64_t t0 = K * q / PSCL_SEED_SPLIT, t1 = K * (q + 1) / PSCL_SEED_SPLIT;
  const int j = a.clust[si];
  for (int64_t t = t0 + threadIdx.x; t < t1; t += 256) {
    const int32_t s = a.pair_snp[pb + t];
    if (((a.mark[s] >> 24) & 0xffull) < 2ull) fmx_seed_merge_pair(a, pb + t, s, j);
  }
}

// ------------------------------------------------------------------------------------------------
// greedy seeding, speculative batches (the default): every cell of a batch decided at once, then proven
// ------------------------------------------------------------------------------------------------
// The chain "cell r joins the cluster it is closest to, given the merges of cells 0..r-1" (cmd_cram_freemux2.cpp:223-260)
// is serial only on paper: once the clusters have some weight, the few merges of the cells just ahead of r almost never
// change r's decision.  A batch of B consecutive cells (in seeding order) is therefore handled like this:
//   E0  every cell, in parallel, against the cluster table as it stands before the batch              -> guess A
//   E1  every cell again, its clusters now being the table PLUS the merges of the batch's earlier cells at the SNPs it
//       shares with them, those cells placed as guess A says (folded in seeding order, per cluster)     -> B
//       Cell 0 of the batch has no predecessor, so B[0] is the chain's decision; if A[0..m) == B[0..m) then B[m] saw exactly
//       the chain's table, so B is the chain's decision up to and including the first index m where A and B differ.
//   E2  (only if they differ somewhere) the cells after m once more, with B as the placement of their predecessors -> C,
//       right up to and including the first index where B and C differ.
//   C   commits the proven prefix: cluster ids out, merges folded into the table per (SNP, cluster) in seeding order, and
//       the next batch (start, size: doubled after a clean batch, shrunk to what was proven otherwise) written to ctrl[k+1].
// Every decision is proven against exactly the table the serial chain would have seen; what differs from the chain is only
// the summation order of a distance (as in the batched form above).  All SMs work on every step: a cell is one CTA.
// The lists of a SNP's cells in seeding order (rk_*) come from one radix sort of (SNP, rank) keys.
struct Seed3Ctrl { int32_t r0, nb; };
struct Seed3Args {
  const int32_t* elig;        // [n_elig] cells that take part, in seeding order (rank -> cell)
  const int64_t* cell_ptr;
  const int32_t* pair_snp;
  const double* gl_soa;
  const double* gl_rk;        // [P][9] the pair GLs as 72-byte records in the order of rk: the earlier cells of a batch at a
                              //        SNP are the records just before a pair's own
  const double* snp_af;
  const int64_t* snp_ptr;     // [V+1] SNP-major lists ...
  const uint32_t* rk;         // [P]   ... rank of each entry's cell; one SNP's entries in ascending rank (cells outside elig
                              //       last, with rank n_elig)
  const uint32_t* pos;        // [P]   entry of each cell-major pair
  double* clust_gl;           // [V][nS][9]
  uint8_t* present;           // [V][nS]
  double* diag;               // [V][nS][3] diagonal of clust_gl, -1 while the cluster has no pileup at the SNP
  double* sc0;                // [bmax][NSM] E0's distances
  double* part;               // [bmax][Q][NSM] partial sums of the CTAs of one cell (a cell's pairs are split over Q CTAs)
  int32_t* arrived;           // [bmax] CTAs of the cell that have delivered their part (the last one adds them up)
  int32_t *dA, *dB, *dC;      // [bmax] decisions of the three evaluation rounds
  Seed3Ctrl* ctrl;            // [batches + 1]
  int32_t* stats;             // [4] batches, batches that needed E2, cells proven by E2, smallest batch
  int32_t* clust;
  pscl_fmx_cell* cells;
  int64_t P;
  int32_t nS, n_elig, bmin, bmax, Q;
};

// first index in [0, n) where x and y differ (n if none), the same value in every thread of the CTA
__device__ __forceinline__ int seed3_first_mismatch(const int32_t* x, const int32_t* y, int n, int* s_m) {
  if (threadIdx.x == 0) *s_m = n;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (x[i] != y[i]) { atomicMin(s_m, i); break; }
  __syncthreads();
  const int m = *s_m;
  __syncthreads();
  return m;
}

// lk2 and lk0 of one (cell pair, cluster) (sc_drop_seq.cpp:556-571), diagonals only on both sides
__device__ __forceinline__ void fmx_seed_lks(double ci0, double ci1, double ci2, double cj0, double cj1, double cj2, double h0, double h1, double h2,
                                             double& lk2, double& lk0) {
  lk2 = ci0 * cj0 * h0 + ci1 * cj1 * h1 + ci2 * cj2 * h2;
  lk0 = (ci0 * h0 + ci1 * h1 + ci2 * h2) * (cj0 * h0 + cj1 * h1 + cj2 * h2);
}
// A distance is sum_s log lk2 - log lk0.  The seeding kernels keep it as a running quotient prod lk2 / prod lk0 per thread
// (mantissa, binary exponent pulled out every few terms) and take the logs per thread and cluster at the end instead of two
// per term: the terms are >= ~1e-13 (normalised GLs clamped at 1e-6, genotype priors that sum to one), so a handful of them
// cannot underflow.  In log space this is at least as exact as the chain's own sum of logs.
__device__ __forceinline__ void seed3_renorm(double& m, int& ex) {
  int k;
  m = frexp(m, &k);
  ex += k;
}
// Sums v[j] over the Q CTAs that share cell b: every CTA reduces its own threads and stores the part; the CTA that arrives
// last adds the Q parts in their fixed order (so the sum does not depend on who was last) into dst[j] (shared memory) and
// returns true, in every one of its threads.
template <int NSM, int NT>
__device__ __forceinline__ bool seed3_cell_sum(const Seed3Args& a, int b, int q, double (&v)[NSM], int nS, double (*s_w)[NSM], double* dst, int* s_last) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int j = 0; j < NSM; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    if (lane == 0) s_w[warp][j] = v[j];
  }
  __syncthreads();
  if (tid < nS) {
    double x = 0.0;
    for (int w = 0; w < NT / 32; ++w) x += s_w[w][tid];
    a.part[((size_t)b * a.Q + q) * NSM + tid] = x;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) {
    const int last = atomicAdd(a.arrived + b, 1) == a.Q - 1;
    if (last) a.arrived[b] = 0;  // ready for the next round (kernels of one stream do not overlap)
    *s_last = last;
  }
  __syncthreads();
  if (!*s_last) return false;
  __threadfence();
  if (tid < nS) {
    double x = 0.0;
    for (int qq = 0; qq < a.Q; ++qq) x += __ldcg(a.part + ((size_t)b * a.Q + qq) * NSM + tid);
    dst[tid] = x;
  }
  __syncthreads();
  return true;
}

// E0: every cell of the batch against the cluster table as it stands before the batch
template <int NSM, int NT>
__global__ void __launch_bounds__(NT, NSM <= 8 ? 3 : 2) k_fmx_seed3_eval0(Seed3Args a, int k) {
  __shared__ double s_w[NT / 32][NSM], s_sc[NSM];
  __shared__ int s_last;
  const Seed3Ctrl c = a.ctrl[k];
  const int b = blockIdx.x, q = blockIdx.y, tid = threadIdx.x, nS = a.nS;
  if (b >= c.nb) return;
  const int si = a.elig[c.r0 + b];
  const int64_t pb = a.cell_ptr[si], K = a.cell_ptr[si + 1] - pb;
  double num[NSM], den[NSM];
  int en[NSM], ed[NSM];
#pragma unroll
  for (int j = 0; j < NSM; ++j) { num[j] = 1.0; den[j] = 1.0; en[j] = 0; ed[j] = 0; }
  int it = 0;
  for (int64_t t = (int64_t)q * NT + tid; t < K; t += (int64_t)a.Q * NT, ++it) {
    const int64_t p = pb + t;
    const int32_t s = a.pair_snp[p];
    const double af = a.snp_af[s];
    const double h0 = (1.0 - af) * (1.0 - af), h1 = 2.0 * af * (1.0 - af), h2 = af * af;
    const double ci0 = a.gl_soa[p], ci1 = a.gl_soa[(size_t)4 * a.P + p], ci2 = a.gl_soa[(size_t)8 * a.P + p];
    const double* dg = a.diag + (size_t)s * nS * 3;
#pragma unroll
    for (int j = 0; j < NSM; ++j) {
      if (j < nS) {
        const double cj0 = dg[3 * j];
        if (cj0 >= 0.0) {  // the cluster has a pileup at this SNP (-1 otherwise)
          double lk2, lk0;
          fmx_seed_lks(ci0, ci1, ci2, cj0, dg[3 * j + 1], dg[3 * j + 2], h0, h1, h2, lk2, lk0);
          num[j] *= lk2; den[j] *= lk0;
        }
      }
    }
    if ((it & 7) == 7) {
#pragma unroll
      for (int j = 0; j < NSM; ++j) { seed3_renorm(num[j], en[j]); seed3_renorm(den[j], ed[j]); }
    }
  }
#pragma unroll
  for (int j = 0; j < NSM; ++j)
    num[j] = (j < nS && it > 0) ? (log(num[j]) - log(den[j])) + (double)(en[j] - ed[j]) * 0.69314718055994530942 : 0.0;
  if (!seed3_cell_sum<NSM, NT>(a, b, q, num, nS, s_w, s_sc, &s_last)) return;
  if (tid < nS) a.sc0[(size_t)b * NSM + tid] = s_sc[tid];
  if (tid == 0) {
    int best = 0;
    double bs = s_sc[0];
    for (int j = 1; j < nS; ++j)
      if (s_sc[j] > bs) { best = j; bs = s_sc[j]; }  // :238-241, first wins
    a.dA[b] = best;
  }
}

// E1 / E2: the same cells with the merges of the batch's earlier cells folded in.  Only the (SNP, cluster) entries such a
// merge touches differ from E0, so this kernel computes the change of each distance — (new lk2 / new lk0) over (old lk2 /
// old lk0) per touched entry, as a running product per (cluster, thread) in shared memory — and adds its log to E0's score.
// A thread walks back from its pair's entry in the SNP's rank-ordered list over the batch's earlier cells (adjacent
// entries), then takes the clusters they are placed in one at a time: cluster pileup (72 B), the earlier cells' GL records
// (72 B each, adjacent entries) folded in seeding order, one quotient.
template <int NSM, int ROUND, int NT>
__global__ void __launch_bounds__(NT) k_fmx_seed3_delta(Seed3Args a, int k) {
  extern __shared__ double s_dyn[];
  __shared__ double s_w[NT / 32][NSM], s_sc[NSM];
  __shared__ int s_m, s_last;
  const Seed3Ctrl c = a.ctrl[k];
  const int b = blockIdx.x, q = blockIdx.y, tid = threadIdx.x, nS = a.nS;
  if (b >= c.nb) return;
  const int32_t* specg = ROUND == 1 ? a.dA : a.dB;   // where the batch's earlier cells are assumed to go
  int32_t* const out = ROUND == 1 ? a.dB : a.dC;
  if (ROUND == 2) {
    const int m1 = seed3_first_mismatch(a.dA, a.dB, c.nb, &s_m);
    if (m1 >= c.nb) return;                          // E1 proved the whole batch
    if (b <= m1) { if (tid == 0 && q == 0) out[b] = a.dB[b]; return; }
  }
  if (b == 0) { if (tid == 0 && q == 0) out[0] = a.dA[0]; return; }  // no earlier cell in the batch: E0 is the chain's decision
  double* const acc = s_dyn;                              // [NSM][NT] running quotient of each cluster, this thread's column
  uint8_t* const spec = (uint8_t*)(s_dyn + (size_t)NSM * NT);  // [b] placements of the earlier cells
  for (int i = tid; i < b; i += NT) spec[i] = (uint8_t)specg[i];
#pragma unroll
  for (int j = 0; j < NSM; ++j) acc[j * NT + tid] = 1.0;
  __syncthreads();
  const uint32_t r0 = (uint32_t)c.r0;
  const int si = a.elig[c.r0 + b];
  const int64_t pb = a.cell_ptr[si], K = a.cell_ptr[si + 1] - pb;
  int ex[NSM];
#pragma unroll
  for (int j = 0; j < NSM; ++j) ex[j] = 0;
  int it = 0;
  for (int64_t t = (int64_t)q * NT + tid; t < K; t += (int64_t)a.Q * NT, ++it) {
    const int64_t p = pb + t;
    const int32_t s = a.pair_snp[p];
    const int64_t qe = (int64_t)a.pos[p], lo = a.snp_ptr[s];
    unsigned mask = 0;     // clusters an earlier cell of the batch is (assumed to be) merged into at this SNP
    int64_t q0 = qe;
    for (; q0 > lo; --q0) {
      const uint32_t r = __ldg(a.rk + q0 - 1);
      if (r < r0) break;
      mask |= 1u << spec[r - r0];
    }
    if (mask) {
      const double af = a.snp_af[s];
      const double h0 = (1.0 - af) * (1.0 - af), h1 = 2.0 * af * (1.0 - af), h2 = af * af;
      const double ci0 = __ldg(a.gl_rk + (size_t)qe * 9), ci1 = __ldg(a.gl_rk + (size_t)qe * 9 + 4), ci2 = __ldg(a.gl_rk + (size_t)qe * 9 + 8);
      while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        const size_t e = (size_t)s * nS + j;
        const double* cj = a.clust_gl + e * 9;
        double gl[9], o[9], lk2, lk0, up = 1.0, dn = 1.0;
#pragma unroll
        for (int g = 0; g < 9; ++g) gl[g] = cj[g];  // 1.0 everywhere while the cluster has nothing at this SNP
        if (a.present[e]) {  // the old term leaves the distance
          fmx_seed_lks(ci0, ci1, ci2, gl[0], gl[4], gl[8], h0, h1, h2, lk2, lk0);
          up = lk0; dn = lk2;
        }
        for (int64_t tt = q0; tt < qe; ++tt) {
          if (spec[__ldg(a.rk + tt) - r0] != j) continue;
          const double* src = a.gl_rk + (size_t)tt * 9;
#pragma unroll
          for (int g = 0; g < 9; ++g) o[g] = __ldg(src + g);
          fmx_merge(gl, o);
        }
        fmx_seed_lks(ci0, ci1, ci2, gl[0], gl[4], gl[8], h0, h1, h2, lk2, lk0);
        acc[j * NT + tid] *= (lk2 * up) / (lk0 * dn);
      }
    }
    if ((it & 3) == 3) {
#pragma unroll
      for (int j = 0; j < NSM; ++j) {
        double m = acc[j * NT + tid];
        if (m != 1.0) { seed3_renorm(m, ex[j]); acc[j * NT + tid] = m; }
      }
    }
  }
  double v[NSM];
#pragma unroll
  for (int j = 0; j < NSM; ++j) {
    const double m = acc[j * NT + tid];
    v[j] = (m == 1.0 && ex[j] == 0) ? 0.0 : log(m) + (double)ex[j] * 0.69314718055994530942;
  }
  if (!seed3_cell_sum<NSM, NT>(a, b, q, v, nS, s_w, s_sc, &s_last)) return;
  if (tid == 0) {
    int best = 0;
    double bs = 0.0;
    for (int j = 0; j < nS; ++j) {
      const double x = a.sc0[(size_t)b * NSM + j] + s_sc[j];
      if (j == 0 || x > bs) { best = j; bs = x; }  // :238-241, first wins
    }
    out[b] = best;
  }
}

template <int NT>
__global__ void __launch_bounds__(NT) k_fmx_seed3_commit(Seed3Args a, int k) {
  extern __shared__ double s_dyn[];
  __shared__ int s_m;
  const Seed3Ctrl c = a.ctrl[k];
  const int b = blockIdx.x, qc = blockIdx.y, tid = threadIdx.x, nS = a.nS;
  if (c.nb <= 0) {
    if (b == 0 && qc == 0 && tid == 0) a.ctrl[k + 1] = c;
    return;
  }
  const int m1 = seed3_first_mismatch(a.dA, a.dB, c.nb, &s_m);
  const int32_t* fing = a.dB;
  int ncommit = c.nb;
  if (m1 < c.nb) {
    const int m2 = seed3_first_mismatch(a.dB, a.dC, c.nb, &s_m);
    fing = a.dC;
    ncommit = m2 + 1 < c.nb ? m2 + 1 : c.nb;
  }
  if (b == 0 && qc == 0 && tid == 0) {
    Seed3Ctrl n;
    n.r0 = c.r0 + ncommit;
    int nb = ncommit == c.nb ? 2 * c.nb : ncommit;   // a clean batch doubles; otherwise as many as were proven
    nb = nb < a.bmin ? a.bmin : nb;
    nb = nb > a.bmax ? a.bmax : nb;
    n.nb = nb < a.n_elig - n.r0 ? nb : a.n_elig - n.r0;
    a.ctrl[k + 1] = n;
    a.stats[0] += 1;
    if (m1 < c.nb) { a.stats[1] += 1; a.stats[2] += ncommit - (m1 + 1); }
    if (a.stats[3] == 0 || c.nb < a.stats[3]) a.stats[3] = c.nb;
  }
  if (b >= ncommit) return;
  uint8_t* const fin = (uint8_t*)s_dyn;  // [ncommit] proven placements
  for (int i = tid; i < ncommit; i += NT) fin[i] = (uint8_t)fing[i];
  __syncthreads();
  const uint32_t r0 = (uint32_t)c.r0, r1 = (uint32_t)(c.r0 + ncommit);
  const int si = a.elig[c.r0 + b];
  const int j = fin[b];
  if (tid == 0 && qc == 0) {
    a.clust[si] = j;
    a.cells[si].clust = a.cells[si].init_clust = j;
    a.cells[si].type = 0;
  }
  const int64_t pb = a.cell_ptr[si], K = a.cell_ptr[si + 1] - pb;
  for (int64_t t = (int64_t)qc * NT + tid; t < K; t += (int64_t)a.Q * NT) {
    const int64_t p = pb + t;
    const int32_t s = a.pair_snp[p];
    const int64_t q = (int64_t)a.pos[p], lo = a.snp_ptr[s], hi = a.snp_ptr[s + 1];
    bool first = true;  // the first of the proven cells that goes into cluster j at this SNP folds all of them, in order
    for (int64_t tt = q; tt > lo; --tt) {
      const uint32_t r = __ldg(a.rk + tt - 1);
      if (r < r0) break;
      if (fin[r - r0] == j) { first = false; break; }
    }
    if (!first) continue;
    const size_t e = (size_t)s * nS + j;
    double* cg = a.clust_gl + e * 9;
    double gl[9], o[9];
#pragma unroll
    for (int g = 0; g < 9; ++g) gl[g] = cg[g];
    for (int64_t tt = q; tt < hi; ++tt) {
      if (tt != q) {
        const uint32_t r = __ldg(a.rk + tt);
        if (r >= r1) break;
        if (fin[r - r0] != j) continue;
      }
      const double* src = a.gl_rk + (size_t)tt * 9;
#pragma unroll
      for (int g = 0; g < 9; ++g) o[g] = __ldg(src + g);
      fmx_merge(gl, o);
    }
#pragma unroll
    for (int g = 0; g < 9; ++g) cg[g] = gl[g];
    a.diag[e * 3] = gl[0]; a.diag[e * 3 + 1] = gl[4]; a.diag[e * 3 + 2] = gl[8];
    a.present[e] = 1;
  }
}

// (SNP, rank) keys of the cell-major pairs, and the sorted list unpacked
__global__ void k_fmx_seed3_keys(const int32_t* __restrict__ pair_snp, const int32_t* __restrict__ pair_cell, const uint32_t* __restrict__ rank_of,
                                 int64_t P, int rank_bits, unsigned long long* __restrict__ key, uint32_t* __restrict__ val) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  key[p] = ((unsigned long long)(uint32_t)pair_snp[p] << rank_bits) | rank_of[pair_cell[p]];
  val[p] = (uint32_t)p;
}
__global__ void k_fmx_seed3_unpack(const unsigned long long* __restrict__ key, const uint32_t* __restrict__ val, const uint32_t* __restrict__ csc_pos,
                                   const double* __restrict__ gl_csc, int64_t P, int rank_bits, uint32_t* __restrict__ rk, uint32_t* __restrict__ pos,
                                   double* __restrict__ gl_rk) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P) return;
  const unsigned long long kk = key[t];
  const uint32_t pp = val[t];
  rk[t] = (uint32_t)(kk & ((1ull << rank_bits) - 1ull));
  pos[pp] = (uint32_t)t;
  const double* src = gl_csc + ((size_t)csc_pos[pp] * 32 + ((size_t)(kk >> rank_bits) & 31)) * 9;  // the M-step's 72-byte record of the pair
#pragma unroll
  for (int g = 0; g < 9; ++g) gl_rk[(size_t)t * 9 + g] = src[g];
}

// ------------------------------------------------------------------------------------------------
// freemuxlet-old seeding: pairwise Bayes factors (cmd_cram_freemuxlet.cpp:174-222) on the device
// ------------------------------------------------------------------------------------------------
// dropDs[i][j] (j < i) accumulates, over the SNPs both droplets cover in ascending SNP order, log lk2 and log lk0 of the
// two droplets' diagonal GLs.  The votes (:245-346) only ever compare llk2 - llk0 with +-bf_thres, so the device keeps two
// bits per pair: 1 = "same donor" (llk2 - llk0 > thres), 2 = "different" (llk0 - llk2 > thres), 0 = undecided.
// One WARP per droplet i walks i's SNPs in ascending order; the 32 lanes take the other droplets of that SNP's list
// (SNP-major view), so every dropDs[i][j] receives its terms in the reference's order; the row's running sums live in a
// per-warp scratch row in global memory (L2).  No C x C matrix of doubles exists anywhere: C^2/4 bytes of trits.
struct PairwiseArgs {
  const int64_t* cell_ptr;
  const int32_t* pair_snp;
  const double* gl_soa;
  const double* snp_af;
  const int64_t* snp_ptr;
  const uint32_t* snp_pair;
  const int32_t* pair_cell;
  double* scratch;   // [warps][2][C]
  uint32_t* trit;    // [C][W] 16 two-bit entries per word; row i holds j < i after this kernel
  int* counter;
  int64_t P;
  int32_t C, W;
  double thres;
};
__global__ void __launch_bounds__(512) k_fmx_pairwise(PairwiseArgs a) {
  const int lane = threadIdx.x & 31;
  const int gw = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  double* const acc2 = a.scratch + (size_t)gw * 2 * a.C;
  double* const acc0 = acc2 + a.C;
  for (;;) {
    int w = 0;
    if (lane == 0) w = atomicAdd(a.counter, 1);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= a.C) break;
    const int i = a.C - 1 - w;  // long rows first
    for (int j = lane; j < i; j += 32) { acc2[j] = 0.0; acc0[j] = 0.0; }
    __syncwarp();
    for (int64_t p = a.cell_ptr[i]; p < a.cell_ptr[i + 1]; ++p) {
      const int32_t s = a.pair_snp[p];
      const double af = a.snp_af[s];
      double gps[3];
      gps[0] = __dmul_rn(1.0 - af, 1.0 - af); gps[1] = __dmul_rn(__dmul_rn(2.0, af), 1.0 - af); gps[2] = __dmul_rn(af, af);
      const double gi[3] = {a.gl_soa[p], a.gl_soa[(size_t)4 * a.P + p], a.gl_soa[(size_t)8 * a.P + p]};
      const int64_t e = a.snp_ptr[s + 1];
      for (int64_t t = a.snp_ptr[s] + lane; t < e; t += 32) {
        const uint32_t q = a.snp_pair[t];
        const int j = a.pair_cell[q];
        if (j >= i) break;  // the list is in ascending droplet order (:196 `jt != it`)
        const double gj[3] = {a.gl_soa[q], a.gl_soa[(size_t)4 * a.P + q], a.gl_soa[(size_t)8 * a.P + q]};
        double lk0 = 0.0, lk2 = 0.0;
#pragma unroll
        for (int g = 0; g < 3; ++g) {  // :204-209, the reference's order of operations, no contraction
          lk2 = __dadd_rn(lk2, __dmul_rn(__dmul_rn(gi[g], gj[g]), gps[g]));
#pragma unroll
          for (int h = 0; h < 3; ++h) lk0 = __dadd_rn(lk0, __dmul_rn(__dmul_rn(__dmul_rn(gi[g], gj[h]), gps[g]), gps[h]));
        }
        acc2[j] += log(lk2);
        acc0[j] += log(lk0);
      }
      __syncwarp();  // the next SNP may hand droplet j to another lane
    }
    uint32_t* row = a.trit + (size_t)i * a.W;
    for (int wd = lane; wd * 16 < i; wd += 32) {
      uint32_t bits = 0;
      for (int k = 0; k < 16; ++k) {
        const int j = wd * 16 + k;
        if (j < i) {
          const double l2 = acc2[j], l0 = acc0[j];
          const uint32_t t = (l0 - l2 > a.thres) ? 2u : (l2 - l0 > a.thres) ? 1u : 0u;  // :275-280
          bits |= t << (2 * k);
        }
      }
      row[wd] = bits;
    }
    __syncwarp();
  }
}
// fills the upper triangle from the lower one (row i, j > i  <-  row j, column i)
__global__ void k_fmx_trit_symmetrize(uint32_t* trit, int32_t C, int32_t W) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)C * W) return;
  const int i = (int)(idx / W), wd = (int)(idx - (int64_t)i * W);
  if (wd * 16 + 15 <= i) return;  // entirely in the lower triangle
  uint32_t bits = trit[idx];
  for (int k = 0; k < 16; ++k) {
    const int j = wd * 16 + k;
    if (j > i && j < C) {
      const uint32_t t = (trit[(size_t)j * W + (i >> 4)] >> (2 * (i & 15))) & 3u;
      bits = (bits & ~(3u << (2 * k))) | (t << (2 * k));
    } else if (j >= i) bits &= ~(3u << (2 * k));
  }
  trit[idx] = bits;
}
__global__ void k_fmx_pair_cell(const int64_t* __restrict__ cell_ptr, int32_t C, int64_t P, int32_t* __restrict__ pair_cell) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  int lo = 0, hi = C;  // largest c with cell_ptr[c] <= p
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cell_ptr[mid] <= p) lo = mid; else hi = mid;
  }
  pair_cell[p] = lo;
}
// clusters decided on the host -> cell records, clust[]
__global__ void k_fmx_set_clusters(const int32_t* __restrict__ cl, int32_t C, pscl_fmx_cell* __restrict__ cells, int32_t* __restrict__ clust) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int v = cl[c];
  cells[c].clust = cells[c].init_clust = v;
  cells[c].type = v >= 0 ? 0 : -1;
  clust[c] = v;
}

__global__ void k_fmx_fill_f64(double* v, size_t n, double x) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = x;
}

// ------------------------------------------------------------------------------------------------
// M-step: cluster pileups from scratch (cmd_cram_freemux2.cpp:277-288, :516-517 + :590-596)
// ------------------------------------------------------------------------------------------------
struct MArgs {
  const int64_t* snp_ptr;
  const int64_t* grp_base;
  const int32_t* cell_w;
  const uint32_t* snp_pair;
  const double* gl_csc;
  const int32_t* member;
  const uint32_t* pair_rd;
  const uint8_t* rd_aq;
  double* clust_diag;
  double* clust_gl;    // written when final
  int32_t* clust_cnt;  // written when final
  int32_t V, nS, final_;
};

// One thread per SNP (a warp = one group of 32 consecutive SNPs) walks the SNP's cell list once (ascending
// cell id: the clamp makes the merge order matter) and merges every singlet into the pileup of ITS cluster;
// the nS pileups of the SNP live in a shared-memory column of the thread ([cluster*9 + g][thread],
// conflict-free).  Entry t of the 32 lists is one coalesced row of cell ids and one contiguous 2304-byte block of
// 32 GL records (warp-grouped layout, k_fmx_csc_fill; records stay whole so that stage 1 writes 72 contiguous
// bytes per pair); the loads of entry t+1 run ahead of the merge of entry t.
// History (config 3, per EM iteration): thread per (SNP, cluster) scanning the whole list 1.08 ms; thread
// per SNP on the [P][9] record layout 1.08 ms (same dependent-load chain), 0.86 ms with the loads run ahead.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_fmx_mstep(MArgs a) {
  extern __shared__ __align__(16) double s_gl[];  // [nS*9][THREADS], then int [nS*3][THREADS] when final
  int* const s_cnt = reinterpret_cast<int*>(s_gl + (size_t)a.nS * 9 * THREADS);
  const int tid = threadIdx.x, lane = tid & 31;
  const int v = blockIdx.x * THREADS + tid;
  const int group = v >> 5;
  const int nS = a.nS;
  for (int q = 0; q < nS * 9; ++q) s_gl[q * THREADS + tid] = 1.0;  // default-constructed pileup (sc_drop_seq.h:72-75)
  if (a.final_)
    for (int q = 0; q < nS * 3; ++q) s_cnt[q * THREADS + tid] = 0;
  if (group * 32 >= a.V) return;  // whole warp out of range
  const int64_t slot0 = a.grp_base[group], L = a.grp_base[group + 1] - slot0;  // the group's longest list
  const int64_t b = v < a.V ? a.snp_ptr[v] : 0;
  int cell_n = -1, j_n = -1;
  double o_n[9];
  if (L > 0) {
    cell_n = a.cell_w[slot0 * 32 + lane];
    j_n = cell_n >= 0 ? a.member[cell_n] : -1;
#pragma unroll
    for (int g = 0; g < 9; ++g) o_n[g] = a.gl_csc[((size_t)slot0 * 32 + lane) * 9 + g];
  }
  int cell_n2 = L > 1 ? a.cell_w[(slot0 + 1) * 32 + lane] : -1;
  for (int64_t t = 0; t < L; ++t) {
    const int j = j_n;
    double o[9];
#pragma unroll
    for (int g = 0; g < 9; ++g) o[g] = o_n[g];
    if (t + 1 < L) {
      j_n = cell_n2 >= 0 ? a.member[cell_n2] : -1;
      const double* src = a.gl_csc + ((size_t)(slot0 + t + 1) * 32 + lane) * 9;
#pragma unroll
      for (int g = 0; g < 9; ++g) o_n[g] = src[g];
      cell_n2 = (t + 2 < L) ? a.cell_w[(slot0 + t + 2) * 32 + lane] : -1;
    }
    if (j < 0) continue;
    double gl[9];
    double* col = s_gl + (size_t)j * 9 * THREADS + tid;
#pragma unroll
    for (int g = 0; g < 9; ++g) gl[g] = col[g * THREADS];
    fmx_merge(gl, o);
#pragma unroll
    for (int g = 0; g < 9; ++g) col[g * THREADS] = gl[g];
    if (a.final_) {
      const uint32_t p = a.snp_pair[b + t];
      int nreads = 0, nref = 0, nalt = 0;
      for (uint32_t r = a.pair_rd[p]; r < a.pair_rd[p + 1]; ++r) {
        const uint32_t al = a.rd_aq[r] >> 6;
        ++nreads;
        if (al == 0) ++nref; else if (al == 1) ++nalt;
      }
      int* cc = s_cnt + (size_t)j * 3 * THREADS + tid;
      cc[0] += nreads; cc[THREADS] += nref; cc[2 * THREADS] += nalt;
    }
  }
  if (v >= a.V) return;
  for (int j = 0; j < nS; ++j) {
    const double* col = s_gl + (size_t)j * 9 * THREADS + tid;
    const size_t e = (size_t)v * nS + j;
    double* d = a.clust_diag + e * 3;
    d[0] = col[0]; d[1] = col[4 * THREADS]; d[2] = col[8 * THREADS];
    if (a.final_) {
      double* cg = a.clust_gl + e * 9;
#pragma unroll
      for (int g = 0; g < 9; ++g) cg[g] = col[g * THREADS];
      int32_t* cc = a.clust_cnt + e * 3;
      const int* sc = s_cnt + (size_t)j * 3 * THREADS + tid;
      cc[0] = sc[0]; cc[1] = sc[THREADS]; cc[2] = sc[2 * THREADS];
    }
  }
}

// posterior of each (SNP, cluster) (cmd_cram_freemux2.cpp:388-390, :402-415)
__global__ void k_fmx_posterior(const double* __restrict__ clust_diag, const double* __restrict__ snp_af, int32_t V,
                                int32_t nS, int32_t US, double geno_error, int apply_err, double* __restrict__ u_tab) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)V * nS) return;
  const int v = (int)(t / nS), j = (int)(t - (int64_t)v * nS);
  const double af = snp_af[v];
  const double h0 = (1.0 - af) * (1.0 - af), h1 = 2 * af * (1.0 - af), h2 = af * af;
  const double* d = clust_diag + (size_t)t * 3;
  double u0 = h0 * d[0], u1 = h1 * d[1], u2 = h2 * d[2];
  const double s = u0 + u1 + u2;
  u0 /= s; u1 /= s; u2 /= s;
  if (apply_err) {
    u0 = (1 - geno_error) * u0 + geno_error * h0;
    u1 = (1 - geno_error) * u1 + geno_error * h1;
    u2 = (1 - geno_error) * u2 + geno_error * h2;
  }
  double* u = u_tab + (size_t)v * US + j * 3;
  u[0] = u0; u[1] = u1; u[2] = u2;
}

// ------------------------------------------------------------------------------------------------
// E-step (cmd_cram_freemux2.cpp:383-456)
// ------------------------------------------------------------------------------------------------
// One warp per work item (<= 2048 pairs of one cell), one lane per (cell,SNP) pair.  A kernel
// instance owns the cluster rows j in [J0, J1) (all k <= j): NA running products per lane, kept as
// mantissa (register) + exponent (shared memory), one log per accumulator per item.
//   lk[j][k] = sum_{a,b} gl[a][b] u_j[a] u_k[b] = u_j . (GL u_k)      (:443, k < j)
//   lk[j][j] = sum_a gl[a][a] u_j[a]                                    (:450)
// NS > 0: the cluster count is a compile-time constant (J0 = 0, J1 = NS); NS == 0: runtime a.nS.
struct EArgs {
  const int32_t* pair_snp;
  const double* gl_soa;
  const double* u_tab;
  const int32_t* item_order;
  const int64_t* item_pbeg;
  const int64_t* item_pend;
  double* item_llk;   // [n_items][npairs]
  int* counter;
  int64_t P;
  int32_t n_items, npairs;
  int32_t nS, US, SD;  // SD: shared-memory row stride in doubles (16-byte units odd -> conflict-free)
};

template <int NS, int J0, int J1, int THREADS>
__global__ void __launch_bounds__(THREADS) k_fmx_estep(EArgs a) {
  constexpr int NR = J1 - J0;
  constexpr int NA = J1 * (J1 + 1) / 2 - J0 * (J0 + 1) / 2;
  constexpr int EBASE = J0 * (J0 + 1) / 2;
  const int nS = NS > 0 ? NS : a.nS;
  const int US = NS > 0 ? ((NS * 3 + 1) & ~1) : a.US;
  // a kernel instance with compile-time NS only needs the posteriors of clusters [0, J1): it gathers and stages that
  // head of the row (RD doubles), which shrinks both the L2 traffic and the shared-memory footprint of the early tiles
  const int RD = NS > 0 ? ((J1 * 3 + 1) & ~1) : US;
  const int SD = NS > 0 ? ((((RD / 2) & 1) == 0) ? RD + 2 : RD) : a.SD;
  const int NCH = RD / 2;                       // 16-byte pieces per (head of a) posterior row
  int lg = 0;
  while ((1 << lg) < NCH && lg < 5) ++lg;       // lanes per row (power of two, <= 32)
  const int LPR = 1 << lg, RPI = 32 >> lg;
  const int PPL = (NCH + LPR - 1) >> lg;        // pieces per lane (1, or 2 when NCH > 32)

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_u = reinterpret_cast<double*>(smem_raw);                         // [2][THREADS][SD]
  int* s_exp = reinterpret_cast<int*>(s_u + (size_t)2 * THREADS * SD);       // [NA][THREADS]
  const int tid = threadIdx.x, lane = tid & 31;
  double* const row0 = s_u + (size_t)tid * SD;
  const size_t buf_d = (size_t)THREADS * SD;

  for (;;) {
    int w = 0;
    if (lane == 0) w = atomicAdd(a.counter, 1);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= a.n_items) break;
    const int item = a.item_order[w];
    const int64_t pb = a.item_pbeg[item], pe = a.item_pend[item];
    const int niter = (int)((pe - pb + 31) >> 5);

    double acc[NA];
#pragma unroll
    for (int e = 0; e < NA; ++e) { acc[e] = 1.0; s_exp[e * THREADS + tid] = 0; }

    // software pipeline: SNP id two iterations ahead; posterior row (cp.async) and the pair's 9 GLs
    // (registers) one iteration ahead
    int32_t snpA = 0; bool okA = false;
    double glB[9]; bool okB = false;
    auto loadA = [&](int it) {
      const int64_t p = pb + ((int64_t)it << 5) + lane;
      okA = (it < niter) && (p < pe);
      if (okA) snpA = a.pair_snp[p];
    };
    auto issueB = [&](int it, int buf) {  // consumes stage A (iteration `it`)
      okB = okA;
      if (okA) {
        const int64_t p = pb + ((int64_t)it << 5) + lane;
#pragma unroll
        for (int g = 0; g < 9; ++g) glB[g] = a.gl_soa[(size_t)g * a.P + p];
      }
      const unsigned okmask = __ballot_sync(0xffffffffu, okA);
      double* dst_base = s_u + (size_t)buf * buf_d + (size_t)(tid & ~31) * SD;
      for (int i = 0; i < LPR; ++i) {
        const int row = i * RPI + (lane >> lg), piece = lane & (LPR - 1);
        const int snp_r = __shfl_sync(0xffffffffu, snpA, row);
        if ((okmask >> row) & 1u) {
          for (int pp = 0; pp < PPL; ++pp) {
            const int pc = piece + (pp << lg);
            if (pc < NCH)
              __pipeline_memcpy_async(reinterpret_cast<char*>(dst_base + (size_t)row * SD) + pc * 16,
                                      reinterpret_cast<const char*>(a.u_tab + (size_t)snp_r * US) + pc * 16, 16);
          }
        }
      }
      __pipeline_commit();
    };
    loadA(0);
    issueB(0, 0);
    loadA(1);

    for (int it = 0; it < niter; ++it) {
      const int buf = it & 1;
      const bool ok = okB;
      double gl[9];
#pragma unroll
      for (int g = 0; g < 9; ++g) gl[g] = glB[g];
      __syncwarp();  // all lanes are done with buffer buf^1 before it is refilled
      issueB(it + 1, buf ^ 1);
      loadA(it + 2);
      __pipeline_wait_prior(1);
      __syncwarp();

      if (ok) {
        const double* row = row0 + (size_t)buf * buf_d;
        double uj[NR][3];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          if (J0 + r < nS) { uj[r][0] = row[(J0 + r) * 3]; uj[r][1] = row[(J0 + r) * 3 + 1]; uj[r][2] = row[(J0 + r) * 3 + 2]; }
          else { uj[r][0] = uj[r][1] = uj[r][2] = 0.0; }
        }
#pragma unroll
        for (int k = 0; k < J1; ++k) {
          if (k < nS) {
            double k0, k1, k2;
            if (k >= J0) { k0 = uj[k - J0 < 0 ? 0 : k - J0][0]; k1 = uj[k - J0 < 0 ? 0 : k - J0][1]; k2 = uj[k - J0 < 0 ? 0 : k - J0][2]; }
            else { k0 = row[k * 3]; k1 = row[k * 3 + 1]; k2 = row[k * 3 + 2]; }
            const double v0 = gl[0] * k0 + gl[1] * k1 + gl[2] * k2;
            const double v1 = gl[3] * k0 + gl[4] * k1 + gl[5] * k2;
            const double v2 = gl[6] * k0 + gl[7] * k1 + gl[8] * k2;
#pragma unroll
            for (int j = (k + 1 > J0 ? k + 1 : J0); j < J1; ++j)
              if (j < nS) acc[j * (j + 1) / 2 + k - EBASE] *= (uj[j - J0][0] * v0 + uj[j - J0][1] * v1 + uj[j - J0][2] * v2);
            if (k >= J0) acc[k * (k + 1) / 2 + k - EBASE] *= (gl[0] * k0 + gl[4] * k1 + gl[8] * k2);
          }
        }
      }
      if ((it & 31) == 31) {  // every factor is >= ~1e-6 (clamped GLs), so 32 of them stay far above underflow
#pragma unroll
        for (int e = 0; e < NA; ++e) { int ex = 0; pscl_renorm(acc[e], ex); s_exp[e * THREADS + tid] += ex; }
      }
    }
    __pipeline_wait_prior(0);
    __syncwarp();

    // ---- item epilogue: product across the warp, one log per accumulator, taken by lane e % 32 ----
    // (rolled through this lane's now idle posterior rows: 2*SD >= NA doubles, checked on the host)
#pragma unroll
    for (int e = 0; e < NA; ++e) { if (e < SD) row0[e] = acc[e]; else row0[buf_d + (e - SD)] = acc[e]; }
    double keep_m0 = 1.0, keep_m1 = 1.0;
    int keep_e0 = 0, keep_e1 = 0;
#pragma unroll 1
    for (int e = 0; e < NA; ++e) {
      double m = (e < SD) ? row0[e] : row0[buf_d + (e - SD)];
      int ex = s_exp[e * THREADS + tid];
      pscl_renorm(m, ex);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        m *= __shfl_xor_sync(0xffffffffu, m, o);
        ex += __shfl_xor_sync(0xffffffffu, ex, o);
      }
      if ((e & 31) == lane) {
        if (e < 32) { keep_m0 = m; keep_e0 = ex; } else { keep_m1 = m; keep_e1 = ex; }
      }
    }
    double* out = a.item_llk + (size_t)item * a.npairs;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int e = lane + 32 * half;
      if (e < NA && e + EBASE < a.npairs) out[e + EBASE] = pscl_prod_log(half ? keep_m1 : keep_m0, half ? keep_e1 : keep_e0);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// E-step at nS = 16, one pass: a team of four warps per work item, one warp per row tile
// ------------------------------------------------------------------------------------------------
// The four row tiles [0,8) [8,11) [11,14) [14,16) used to be four launches, each gathering (the head of) the posterior row of
// every pair again: 1.2 KB per pair, and at 500 k SNPs the 192 MB table does not fit L2, so that was DRAM traffic (48.6 GB
// per iteration at 4.8e7 pairs, profiles/r2j_k_fmx_estep16_tile*_ncu.txt).  Here the four warps of a CTA take the SAME 32
// pairs per step: the 32 rows (384 B each) are gathered once into shared memory, 8 rows per warp with cp.async, double
// buffered; every warp then runs its own tile's products out of those rows (warp-uniform code, the tile is chosen by the
// warp's index) — the same arithmetic on the same doubles in the same order as the tile kernels, so the same bits.
template <int J0, int J1>
__device__ __forceinline__ void estep16_tile(const double* __restrict__ row, const double (&gl)[9], double (&acc)[40]) {
  constexpr int NR = J1 - J0;
  constexpr int EBASE = J0 * (J0 + 1) / 2;
  double uj[NR][3];
#pragma unroll
  for (int r = 0; r < NR; ++r) { uj[r][0] = row[(J0 + r) * 3]; uj[r][1] = row[(J0 + r) * 3 + 1]; uj[r][2] = row[(J0 + r) * 3 + 2]; }
#pragma unroll
  for (int k = 0; k < J1; ++k) {
    double k0, k1, k2;
    if (k >= J0) { k0 = uj[k - J0 < 0 ? 0 : k - J0][0]; k1 = uj[k - J0 < 0 ? 0 : k - J0][1]; k2 = uj[k - J0 < 0 ? 0 : k - J0][2]; }
    else { k0 = row[k * 3]; k1 = row[k * 3 + 1]; k2 = row[k * 3 + 2]; }
    const double v0 = gl[0] * k0 + gl[1] * k1 + gl[2] * k2;
    const double v1 = gl[3] * k0 + gl[4] * k1 + gl[5] * k2;
    const double v2 = gl[6] * k0 + gl[7] * k1 + gl[8] * k2;
#pragma unroll
    for (int j = (k + 1 > J0 ? k + 1 : J0); j < J1; ++j)
      acc[j * (j + 1) / 2 + k - EBASE] *= (uj[j - J0][0] * v0 + uj[j - J0][1] * v1 + uj[j - J0][2] * v2);
    if (k >= J0) acc[k * (k + 1) / 2 + k - EBASE] *= (gl[0] * k0 + gl[4] * k1 + gl[8] * k2);
  }
}
template <int J0, int J1>
__device__ __forceinline__ void estep16_renorm(double (&acc)[40], int* s_exp, int tid) {
  constexpr int NA = J1 * (J1 + 1) / 2 - J0 * (J0 + 1) / 2;
#pragma unroll
  for (int e = 0; e < NA; ++e) { int ex = 0; pscl_renorm(acc[e], ex); s_exp[e * 128 + tid] += ex; }
}
// product across the warp of every accumulator of the tile (transpose-reduction from registers, as k_demux_default's item
// epilogue), one log per accumulator, stored by the lane that ends up holding it
template <int J0, int J1>
__device__ __forceinline__ void estep16_store(const double (&acc)[40], const int* s_exp, int tid, double* __restrict__ out, int npairs) {
  constexpr int NA = J1 * (J1 + 1) / 2 - J0 * (J0 + 1) / 2;
  constexpr int EBASE = J0 * (J0 + 1) / 2;
  static_assert(NA <= 40, "tile too large for the two epilogue groups");
  const int lane = tid & 31;
  double m1[32]; int x1[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    if (e < NA) { m1[e] = acc[e]; x1[e] = s_exp[e * 128 + tid]; pscl_renorm(m1[e], x1[e]); } else { m1[e] = 1.0; x1[e] = 0; }
  }
  pscl_transpose_prod<32>(m1, x1, lane);
  pscl_renorm(m1[0], x1[0]);
  if (lane < NA && lane + EBASE < npairs) out[lane + EBASE] = pscl_prod_log(m1[0], x1[0]);
  if (NA > 32) {
    double m2[8]; int x2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (32 + e < NA) { m2[e] = acc[32 + e]; x2[e] = s_exp[(32 + e) * 128 + tid]; pscl_renorm(m2[e], x2[e]); } else { m2[e] = 1.0; x2[e] = 0; }
    }
    pscl_transpose_prod<8>(m2, x2, lane);
    pscl_renorm(m2[0], x2[0]);
    const int e = 32 + (lane >> 2);  // element 32 + i sits in lane 4 * i
    if ((lane & 3) == 0 && e < NA && e + EBASE < npairs) out[e + EBASE] = pscl_prod_log(m2[0], x2[0]);
  }
}

__global__ void __launch_bounds__(128, 2) k_fmx_estep16_team(EArgs a) {
  constexpr int US = 48, SD = 50, NCH = 24;  // posterior row: 16 clusters x 3 doubles = 24 pieces of 16 bytes; odd 16-byte stride
  constexpr int NBUF = 3;                    // row buffers: the gathers run two steps ahead of the products
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_u = reinterpret_cast<double*>(smem_raw);                    // [NBUF][32][SD] rows of a step's 32 pairs
  int* s_exp = reinterpret_cast<int*>(s_u + (size_t)NBUF * 32 * SD);    // [40][128]
  int* s_snp = s_exp + 40 * 128;                                        // [2][1024] SNP ids of 32 steps, two chunks
  __shared__ int s_item;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (;;) {
    if (tid == 0) s_item = atomicAdd(a.counter, 1);
    __syncthreads();
    const int w = s_item;
    __syncthreads();
    if (w >= a.n_items) break;
    const int item = a.item_order[w];
    const int64_t pb = a.item_pbeg[item], pe = a.item_pend[item];
    const int niter = (int)((pe - pb + 31) >> 5);
    double acc[40];
#pragma unroll
    for (int e = 0; e < 40; ++e) { acc[e] = 1.0; s_exp[e * 128 + tid] = 0; }
    // SNP ids: 32 steps (1024 pairs) per chunk, loaded by the whole team a chunk ahead (-1 = past the item's end)
    auto load_chunk = [&](int c) {
      int* dst = s_snp + (c & 1) * 1024;
      const int64_t base = pb + (int64_t)c * 1024;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int64_t p = base + tid + 128 * q;
        dst[tid + 128 * q] = p < pe ? a.pair_snp[p] : -1;
      }
    };
    auto snp_at = [&](int step, int ln) { return s_snp[((step >> 5) & 1) * 1024 + (step & 31) * 32 + ln]; };
    auto issue_rows = [&](int step, int buf) {  // this warp's 8 rows of the step's 32
      if (step < niter) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = warp * 8 + i;
          const int snp_r = snp_at(step, r);
          if (snp_r >= 0 && lane < NCH)
            __pipeline_memcpy_async(reinterpret_cast<char*>(s_u + ((size_t)buf * 32 + r) * SD) + lane * 16,
                                    reinterpret_cast<const char*>(a.u_tab + (size_t)snp_r * US) + lane * 16, 16);
        }
      }
      __pipeline_commit();
    };
    double glB[9]; bool okB = false;
    auto load_gl = [&](int step) {
      const int64_t p = pb + ((int64_t)step << 5) + lane;
      okB = step < niter && p < pe;
      if (okB) {
#pragma unroll
        for (int g = 0; g < 9; ++g) glB[g] = a.gl_soa[(size_t)g * a.P + p];
      }
    };
    load_chunk(0);
    __syncthreads();
    issue_rows(0, 0);
    issue_rows(1, 1);
    load_gl(0);
    for (int it = 0; it < niter; ++it) {
      const bool ok = okB;
      double gl[9];
#pragma unroll
      for (int g = 0; g < 9; ++g) gl[g] = glB[g];
      __syncthreads();  // every warp is done with the buffer of step it-1 (refilled now) and with the SNP chunk before this one
      if ((it & 31) == 0) load_chunk((it >> 5) + 1);
      issue_rows(it + 2, (it + 2) % NBUF);
      load_gl(it + 1);
      __pipeline_wait_prior(2);
      __syncthreads();  // ... and every warp's share of this step's rows has landed (and the new SNP chunk is visible)
      if (ok) {
        const double* row = s_u + ((size_t)(it % NBUF) * 32 + lane) * SD;
        switch (warp) {
          case 0: estep16_tile<0, 8>(row, gl, acc); break;
          case 1: estep16_tile<8, 11>(row, gl, acc); break;
          case 2: estep16_tile<11, 14>(row, gl, acc); break;
          default: estep16_tile<14, 16>(row, gl, acc); break;
        }
      }
      if ((it & 31) == 31) {  // every factor is >= ~1e-6 (clamped GLs), so 32 of them stay far above underflow
        switch (warp) {
          case 0: estep16_renorm<0, 8>(acc, s_exp, tid); break;
          case 1: estep16_renorm<8, 11>(acc, s_exp, tid); break;
          case 2: estep16_renorm<11, 14>(acc, s_exp, tid); break;
          default: estep16_renorm<14, 16>(acc, s_exp, tid); break;
        }
      }
    }
    __pipeline_wait_prior(0);
    double* out = a.item_llk + (size_t)item * a.npairs;
    switch (warp) {
      case 0: estep16_store<0, 8>(acc, s_exp, tid, out, a.npairs); break;
      case 1: estep16_store<8, 11>(acc, s_exp, tid, out, a.npairs); break;
      case 2: estep16_store<11, 14>(acc, s_exp, tid, out, a.npairs); break;
      default: estep16_store<14, 16>(acc, s_exp, tid, out, a.npairs); break;
    }
  }
}

// llk[c][pair] = sum over the cell's items, in item order (a cell without pairs keeps 0, :385 init)
__global__ void k_fmx_llk_reduce(const int32_t* __restrict__ cell_item_ptr, const double* __restrict__ item_llk,
                                 int32_t C, int32_t npairs, double* __restrict__ llk) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)C * npairs) return;
  const int c = (int)(t / npairs), q = (int)(t - (int64_t)c * npairs);
  double x = 0.0;
  for (int it = cell_item_ptr[c]; it < cell_item_ptr[c + 1]; ++it) x += item_llk[(size_t)it * npairs + q];
  llk[t] = x;
}

// ------------------------------------------------------------------------------------------------
// per-cell epilogue + classify (cmd_cram_freemux2.cpp:458-584; old mode cmd_cram_freemuxlet.cpp:524-648)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double fmx_log_add(double la, double lb) {  // sc_drop_seq.cpp:5-8
  return (la > lb) ? la + log(1.0 + exp(lb - la)) : lb + log(1.0 + exp(la - lb));
}

__global__ void __launch_bounds__(128) k_fmx_classify(const double* __restrict__ llk, int32_t C, int32_t nS,
                                                      double doublet_prior, int mode_old, pscl_fmx_cell* __restrict__ cells,
                                                      int32_t* __restrict__ clust, int32_t* __restrict__ member,
                                                      int32_t* __restrict__ counters) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int npairs = nS * (nS + 1) / 2;
  const double* llks = llk + (size_t)c * npairs;
  const double lsp = log((1.0 - doublet_prior) / nS);            // :379
  const double ldp = log(doublet_prior / nS / (nS - 1) * 2.0);   // :380
  int sBest = -1, sNext = -1, dBest1 = -1, dBest2 = -1, dNext1 = -1, dNext2 = -1;
  double sngBestLLK = -1e300, sngNextLLK = -1e300, dblBestLLK = -1e300, dblNextLLK = -1e300;
  double sumLLK = -1e300, sngLLK = -1e300;
  for (int j = 0; j < nS; ++j) {  // :468-497, same scan order
    for (int k = 0; k < j; ++k) {
      const double x = llks[j * (j + 1) / 2 + k];
      if (x > dblBestLLK) { dNext1 = dBest1; dNext2 = dBest2; dblNextLLK = dblBestLLK; dBest1 = j; dBest2 = k; dblBestLLK = x; }
      else if (x > dblNextLLK) { dNext1 = j; dNext2 = k; dblNextLLK = x; }
      sumLLK = fmx_log_add(sumLLK, x + ldp);
    }
    const double x = llks[j * (j + 1) / 2 + j];
    if (x > sngBestLLK) { sNext = sBest; sngNextLLK = sngBestLLK; sBest = j; sngBestLLK = x; }
    else if (x > sngNextLLK) { sNext = j; sngNextLLK = x; }
    sumLLK = fmx_log_add(sumLLK, x + lsp);
    sngLLK = fmx_log_add(sngLLK, x + lsp);
  }
  pscl_fmx_cell r = cells[c];
  const int prev_type = r.type;
  r.sng_best = sBest; r.sng_best_llk = sngBestLLK; r.sng_next = sNext; r.sng_next_llk = sngNextLLK;
  r.dbl_best_j = dBest1; r.dbl_best_k = dBest2; r.dbl_best_llk = dblBestLLK;
  r.dbl_next_j = dNext1; r.dbl_next_k = dNext2; r.dbl_next_llk = dblNextLLK;
  r.sng_pp = exp(sngLLK - sumLLK);            // :510
  r.sng_only_pp = exp(sngBestLLK + lsp - sngLLK);  // :511
  r.sum_llk = sumLLK;
  int cl = mode_old ? r.clust : -1;  // :520 resets clusts; freemuxlet-old never touches it
  bool changed;
  if (dblBestLLK > sngBestLLK + 2) {  // :521-543
    changed = prev_type != 1;
    r.type = 1;
    r.best_pp = dblBestLLK + ldp - sumLLK;  // kept as a log (:526)
    r.best_j = dBest1; r.best_k = dBest2; r.best_llk = dblBestLLK;
    if (dblNextLLK > sngBestLLK + 2) { r.next_j = dNext1; r.next_k = dNext2; r.next_llk = dblNextLLK; }
    else { r.next_j = r.next_k = sBest; r.next_llk = sngBestLLK; }
  } else if (sngBestLLK > sngNextLLK + 2) {  // :544-565
    changed = (prev_type != 0) || (r.best_j != sBest) || (r.best_k != sBest);
    r.type = 0;
    atomicAdd(&counters[1], 1);
    r.best_pp = sngBestLLK + lsp - sumLLK;
    r.best_j = r.best_k = sBest; r.best_llk = sngBestLLK;
    if (!mode_old) cl = sBest;  // :553
    if (dblBestLLK > sngNextLLK + 2) { r.next_j = dBest1; r.next_k = dBest2; r.next_llk = dblBestLLK; }
    else { r.next_j = r.next_k = sNext; r.next_llk = sngNextLLK; }
  } else {  // :566-584
    changed = prev_type != 2;
    r.type = 2;
    atomicAdd(&counters[2], 1);
    r.best_pp = sngBestLLK + lsp - sumLLK;
    r.best_j = r.best_k = sBest; r.best_llk = sngBestLLK;
    if (dblBestLLK > sngNextLLK + 2) { r.next_j = dBest1; r.next_k = dBest2; r.next_llk = dblNextLLK; /* sic :578 */ }
    else { r.next_j = r.next_k = sNext; r.next_llk = sngNextLLK; }
  }
  if (changed) atomicAdd(&counters[0], 1);
  r.clust = cl;
  cells[c] = r;
  clust[c] = cl;
  member[c] = (r.type == 0 && r.best_j == r.best_k) ? r.best_j : -1;  // who the M-step merges (:592-594)
}

// ------------------------------------------------------------------------------------------------
// host API
// ------------------------------------------------------------------------------------------------
#define FMX_GRID(n, b) (unsigned)(((n) + (b)-1) / (b))

static int fmx_build_csc(pscl_ctx* ctx, pscl_fmx_state* s) {
  const pscl_plp* plp = s->plp;
  const int64_t P = s->P;
  const int32_t NG = (s->V + 31) / 32;
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->snp_ptr, sizeof(int64_t) * ((size_t)s->V + 1)));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->snp_pair, sizeof(uint32_t) * (size_t)(P ? P : 1)));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->csc_pos, sizeof(uint32_t) * (size_t)(P ? P : 1)));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->grp_base, sizeof(int64_t) * ((size_t)NG + 1)));
  s->n_slots = 0;
  if (P == 0) {
    k_fmx_fill_i64<<<FMX_GRID(s->V + 1, 256), 256, 0, ctx->stream>>>(s->snp_ptr, (int64_t)s->V + 1, 0);
    k_fmx_fill_i64<<<FMX_GRID(NG + 1, 256), 256, 0, ctx->stream>>>(s->grp_base, (int64_t)NG + 1, 0);
    ctx->launches += 2;
    PSCL_CUDA(ctx, cudaGetLastError());
    PSCL_CUDA(ctx, cudaMalloc((void**)&s->cell_w, 16));
    return PSCL_OK;
  }
  int32_t* key_out = nullptr;
  uint32_t* val_in = nullptr;
  int64_t* glen = nullptr;
  void *tmp = nullptr, *tmp2 = nullptr;
  size_t tmp_bytes = 0, tmp2_bytes = 0;
  PSCL_CUDA(ctx, cudaMalloc((void**)&key_out, sizeof(int32_t) * (size_t)P));
  PSCL_CUDA(ctx, cudaMalloc((void**)&val_in, sizeof(uint32_t) * (size_t)P));
  PSCL_CUDA(ctx, cudaMalloc((void**)&glen, sizeof(int64_t) * ((size_t)NG + 1)));
  k_fmx_iota<<<FMX_GRID(P, 256), 256, 0, ctx->stream>>>(val_in, P);
  ctx->launches++;
  int end_bit = 1;
  while (end_bit < 31 && ((int64_t)1 << end_bit) < (int64_t)s->V) ++end_bit;
  // stable LSD radix sort: equal SNP ids keep their cell-major order = ascending cell id
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, plp->pair_snp, key_out, val_in, s->snp_pair, (int64_t)P, 0, end_bit, ctx->stream);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp2_bytes, glen, s->grp_base, NG + 1, ctx->stream);
  PSCL_CUDA(ctx, cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
  PSCL_CUDA(ctx, cudaMalloc(&tmp2, tmp2_bytes ? tmp2_bytes : 16));
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, plp->pair_snp, key_out, val_in, s->snp_pair, (int64_t)P, 0, end_bit, ctx->stream);
  ctx->launches += 4;
  if (e == cudaSuccess) {
    k_fmx_snp_ptr<<<FMX_GRID(P, 256), 256, 0, ctx->stream>>>(key_out, s->V, P, s->snp_ptr);
    k_fmx_group_len<<<FMX_GRID(NG + 1, 256), 256, 0, ctx->stream>>>(s->snp_ptr, s->V, NG, glen);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp2, tmp2_bytes, glen, s->grp_base, NG + 1, ctx->stream);  // grp_base[NG] = slots
  ctx->launches += 3;
  int64_t n_slots = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&n_slots, s->grp_base + NG, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess && n_slots >= ((int64_t)1 << 32)) {
    cudaFree(key_out); cudaFree(val_in); cudaFree(glen); cudaFree(tmp); cudaFree(tmp2);
    return pscl_fail(ctx, PSCL_EINVAL, "SNP-major view needs %lld slots (>= 2^32): shard the SNPs", (long long)n_slots);
  }
  s->n_slots = n_slots;
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->cell_w, sizeof(int32_t) * 32 * (size_t)(n_slots ? n_slots : 1));
  if (e == cudaSuccess) e = cudaMemsetAsync(s->cell_w, 0xff, sizeof(int32_t) * 32 * (size_t)(n_slots ? n_slots : 1), ctx->stream);  // -1 = padding
  if (e == cudaSuccess) {
    k_fmx_csc_fill<<<FMX_GRID(P, 256), 256, 0, ctx->stream>>>(key_out, s->snp_pair, plp->cell_ptr, s->C, P, s->snp_ptr, s->grp_base, s->cell_w, s->csc_pos);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(key_out); cudaFree(val_in); cudaFree(glen); cudaFree(tmp); cudaFree(tmp2);
  if (e != cudaSuccess) return pscl_fail(ctx, PSCL_ECUDA, "SNP-major view build failed: %s", cudaGetErrorString(e));
  return PSCL_OK;
}

extern "C" int pscl_fmx_init(pscl_ctx* ctx, const pscl_plp* plp, const pscl_fmx_opts* o) {
  if (!ctx) return PSCL_EINVAL;
  PsclScope scope__(ctx);
  if (!plp || !o) return pscl_fail(ctx, PSCL_EINVAL, "pscl_fmx_init: NULL argument");
  if (o->n_clusters < 2 || o->n_clusters > PSCL_FMX_MAX_CLUSTERS)
    return pscl_fail(ctx, PSCL_EINVAL, "n_clusters must be in [2,%d] (the reference divides by nSamples-1, cmd_cram_freemux2.cpp:380)", PSCL_FMX_MAX_CLUSTERS);
  if (!plp->snp_af) return pscl_fail(ctx, PSCL_EINVAL, "freemuxlet needs the AF column of .var.gz (pscl_pileup.snp_af)");
  if (!(o->doublet_prior > 0.0 && o->doublet_prior < 1.0)) return pscl_fail(ctx, PSCL_EINVAL, "doublet_prior must be in (0,1)");
  if (!(o->geno_error >= 0.0 && o->geno_error <= 1.0)) return pscl_fail(ctx, PSCL_EINVAL, "geno_error must be in [0,1]");
  if (o->max_iter < 0) return pscl_fail(ctx, PSCL_EINVAL, "max_iter must be >= 0");
  PSCL_CUDA(ctx, cudaSetDevice(ctx->device));
  fmx_state_free(ctx);
  pscl_fmx_state* s = new pscl_fmx_state();
  ctx->fmx = s;
  s->plp = plp; s->o = *o;
  s->nS = o->n_clusters; s->npairs = s->nS * (s->nS + 1) / 2;
  s->US = (s->nS * 3 + 1) & ~1;
  s->SD = (((s->US / 2) & 1) == 0) ? s->US + 2 : s->US;
  s->C = plp->C; s->V = plp->V; s->P = plp->P;
  const size_t P1 = (size_t)(s->P ? s->P : 1), VS = (size_t)(s->V ? s->V : 1) * s->nS, C1 = (size_t)(s->C ? s->C : 1);
  const size_t NI = (size_t)(plp->n_items ? plp->n_items : 1);
  int rc = fmx_build_csc(ctx, s);
  if (rc != PSCL_OK) return rc;
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->gl_soa, sizeof(double) * 9 * P1));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->gl_csc, sizeof(double) * 9 * 32 * (size_t)(s->n_slots ? s->n_slots : 1)));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->clust_diag, sizeof(double) * 3 * VS));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->u_tab, sizeof(double) * (size_t)(s->V ? s->V : 1) * s->US));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->clust_gl, sizeof(double) * 9 * VS));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->clust_cnt, sizeof(int32_t) * 3 * VS));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->present, VS));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->cells, sizeof(pscl_fmx_cell) * C1));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->member, sizeof(int32_t) * C1));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->item_s1, sizeof(double) * 2 * NI));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->item_nrd, sizeof(int32_t) * NI));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->item_llk, sizeof(double) * NI * s->npairs));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->counters, sizeof(int32_t) * 4));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->order, sizeof(int32_t) * C1));
  PSCL_CUDA(ctx, cudaMalloc((void**)&s->work_counter, 64));
  PSCL_CUDA(ctx, cudaMemsetAsync(s->u_tab, 0, sizeof(double) * (size_t)(s->V ? s->V : 1) * s->US, ctx->stream));
  return PSCL_OK;
}

#define FMX_STATE(ctx, s)                                                        \
  if (!ctx) return PSCL_EINVAL;                                                  \
  PsclScope scope__(ctx);                                                        \
  pscl_fmx_state* s = ctx->fmx;                                                  \
  if (!s) return pscl_fail(ctx, PSCL_ESTATE, "call pscl_fmx_init first");        \
  PSCL_CUDA(ctx, cudaSetDevice(ctx->device))

extern "C" int pscl_fmx_stage1(pscl_ctx* ctx, double* stage1_dev) {
  FMX_STATE(ctx, s);
  if (!stage1_dev) return pscl_fail(ctx, PSCL_EINVAL, "pscl_fmx_stage1: NULL output");
  const pscl_plp* plp = s->plp;
  if (plp->n_items > 0) {
    S1Args a;
    a.pair_snp = plp->pair_snp; a.pair_rd = plp->pair_rd; a.rd_aq = plp->rd_aq; a.snp_af = plp->snp_af;
    a.phred_err = ctx->phred_err; a.csc_pos = s->csc_pos; a.item_pbeg = plp->item_pbeg; a.item_pend = plp->item_pend;
    a.gl_soa = s->gl_soa; a.gl_csc = s->gl_csc; a.item_s1 = s->item_s1; a.item_nrd = s->item_nrd;
    a.P = s->P; a.n_items = plp->n_items;
    int grid = std::min((plp->n_items + 7) / 8, ctx->sm_count * 8);
    k_fmx_stage1<<<grid, 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    PSCL_CUDA(ctx, cudaGetLastError());
  }
  if (s->C > 0) {
    k_fmx_stage1_cells<<<FMX_GRID(s->C, 128), 128, 0, ctx->stream>>>(plp->cell_item_ptr, plp->cell_ptr, s->item_s1, s->item_nrd, s->C, stage1_dev);
    ctx->launches++;
    PSCL_CUDA(ctx, cudaGetLastError());
  }
  s->stage1_done = true;
  return PSCL_OK;
}

// sc_drop_comp_t (sc_drop_seq.h:190-198): score descending, ties by larger id first
static void fmx_sort_order(const std::vector<double>& score, std::vector<int32_t>& order) {
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int32_t)i;
  std::sort(order.begin(), order.end(), [&](int32_t l, int32_t r) {
    double cmp = score[l] - score[r];
    if (cmp != 0) return cmp > 0;
    return l > r;
  });
}

// freemuxlet-old's seeding (cmd_cram_freemuxlet.cpp:165-346): the Bayes-factor trits on the device, the votes on the host
// in the reference's own order and on its own random numbers (glibc rand() from its default seed; std::random_shuffle of
// libstdc++ draws rand() % (i + 1) for i = 1..n-1) — the vote loops are a Gauss-Seidel sweep over the droplets and cost
// C^2 two-bit look-ups each, which is minutes on the host only beyond ~50k droplets (the reference's own matrix of
// doubles would need 40 GB there).
static int fmx_seed_old(pscl_ctx* ctx, pscl_fmx_state* s, const double* score_dev, const int32_t* init_clust_dev, int32_t* clust_dev) {
  const pscl_plp* plp = s->plp;
  const int32_t C = s->C, nS = s->nS, W = (C + 15) / 16;
  const int64_t P = s->P;
  std::vector<double> h_score(C);
  std::vector<int32_t> order(C), clusts(C, -1);
  PSCL_CUDA(ctx, cudaMemcpyAsync(h_score.data(), score_dev, sizeof(double) * C, cudaMemcpyDeviceToHost, ctx->stream));
  if (init_clust_dev) PSCL_CUDA(ctx, cudaMemcpyAsync(clusts.data(), init_clust_dev, sizeof(int32_t) * C, cudaMemcpyDeviceToHost, ctx->stream));
  // ---- pairwise trits ----
  int32_t* d_pair_cell = nullptr; uint32_t* d_trit = nullptr; double* d_scratch = nullptr; int* d_counter = nullptr; int32_t* d_cl = nullptr;
  const int warps_per_cta = 16, ctas = ctx->sm_count * 2, n_warps = warps_per_cta * ctas;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** d, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(d, bytes ? bytes : 16); };
  alloc((void**)&d_pair_cell, sizeof(int32_t) * (size_t)(P ? P : 1));
  alloc((void**)&d_trit, sizeof(uint32_t) * (size_t)C * W);
  alloc((void**)&d_scratch, sizeof(double) * 2 * (size_t)C * n_warps);
  alloc((void**)&d_counter, sizeof(int));
  alloc((void**)&d_cl, sizeof(int32_t) * (size_t)C);
  std::vector<uint32_t> trit((size_t)C * W);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_counter, 0, sizeof(int), ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_trit, 0, sizeof(uint32_t) * (size_t)C * W, ctx->stream);
  if (e == cudaSuccess) {
    if (P > 0) k_fmx_pair_cell<<<FMX_GRID(P, 256), 256, 0, ctx->stream>>>(plp->cell_ptr, C, P, d_pair_cell);
    PairwiseArgs a;
    a.cell_ptr = plp->cell_ptr; a.pair_snp = plp->pair_snp; a.gl_soa = s->gl_soa; a.snp_af = plp->snp_af; a.snp_ptr = s->snp_ptr;
    a.snp_pair = s->snp_pair; a.pair_cell = d_pair_cell; a.scratch = d_scratch; a.trit = d_trit; a.counter = d_counter;
    a.P = P; a.C = C; a.W = W; a.thres = s->o.bf_thres;
    k_fmx_pairwise<<<ctas, warps_per_cta * 32, 0, ctx->stream>>>(a);
    k_fmx_trit_symmetrize<<<FMX_GRID((int64_t)C * W, 256), 256, 0, ctx->stream>>>(d_trit, C, W);
    ctx->launches += 3;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(trit.data(), d_trit, sizeof(uint32_t) * (size_t)C * W, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) {
    auto tr = [&](int32_t i, int32_t j) { return (trit[(size_t)i * W + (j >> 4)] >> (2 * (j & 15))) & 3u; };
    fmx_sort_order(h_score, order);  // :166-171
    srand(1);  // the reference never seeds: every run of it draws glibc's default stream
    std::vector<double> votes(nS);
    if (!init_clust_dev) {  // :245-296
      for (int32_t i = 0; i < C; ++i) {
        const int32_t si = order[i];
        if ((double)i > (double)C * s->o.frac_init_clust) continue;  // :248
        for (int j = 0; j < nS; ++j) votes[j] = rand() / (RAND_MAX + 1.) / 1000.;
        for (int32_t j = 0; j < i; ++j) {
          const int32_t sj = order[j];
          const uint32_t t = tr(si, sj);
          if (t == 2u) votes[clusts[sj]] -= 1.0;
          else if (t == 1u) votes[clusts[sj]] += 1.0;
        }
        int elected = 0;
        double maxvote = votes[0];
        for (int j = 1; j < nS; ++j)
          if (maxvote < votes[j]) { elected = j; maxvote = votes[j]; }
        clusts[si] = elected;
      }
    }
    if (s->o.iter_init > 0) {  // :300-346, always ten sweeps
      std::vector<int32_t> orand(C);
      for (int sweep = 0; sweep < 10; ++sweep) {
        for (int32_t i = 0; i < C; ++i) orand[i] = i;
        for (int32_t i = 1; i < C; ++i) {  // std::random_shuffle (libstdc++)
          const int32_t j = rand() % (i + 1);
          if (i != j) std::swap(orand[i], orand[j]);
        }
        for (int32_t i = 0; i < C; ++i) {
          const int32_t si = orand[i];
          for (int j = 0; j < nS; ++j) votes[j] = rand() / (RAND_MAX + 1.) / 1000.;
          const uint32_t* row = trit.data() + (size_t)si * W;
          for (int32_t j = 0; j < C; ++j) {
            const uint32_t t = (row[j >> 4] >> (2 * (j & 15))) & 3u;  // the diagonal entry is 0
            if (t && clusts[j] >= 0) { if (t == 1u) ++votes[clusts[j]]; else --votes[clusts[j]]; }
          }
          int elected = 0;
          double maxvote = votes[0];
          for (int j = 1; j < nS; ++j)
            if (maxvote < votes[j]) { elected = j; maxvote = votes[j]; }
          if (clusts[si] >= 0 || !s->o.keep_init_missing) clusts[si] = elected;  // :336-340
        }
      }
    }
    e = cudaMemcpyAsync(d_cl, clusts.data(), sizeof(int32_t) * C, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
      k_fmx_set_clusters<<<FMX_GRID(C, 128), 128, 0, ctx->stream>>>(d_cl, C, s->cells, clust_dev);
      ctx->launches++;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(d_pair_cell); cudaFree(d_trit); cudaFree(d_scratch); cudaFree(d_counter); cudaFree(d_cl);
  if (e != cudaSuccess)
    return pscl_fail(ctx, e == cudaErrorMemoryAllocation ? PSCL_ENOMEM : PSCL_ECUDA, "freemuxlet-old seeding failed: %s", cudaGetErrorString(e));
  return PSCL_OK;
}

// the four launches of one batch; dynamic shared memory = the quotient columns [NSM][NT] + one byte per cell of the batch
template <int NSM, int NT>
static void seed3_batch_t(const Seed3Args& a, int k, cudaStream_t st) {
  const size_t sm_delta = sizeof(double) * NSM * NT + (size_t)a.bmax, sm_commit = ((size_t)a.bmax + 7) & ~(size_t)7;
  static thread_local int attr_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_dev != dev) {
    cudaFuncSetAttribute(k_fmx_seed3_delta<NSM, 1, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * NSM * NT + 2048));
    cudaFuncSetAttribute(k_fmx_seed3_delta<NSM, 2, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * NSM * NT + 2048));
    attr_dev = dev;
  }
  const dim3 grid((unsigned)a.bmax, (unsigned)a.Q);
  k_fmx_seed3_eval0<NSM, NT><<<grid, NT, 0, st>>>(a, k);
  k_fmx_seed3_delta<NSM, 1, NT><<<grid, NT, sm_delta, st>>>(a, k);
  k_fmx_seed3_delta<NSM, 2, NT><<<grid, NT, sm_delta, st>>>(a, k);
  k_fmx_seed3_commit<NT><<<grid, NT, sm_commit, st>>>(a, k);
}
static void seed3_batch(const Seed3Args& a, int k, cudaStream_t st) {
  if (a.nS <= 8) seed3_batch_t<8, 256>(a, k, st);
  else if (a.nS <= 16) seed3_batch_t<16, 256>(a, k, st);
  else seed3_batch_t<PSCL_FMX_MAX_CLUSTERS, 256>(a, k, st);
}

// greedy seeding in speculative batches (k_fmx_seed3_*): the rank-ordered SNP lists, then E0 / E1 / E2 / commit per batch;
// the batch sequence lives on the device (ctrl[]), the host reads it back every few batches to see whether it is done
static cudaError_t fmx_seed_speculative(pscl_ctx* ctx, pscl_fmx_state* s, const std::vector<int32_t>& elig, int32_t* clust_dev) {
  const pscl_plp* plp = s->plp;
  const int64_t P = s->P;
  const int n_elig = (int)elig.size();
  int bmax = 128;
  if (const char* bv = getenv("PSCL_SEED_BATCH")) bmax = std::max(1, std::min(2048, atoi(bv)));
  const int bmin = std::min(8, bmax);
  const int NCH = 8;  // batches launched between two looks at the cursor
  const auto t_begin = std::chrono::steady_clock::now();
  int rank_bits = 1;
  while ((1ll << rank_bits) <= (long long)n_elig) ++rank_bits;
  int snp_bits = 1;
  while ((1ll << snp_bits) < (long long)s->V) ++snp_bits;
  std::vector<uint32_t> h_rank((size_t)s->C, (uint32_t)n_elig);
  for (int r = 0; r < n_elig; ++r) h_rank[elig[r]] = (uint32_t)r;
  cudaError_t e = cudaSuccess;
  int32_t *d_elig = nullptr, *d_pair_cell = nullptr, *d_dec = nullptr, *d_stats = nullptr;
  double *d_diag = nullptr, *d_sc0 = nullptr, *d_part = nullptr;
  int32_t* d_arrived = nullptr;
  // a cell's pairs are split over Q CTAs of 256 threads: about one pair per thread, so a kernel lasts one chain of dependent loads
  int Q = (int)std::max<int64_t>(1, std::min<int64_t>(16, (P / std::max(s->C, 1) + 255) / 256));
  if (const char* qv = getenv("PSCL_SEED_SPLIT")) Q = std::max(1, std::min(64, atoi(qv)));
  uint32_t *d_rank_of = nullptr, *d_val = nullptr, *d_val2 = nullptr, *d_pos = nullptr;
  uint32_t* d_rk = nullptr;
  double* d_gl_rk = nullptr;
  unsigned long long *d_key = nullptr, *d_key2 = nullptr;
  Seed3Ctrl* d_ctrl = nullptr;
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0;
  const size_t n_ctrl = (size_t)n_elig + NCH + 2;
  auto alloc = [&](void** d, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(d, bytes ? bytes : 16); };
  alloc((void**)&d_elig, sizeof(int32_t) * (size_t)n_elig);
  alloc((void**)&d_rank_of, sizeof(uint32_t) * (size_t)s->C);
  alloc((void**)&d_pair_cell, sizeof(int32_t) * (size_t)P);
  alloc((void**)&d_key, sizeof(unsigned long long) * (size_t)P);
  alloc((void**)&d_key2, sizeof(unsigned long long) * (size_t)P);
  alloc((void**)&d_val, sizeof(uint32_t) * (size_t)P);
  alloc((void**)&d_val2, sizeof(uint32_t) * (size_t)P);
  alloc((void**)&d_rk, sizeof(uint32_t) * (size_t)P);
  alloc((void**)&d_gl_rk, sizeof(double) * 9 * (size_t)P);
  alloc((void**)&d_pos, sizeof(uint32_t) * (size_t)P);
  alloc((void**)&d_dec, sizeof(int32_t) * 3 * (size_t)bmax);
  alloc((void**)&d_diag, sizeof(double) * 3 * (size_t)s->V * s->nS);
  alloc((void**)&d_sc0, sizeof(double) * (size_t)bmax * PSCL_FMX_MAX_CLUSTERS);
  alloc((void**)&d_part, sizeof(double) * (size_t)bmax * Q * PSCL_FMX_MAX_CLUSTERS);
  alloc((void**)&d_arrived, sizeof(int32_t) * (size_t)bmax);
  alloc((void**)&d_stats, sizeof(int32_t) * 4);
  alloc((void**)&d_ctrl, sizeof(Seed3Ctrl) * n_ctrl);
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_key, d_key2, d_val, d_val2, (int64_t)P, 0, rank_bits + snp_bits, ctx->stream);
  alloc(&d_tmp, tmp_bytes);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_elig, elig.data(), sizeof(int32_t) * (size_t)n_elig, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_rank_of, h_rank.data(), sizeof(uint32_t) * (size_t)s->C, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_stats, 0, sizeof(int32_t) * 4, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_arrived, 0, sizeof(int32_t) * (size_t)bmax, ctx->stream);
  const Seed3Ctrl first = {0, std::min(bmin, n_elig)};
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_ctrl, &first, sizeof first, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && P > 0) {
    k_fmx_fill_f64<<<FMX_GRID((int64_t)3 * s->V * s->nS, 256), 256, 0, ctx->stream>>>(d_diag, (size_t)3 * s->V * s->nS, -1.0);
    k_fmx_pair_cell<<<FMX_GRID(P, 256), 256, 0, ctx->stream>>>(plp->cell_ptr, s->C, P, d_pair_cell);
    k_fmx_seed3_keys<<<FMX_GRID(P, 256), 256, 0, ctx->stream>>>(plp->pair_snp, d_pair_cell, d_rank_of, P, rank_bits, d_key, d_val);
    e = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_key, d_key2, d_val, d_val2, (int64_t)P, 0, rank_bits + snp_bits, ctx->stream);
    k_fmx_seed3_unpack<<<FMX_GRID(P, 256), 256, 0, ctx->stream>>>(d_key2, d_val2, s->csc_pos, s->gl_csc, P, rank_bits, d_rk, d_pos, d_gl_rk);
    ctx->launches += 5;
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  Seed3Args a;
  a.elig = d_elig; a.cell_ptr = plp->cell_ptr; a.pair_snp = plp->pair_snp; a.gl_soa = s->gl_soa; a.gl_rk = d_gl_rk; a.snp_af = plp->snp_af; a.snp_ptr = s->snp_ptr;
  a.rk = d_rk; a.pos = d_pos; a.clust_gl = s->clust_gl; a.present = s->present;
  a.diag = d_diag; a.sc0 = d_sc0; a.part = d_part; a.arrived = d_arrived; a.Q = Q; a.dA = d_dec; a.dB = d_dec + bmax; a.dC = d_dec + 2 * (size_t)bmax; a.ctrl = d_ctrl; a.stats = d_stats; a.clust = clust_dev; a.cells = s->cells;
  a.P = P; a.nS = s->nS; a.n_elig = n_elig; a.bmin = bmin; a.bmax = bmax;
  const bool trace = getenv("PSCL_TRACE") != nullptr;
  if (trace && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  const auto t_lists = std::chrono::steady_clock::now();
  Seed3Ctrl cur = first;
  for (int k = 0; e == cudaSuccess && cur.r0 < n_elig;) {
    if ((size_t)k + NCH + 1 > n_ctrl) { e = cudaErrorUnknown; break; }  // cannot happen: every batch proves at least one cell
    for (int i = 0; i < NCH; ++i, ++k) {
      seed3_batch(a, k, ctx->stream);
      ctx->launches += 4;
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&cur, d_ctrl + k, sizeof cur, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  if (e == cudaSuccess && trace) {
    int32_t st[4] = {0, 0, 0, 0};
    cudaMemcpy(st, d_stats, sizeof st, cudaMemcpyDeviceToHost);
    const auto t_end = std::chrono::steady_clock::now();
    fprintf(stderr, "[pscl_fmx_seed] %d cells in %d batches of <= %d (%d needed a third round, which proved %d more cells; smallest batch %d); "
            "%d CTAs per cell; rank-ordered SNP lists %.2f ms, batches %.2f ms\n", n_elig, st[0], bmax, st[1], st[2], st[3], Q,
            std::chrono::duration<double, std::milli>(t_lists - t_begin).count(), std::chrono::duration<double, std::milli>(t_end - t_lists).count());
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_elig); cudaFree(d_rank_of); cudaFree(d_pair_cell); cudaFree(d_key); cudaFree(d_key2); cudaFree(d_val); cudaFree(d_val2);
  cudaFree(d_rk); cudaFree(d_gl_rk); cudaFree(d_pos); cudaFree(d_dec); cudaFree(d_stats); cudaFree(d_ctrl); cudaFree(d_tmp); cudaFree(d_diag); cudaFree(d_sc0); cudaFree(d_part); cudaFree(d_arrived);
  return e;
}

extern "C" int pscl_fmx_seed(pscl_ctx* ctx, const double* stage1_dev, const int32_t* init_clust_dev, int32_t* clust_dev) {
  FMX_STATE(ctx, s);
  if (!s->stage1_done) return pscl_fail(ctx, PSCL_ESTATE, "pscl_fmx_seed before pscl_fmx_stage1");
  if (!stage1_dev || !clust_dev) return pscl_fail(ctx, PSCL_EINVAL, "pscl_fmx_seed: NULL argument");
  const pscl_plp* plp = s->plp;
  double* score = nullptr;
  PSCL_CUDA(ctx, cudaMalloc((void**)&score, sizeof(double) * (size_t)(s->C ? s->C : 1)));
  if (s->C > 0) {
    k_fmx_begin<<<FMX_GRID(s->C, 128), 128, 0, ctx->stream>>>(stage1_dev, init_clust_dev, s->C, s->cells, clust_dev, score);
    ctx->launches++;
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && s->o.mode_old && s->C > 0 && (!init_clust_dev || s->o.iter_init > 0)) {
    const int rc_old = fmx_seed_old(ctx, s, score, init_clust_dev, clust_dev);
    cudaFree(score);
    cudaFree(s->csc_pos); s->csc_pos = nullptr;
    if (rc_old != PSCL_OK) return rc_old;
    s->begun = true;
    s->iters = 0;
    return PSCL_OK;
  }
  if (e == cudaSuccess && !init_clust_dev && s->C > 0) {
    std::vector<double> h_score(s->C);
    std::vector<int32_t> h_order(s->C);
    e = cudaMemcpyAsync(h_score.data(), score, sizeof(double) * s->C, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && s->o.randomize_singlet_score) {
      // --randomize-singlet-score (cmd_cram_freemux2.cpp:164-181): Fisher-Yates on the scores with the libc rand()
      // stream the reference uses; the kernel's threshold test (:226) must see the shuffled scores too
      srand(s->o.seed == 0 ? (unsigned)time(nullptr) : (unsigned)s->o.seed);
      for (int32_t i = 0; i < s->C - 1; ++i) {
        const int32_t j = i + rand() % (s->C - i);
        if (i < j) std::swap(h_score[i], h_score[j]);
      }
      e = cudaMemcpyAsync(score, h_score.data(), sizeof(double) * s->C, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e == cudaSuccess) {
      fmx_sort_order(h_score, h_order);
      e = cudaMemcpyAsync(s->order, h_order.data(), sizeof(int32_t) * s->C, cudaMemcpyHostToDevice, ctx->stream);
    }
    const size_t VS = (size_t)s->V * s->nS;
    if (e == cudaSuccess && VS > 0) {
      k_fmx_fill_f64<<<FMX_GRID(VS * 9, 256), 256, 0, ctx->stream>>>(s->clust_gl, VS * 9, 1.0);
      ctx->launches++;
      e = cudaMemsetAsync(s->present, 0, VS, ctx->stream);
    }
    // the one-CTA chain (kept as the cross-check of the batched forms; also what a state seeded a second time takes beyond 24
    // clusters, where the older batched form's 24-bit cluster mask ends)
    const bool serial = getenv("PSCL_SEED_SERIAL") != nullptr || (s->nS > 24 && (getenv("PSCL_SEED_V2") || !s->csc_pos));
    if (e == cudaSuccess && serial) {
      SeedArgs a;
      a.order = s->order; a.score = score; a.cell_ptr = plp->cell_ptr; a.pair_snp = plp->pair_snp; a.gl_soa = s->gl_soa;
      a.snp_af = plp->snp_af; a.clust_gl = s->clust_gl; a.present = s->present; a.clust = clust_dev; a.cells = s->cells;
      a.P = s->P; a.C = s->C; a.nS = s->nS; a.frac = s->o.frac_init_clust; a.thres = s->o.singlet_score_thres;
      k_fmx_seed<<<1, 1024, 0, ctx->stream>>>(a);
      ctx->launches++;
      e = cudaGetLastError();
    } else if (e == cudaSuccess) {
      // cells that take part (:225-226), in seeding order, and the running pair count of their pileups
      std::vector<int32_t> elig;
      std::vector<int64_t> epair(1, 0);
      elig.reserve(s->C);
      for (int32_t i = 0; i < s->C; ++i) {
        const int32_t si = h_order[i];
        if ((double)i > (double)s->C * s->o.frac_init_clust) continue;
        if (h_score[si] < s->o.singlet_score_thres) continue;
        elig.push_back(si);
        epair.push_back(epair.back() + (plp->h_cell_ptr[si + 1] - plp->h_cell_ptr[si]));
      }
      const int n_elig = (int)elig.size();
      // the batched form above (8 cells per batch, serial corrections) is kept as a cross-check, and takes over when a state is
      // seeded a second time (csc_pos is gone)
      if (getenv("PSCL_SEED_V2") || !s->csc_pos || false) {
        int B = 8;
        if (const char* bv = getenv("PSCL_SEED_BATCH")) B = std::max(1, std::min(256, atoi(bv)));
        int64_t max_pairs = 1;
        for (int b0 = 0; b0 < n_elig; b0 += B) max_pairs = std::max(max_pairs, epair[std::min(n_elig, b0 + B)] - epair[b0]);
        int32_t* d_elig = nullptr; int64_t* d_epair = nullptr; unsigned long long* d_mark = nullptr; double *d_contrib = nullptr, *d_d0p = nullptr;
        auto alloc = [&](void** d, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(d, bytes ? bytes : 16); };
        alloc((void**)&d_elig, sizeof(int32_t) * (size_t)std::max(n_elig, 1));
        alloc((void**)&d_epair, sizeof(int64_t) * ((size_t)n_elig + 1));
        alloc((void**)&d_mark, sizeof(unsigned long long) * (size_t)std::max(s->V, 1));
        alloc((void**)&d_contrib, sizeof(double) * (size_t)max_pairs * s->nS);
        alloc((void**)&d_d0p, sizeof(double) * (size_t)B * PSCL_SEED_SPLIT * s->nS);
        if (e == cudaSuccess && n_elig) e = cudaMemcpyAsync(d_elig, elig.data(), sizeof(int32_t) * n_elig, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_epair, epair.data(), sizeof(int64_t) * ((size_t)n_elig + 1), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_mark, 0, sizeof(unsigned long long) * (size_t)std::max(s->V, 1), ctx->stream);
        SeedBatchArgs a;
        a.elig = d_elig; a.epair = d_epair; a.cell_ptr = plp->cell_ptr; a.pair_snp = plp->pair_snp; a.gl_soa = s->gl_soa; a.snp_af = plp->snp_af;
        a.clust_gl = s->clust_gl; a.present = s->present; a.mark = d_mark; a.contrib = d_contrib; a.d0p = d_d0p; a.clust = clust_dev;
        a.cells = s->cells; a.P = s->P; a.nS = s->nS;
        for (int b0 = 0, batch = 0; b0 < n_elig && e == cudaSuccess; b0 += B, ++batch) {
          a.base = b0; a.nb = std::min(B, n_elig - b0); a.batch = batch;
          const dim3 grid((unsigned)a.nb, PSCL_SEED_SPLIT);
          if (s->nS <= 8) { k_fmx_seed_dist<8><<<grid, 256, 0, ctx->stream>>>(a); k_fmx_seed_commit<8><<<1, 1024, 0, ctx->stream>>>(a); }
          else if (s->nS <= 16) { k_fmx_seed_dist<16><<<grid, 256, 0, ctx->stream>>>(a); k_fmx_seed_commit<16><<<1, 1024, 0, ctx->stream>>>(a); }
          else { k_fmx_seed_dist<PSCL_FMX_MAX_CLUSTERS><<<grid, 256, 0, ctx->stream>>>(a); k_fmx_seed_commit<PSCL_FMX_MAX_CLUSTERS><<<1, 1024, 0, ctx->stream>>>(a); }
          k_fmx_seed_merge<<<grid, 256, 0, ctx->stream>>>(a);
          ctx->launches += 3;
          e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // elig / epair are pageable sources
        cudaFree(d_elig); cudaFree(d_epair); cudaFree(d_mark); cudaFree(d_contrib); cudaFree(d_d0p);
      } else if (n_elig > 0) {
        e = fmx_seed_speculative(ctx, s, elig, clust_dev);
      }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // h_order is a pageable source
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(score);
  cudaFree(s->csc_pos); s->csc_pos = nullptr;  // stage 1 scatters through it, the speculative seeding finds the GL records with it
  if (e != cudaSuccess) return pscl_fail(ctx, PSCL_ECUDA, "freemuxlet seeding failed: %s", cudaGetErrorString(e));
  s->begun = true;
  s->iters = 0;
  return PSCL_OK;
}

static int fmx_mstep_launch(pscl_ctx* ctx, pscl_fmx_state* s, const int32_t* member, int final_) {
  const int64_t n = (int64_t)s->V * s->nS;
  if (n == 0) return PSCL_OK;
  MArgs a;
  a.snp_ptr = s->snp_ptr; a.grp_base = s->grp_base; a.cell_w = s->cell_w; a.snp_pair = s->snp_pair; a.gl_csc = s->gl_csc; a.member = member;
  a.pair_rd = s->plp->pair_rd; a.rd_aq = s->plp->rd_aq; a.clust_diag = s->clust_diag; a.clust_gl = s->clust_gl;
  a.clust_cnt = s->clust_cnt; a.V = s->V; a.nS = s->nS; a.final_ = final_;
  {
    const int threads = s->nS <= 16 ? 128 : 64;
    const size_t smem = (size_t)s->nS * threads * (9 * sizeof(double) + 3 * sizeof(int));
    static bool attr_set[64] = {false};
    if (!attr_set[ctx->device & 63]) {
      PSCL_CUDA(ctx, cudaFuncSetAttribute(k_fmx_mstep<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      PSCL_CUDA(ctx, cudaFuncSetAttribute(k_fmx_mstep<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set[ctx->device & 63] = true;
    }
    const int64_t vpad = ((int64_t)s->V + 31) / 32 * 32;  // whole warps: a group of 32 SNPs is never split
    if (threads == 128) k_fmx_mstep<128><<<FMX_GRID(vpad, 128), 128, smem, ctx->stream>>>(a);
    else k_fmx_mstep<64><<<FMX_GRID(vpad, 64), 64, smem, ctx->stream>>>(a);
  }
  ctx->launches++;
  PSCL_CUDA(ctx, cudaGetLastError());
  s->final_tab = final_ != 0;
  return PSCL_OK;
}

extern "C" int pscl_fmx_mstep(pscl_ctx* ctx, const int32_t* clust_dev) {
  FMX_STATE(ctx, s);
  if (!s->begun) return pscl_fail(ctx, PSCL_ESTATE, "pscl_fmx_mstep before pscl_fmx_seed");
  if (clust_dev) PSCL_CUDA(ctx, cudaMemcpyAsync(s->member, clust_dev, sizeof(int32_t) * (size_t)s->C, cudaMemcpyDeviceToDevice, ctx->stream));
  PSCL_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = fmx_mstep_launch(ctx, s, s->member, 0);
  PSCL_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  return rc;
}

template <int NS, int J0, int J1, int THREADS>
static int fmx_estep_launch(pscl_ctx* ctx, pscl_fmx_state* s, const EArgs& a) {
  constexpr int NA = J1 * (J1 + 1) / 2 - J0 * (J0 + 1) / 2;
  constexpr int RD = (J1 * 3 + 1) & ~1;
  const int SD = NS > 0 ? ((((RD / 2) & 1) == 0) ? RD + 2 : RD) : s->SD;  // as in the kernel
  if (2 * SD < NA) return pscl_fail(ctx, PSCL_EINVAL, "internal: E-step staging rows too short (SD=%d, NA=%d)", SD, NA);
  const size_t smem = sizeof(double) * 2 * THREADS * (size_t)SD + sizeof(int) * (size_t)NA * THREADS;
  auto kern = k_fmx_estep<NS, J0, J1, THREADS>;
  PSCL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  PSCL_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
  if (per_sm < 1) return pscl_fail(ctx, PSCL_EINVAL, "E-step kernel does not fit one SM (smem %zu B)", smem);
  int grid = ctx->sm_count * per_sm;
  const int need = (a.n_items + THREADS / 32 - 1) / (THREADS / 32);
  if (grid > need) grid = need;
  PSCL_CUDA(ctx, cudaMemsetAsync(s->work_counter, 0, sizeof(int), ctx->stream));
  kern<<<grid, THREADS, smem, ctx->stream>>>(a);
  ctx->launches++;
  PSCL_CUDA(ctx, cudaGetLastError());
  return PSCL_OK;
}

static int fmx_estep16_team_launch(pscl_ctx* ctx, pscl_fmx_state* s, const EArgs& a) {
  const size_t smem = sizeof(double) * 3 * 32 * 50 + sizeof(int) * 40 * 128 + sizeof(int) * 2 * 1024;
  auto kern = k_fmx_estep16_team;
  PSCL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  PSCL_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
  if (per_sm < 1) return pscl_fail(ctx, PSCL_EINVAL, "E-step team kernel does not fit one SM (smem %zu B)", smem);
  int grid = ctx->sm_count * per_sm;
  if (grid > a.n_items) grid = a.n_items;
  PSCL_CUDA(ctx, cudaMemsetAsync(s->work_counter, 0, sizeof(int), ctx->stream));
  kern<<<grid, 128, smem, ctx->stream>>>(a);
  ctx->launches++;
  PSCL_CUDA(ctx, cudaGetLastError());
  return PSCL_OK;
}

extern "C" int pscl_fmx_estep(pscl_ctx* ctx, int32_t iter, double* llk_dev) {
  FMX_STATE(ctx, s);
  if (!s->begun) return pscl_fail(ctx, PSCL_ESTATE, "pscl_fmx_estep before pscl_fmx_seed / pscl_fmx_mstep");
  if (!llk_dev) return pscl_fail(ctx, PSCL_EINVAL, "pscl_fmx_estep: NULL output");
  const pscl_plp* plp = s->plp;
  // geno_error is mixed in every iteration by freemux2 (:410), only in the last one by freemuxlet-old (:485)
  const int apply_err = s->o.mode_old ? (s->o.geno_error > 0 && iter + 1 == s->o.max_iter) : (s->o.geno_error > 0);
  PSCL_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  const int64_t n = (int64_t)s->V * s->nS;
  if (n > 0) {
    k_fmx_posterior<<<FMX_GRID(n, 256), 256, 0, ctx->stream>>>(s->clust_diag, plp->snp_af, s->V, s->nS, s->US, s->o.geno_error, apply_err, s->u_tab);
    ctx->launches++;
    PSCL_CUDA(ctx, cudaGetLastError());
  }
  if (plp->n_items > 0) {
    EArgs a;
    a.pair_snp = plp->pair_snp; a.gl_soa = s->gl_soa; a.u_tab = s->u_tab; a.item_order = plp->item_order;
    a.item_pbeg = plp->item_pbeg; a.item_pend = plp->item_pend; a.item_llk = s->item_llk; a.counter = s->work_counter;
    a.P = s->P; a.n_items = plp->n_items; a.npairs = s->npairs; a.nS = s->nS; a.US = s->US; a.SD = s->SD;
    int rc = PSCL_OK;
    const int nS = s->nS;
    switch (nS) {
      case 2: rc = fmx_estep_launch<2, 0, 2, 128>(ctx, s, a); break;
      case 3: rc = fmx_estep_launch<3, 0, 3, 128>(ctx, s, a); break;
      case 4: rc = fmx_estep_launch<4, 0, 4, 128>(ctx, s, a); break;
      case 5: rc = fmx_estep_launch<5, 0, 5, 128>(ctx, s, a); break;
      case 6: rc = fmx_estep_launch<6, 0, 6, 128>(ctx, s, a); break;
      case 7: rc = fmx_estep_launch<7, 0, 7, 128>(ctx, s, a); break;
      case 8: rc = fmx_estep_launch<8, 0, 8, 128>(ctx, s, a); break;
      case 16:  // configs[4]'s cluster count: one pass, a team of four warps per work item (PSCL_ESTEP_TILES=1: the four
                // compile-time tile launches it replaces, kept as its cross-check)
        if (!getenv("PSCL_ESTEP_TILES")) { rc = fmx_estep16_team_launch(ctx, s, a); break; }
        rc = fmx_estep_launch<16, 0, 8, 64>(ctx, s, a);
        if (rc == PSCL_OK) rc = fmx_estep_launch<16, 8, 11, 64>(ctx, s, a);
        if (rc == PSCL_OK) rc = fmx_estep_launch<16, 11, 14, 64>(ctx, s, a);
        if (rc == PSCL_OK) rc = fmx_estep_launch<16, 14, 16, 64>(ctx, s, a);
        break;
      default:
        rc = fmx_estep_launch<0, 0, 8, 64>(ctx, s, a);
        if (rc == PSCL_OK && nS > 8) rc = fmx_estep_launch<0, 8, 11, 64>(ctx, s, a);
        if (rc == PSCL_OK && nS > 11) rc = fmx_estep_launch<0, 11, 14, 64>(ctx, s, a);
        if (rc == PSCL_OK && nS > 14) rc = fmx_estep_launch<0, 14, 16, 64>(ctx, s, a);
        if (rc == PSCL_OK && nS > 16) rc = fmx_estep_launch<0, 16, 18, 64>(ctx, s, a);
        if (rc == PSCL_OK && nS > 18) rc = fmx_estep_launch<0, 18, 20, 64>(ctx, s, a);
        if (rc == PSCL_OK && nS > 20) rc = fmx_estep_launch<0, 20, 22, 64>(ctx, s, a);
        if (rc == PSCL_OK && nS > 22) rc = fmx_estep_launch<0, 22, 24, 64>(ctx, s, a);
        if (rc == PSCL_OK && nS > 24) rc = fmx_estep_launch<0, 24, 25, 64>(ctx, s, a);  // one row per launch from here: 25..32 accumulators
        if (rc == PSCL_OK && nS > 25) rc = fmx_estep_launch<0, 25, 26, 64>(ctx, s, a);
        if (rc == PSCL_OK && nS > 26) rc = fmx_estep_launch<0, 26, 27, 64>(ctx, s, a);
        if (rc == PSCL_OK && nS > 27) rc = fmx_estep_launch<0, 27, 28, 64>(ctx, s, a);
        if (rc == PSCL_OK && nS > 28) rc = fmx_estep_launch<0, 28, 29, 64>(ctx, s, a);
        if (rc == PSCL_OK && nS > 29) rc = fmx_estep_launch<0, 29, 30, 64>(ctx, s, a);
        if (rc == PSCL_OK && nS > 30) rc = fmx_estep_launch<0, 30, 31, 64>(ctx, s, a);
        if (rc == PSCL_OK && nS > 31) rc = fmx_estep_launch<0, 31, 32, 64>(ctx, s, a);
    }
    if (rc != PSCL_OK) return rc;
  }
  if ((int64_t)s->C * s->npairs > 0) {
    k_fmx_llk_reduce<<<FMX_GRID((int64_t)s->C * s->npairs, 256), 256, 0, ctx->stream>>>(plp->cell_item_ptr, s->item_llk, s->C, s->npairs, llk_dev);
    ctx->launches++;
    PSCL_CUDA(ctx, cudaGetLastError());
  }
  PSCL_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  return PSCL_OK;
}

extern "C" int pscl_fmx_classify(pscl_ctx* ctx, const double* llk_dev, int32_t* clust_dev, pscl_fmx_result* res) {
  FMX_STATE(ctx, s);
  if (!s->begun) return pscl_fail(ctx, PSCL_ESTATE, "pscl_fmx_classify before pscl_fmx_seed");
  if (!llk_dev || !clust_dev) return pscl_fail(ctx, PSCL_EINVAL, "pscl_fmx_classify: NULL argument");
  PSCL_CUDA(ctx, cudaMemsetAsync(s->counters, 0, sizeof(int32_t) * 4, ctx->stream));
  if (s->C > 0) {
    k_fmx_classify<<<FMX_GRID(s->C, 128), 128, 0, ctx->stream>>>(llk_dev, s->C, s->nS, s->o.doublet_prior, s->o.mode_old, s->cells, clust_dev, s->member, s->counters);
    ctx->launches++;
    PSCL_CUDA(ctx, cudaGetLastError());
  }
  int32_t h[4] = {0, 0, 0, 0};
  PSCL_CUDA(ctx, cudaMemcpyAsync(h, s->counters, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  PSCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  s->iters++;
  s->last.n_iter = s->iters; s->last.n_changed = h[0]; s->last.n_singlet = h[1]; s->last.n_ambiguous = h[2];
  s->last.n_doublet = s->C - h[1] - h[2];
  if (res) *res = s->last;
  return PSCL_OK;
}

// Per-cell records (nullable) and the cluster pileups of SNPs [v0, v1) of the current membership, written to rows
// [v0, v1) of the caller's [V][nS][9] / [V][nS][3] arrays (an SNP shard owns exactly its own rows).
static int fmx_fetch_range(pscl_ctx* ctx, pscl_fmx_cell* out, double* clust_gl, int32_t* clust_cnt, int32_t v0, int32_t v1) {
  pscl_fmx_state* s = ctx->fmx;
  if (!s || !s->begun) return pscl_fail(ctx, PSCL_ESTATE, "pscl_fmx_fetch before pscl_fmx_seed");
  if (out && s->C > 0)
    PSCL_CUDA(ctx, cudaMemcpyAsync(out, s->cells, sizeof(pscl_fmx_cell) * (size_t)s->C, cudaMemcpyDeviceToHost, ctx->stream));
  if (clust_gl || clust_cnt) {
    // full 9-GL pileups and read counts of the current membership (what clustPileup holds when the
    // reference writes .clust1.vcf.gz, cmd_cram_freemux2.cpp:608-658)
    int rc = fmx_mstep_launch(ctx, s, s->member, 1);
    if (rc != PSCL_OK) return rc;
    const size_t off = (size_t)v0 * s->nS, n = (size_t)(v1 > v0 ? v1 - v0 : 0) * s->nS;
    if (clust_gl && n) PSCL_CUDA(ctx, cudaMemcpyAsync(clust_gl + 9 * off, s->clust_gl + 9 * off, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (clust_cnt && n) PSCL_CUDA(ctx, cudaMemcpyAsync(clust_cnt + 3 * off, s->clust_cnt + 3 * off, sizeof(int32_t) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  PSCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PSCL_OK;
}

extern "C" int pscl_fmx_fetch(pscl_ctx* ctx, pscl_fmx_cell* out, double* clust_gl, int32_t* clust_cnt) {
  FMX_STATE(ctx, s);
  return fmx_fetch_range(ctx, out, clust_gl, clust_cnt, 0, s->V);
}

extern "C" int pscl_fmx_last_kernel_ms(pscl_ctx* ctx, float* ms) {
  FMX_STATE(ctx, s);
  (void)s;
  PSCL_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
  float t = 0.f;
  PSCL_CUDA(ctx, cudaEventElapsedTime(&t, ctx->ev0, ctx->ev1));
  if (ms) *ms = t;
  return PSCL_OK;
}

extern "C" int pscl_fmx_run_aux(pscl_ctx* ctx, const pscl_pileup* host, const pscl_fmx_opts* opts, const int32_t* init_clust,
                                pscl_fmx_cell* out, double* clust_gl, int32_t* clust_cnt, pscl_fmx_result* res,
                                double* clust_gl0, int32_t* clust_cnt0);
extern "C" int pscl_fmx_run(pscl_ctx* ctx, const pscl_pileup* host, const pscl_fmx_opts* opts, const int32_t* init_clust,
                            pscl_fmx_cell* out, double* clust_gl, int32_t* clust_cnt, pscl_fmx_result* res) {
  return pscl_fmx_run_aux(ctx, host, opts, init_clust, out, clust_gl, clust_cnt, res, nullptr, nullptr);
}

extern "C" int pscl_fmx_run_aux(pscl_ctx* ctx, const pscl_pileup* host, const pscl_fmx_opts* opts, const int32_t* init_clust,
                                pscl_fmx_cell* out, double* clust_gl, int32_t* clust_cnt, pscl_fmx_result* res,
                                double* clust_gl0, int32_t* clust_cnt0) {
  if (!ctx) return PSCL_EINVAL;
  PsclScope scope__(ctx);
  if (!host || !opts || !out) return pscl_fail(ctx, PSCL_EINVAL, "pscl_fmx_run: NULL argument");
  pscl_plp* plp = nullptr;
  int rc = pscl_plp_upload(ctx, host, &plp);
  if (rc != PSCL_OK) return rc;
  int32_t* d_init = nullptr;
  auto body = [&]() -> int {
    int r = pscl_fmx_init(ctx, plp, opts);
    if (r != PSCL_OK) return r;
    pscl_fmx_state* s = ctx->fmx;
    const size_t C1 = (size_t)(s->C ? s->C : 1);
    PSCL_CUDA(ctx, cudaMalloc((void**)&s->own_stage1, sizeof(double) * 4 * C1));
    PSCL_CUDA(ctx, cudaMalloc((void**)&s->own_llk, sizeof(double) * C1 * s->npairs));
    PSCL_CUDA(ctx, cudaMalloc((void**)&s->own_clust, sizeof(int32_t) * C1));
    if (init_clust) {
      for (int32_t c = 0; c < s->C; ++c)
        if (init_clust[c] >= s->nS) return pscl_fail(ctx, PSCL_EINVAL, "init_clust[%d] = %d is not below n_clusters = %d", c, init_clust[c], s->nS);
      PSCL_CUDA(ctx, cudaMalloc((void**)&d_init, sizeof(int32_t) * C1));
      PSCL_CUDA(ctx, cudaMemcpyAsync(d_init, init_clust, sizeof(int32_t) * (size_t)s->C, cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((r = pscl_fmx_stage1(ctx, s->own_stage1)) != PSCL_OK) return r;
    if ((r = pscl_fmx_seed(ctx, s->own_stage1, d_init, s->own_clust)) != PSCL_OK) return r;
    if ((r = pscl_fmx_mstep(ctx, s->own_clust)) != PSCL_OK) return r;  // :277-288
    // --aux-files: the cluster pileups of the initial assignment, as the reference writes them to .clust0.vcf.gz (:291-347)
    if ((clust_gl0 || clust_cnt0) && (r = pscl_fmx_fetch(ctx, nullptr, clust_gl0, clust_cnt0)) != PSCL_OK) return r;
    pscl_fmx_result rr;
    memset(&rr, 0, sizeof(rr));
    for (int iter = 0; iter < s->o.max_iter; ++iter) {
      if ((r = pscl_fmx_estep(ctx, iter, s->own_llk)) != PSCL_OK) return r;
      if ((r = pscl_fmx_classify(ctx, s->own_llk, s->own_clust, &rr)) != PSCL_OK) return r;
      if ((r = pscl_fmx_mstep(ctx, nullptr)) != PSCL_OK) return r;
      if (!s->o.mode_old && s->o.early_stop && rr.n_changed == 0) break;  // :601-604
    }
    if (res) *res = rr;
    return pscl_fmx_fetch(ctx, out, clust_gl, clust_cnt);
  };
  rc = body();
  std::string err = ctx->err;
  cudaFree(d_init);
  fmx_state_free(ctx);
  pscl_plp_free(ctx, plp);
  ctx->err = err;
  return rc;
}
