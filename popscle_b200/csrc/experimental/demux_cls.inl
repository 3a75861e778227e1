// experimental/demux_cls.inl (compiled only with -DPSCL_EXPERIMENTAL; measured slower than k_demux_default, kept as a record) — class-split demuxlet kernel for the default alpha grid {0, 0.5}, 2 <= nv <= 8
// (part of popscle_b200.cu; replaces cmd_cram_demuxlet.cpp:655-747).  k_demux_default (demux.inl) is
// the lane-per-pair baseline this kernel is measured against; both are parity-tested.
//
// What changes against k_demux_default (ncu: FP64-issue bound, 225 FP64 of 637 instructions per 32
// pairs, profiles/r01a_*):
//
//  1. Fewer FP64 instructions.  For a (cell,SNP) pair with at most one usable base-call — three
//     quarters of all pairs in droplet data — the per-read factor pR(1-p)+pA*p (:685) is LINEAR in
//     the mixing fraction p = (l+m)/4, so with h[i] = a + b*i
//         sum_{l,m} g_j[l] g_k[m] h[l+m] = S_k (a S_j + b M_j) + M_k (b S_j),
//         S_j = sum_l g_j[l],  M_j = sum_l l g_j[l]
//     i.e. 2 FP64 instructions per doublet instead of 3 (+3 instead of 9 per sample), on a
//     (S,M) genotype table of 16 B per sample instead of 24 B.  S_j is carried exactly (it is 1 only
//     up to the float32 rounding of the VCF values), so nothing is approximated.  Pairs with more
//     usable base-calls ("class M": 2-3, folded from a shared-memory table; "class D": > 3, folded
//     once per upload into a 5-value table) keep the Hankel form of k_demux_default.
//  2. No index chasing in the hot loop.  The pileup image is re-ordered once per upload into 8-byte
//     records {snp, up to three packed base-calls}, class S first, then M, then D inside each cell
//     (k_dmx_classify / k_dmx_scatter): one coalesced 8-byte load per lane per batch replaces the
//     pair_snp / pair_rd / rd_aq / has_gp loads, and a warp never mixes the two forms.  SNPs without
//     genotypes are marked in the genotype tables themselves (first value -1).
//
// Skeleton as k_demux_default: persistent CTAs, a warp owns a work item, ONE LANE PER PAIR, the
// running products live in that lane's registers (mantissa + exponent, one log per accumulator per
// item); records two batches ahead in registers (+ an L2 prefetch eight batches ahead), the 32
// genotype rows of the next batch gathered cooperatively with cp.async into the other half of a
// shared-memory double buffer while the current batch is multiplied.
//
// Status (config 2, one B200, profiles/r02*): 0.79 ms against 0.75 ms for k_demux_default, i.e. NOT
// faster although it executes a third fewer instructions (421 against 637 per 32 pairs).  Both kernels
// sit on the same wall: the per-pair gather of a 128-192 B genotype row out of L2.  tools/gather_bench.cu
// measures that gather alone at 0.33-0.37 ms for config 2's 19.9 M rows with 8 warps per SM, 0.45 ms with
// 128 independent FP64 FMAs per pair beside it and 0.57 ms with 256 — the floor of any lane-per-pair
// design here.  PSCL_CLS timing variants (no gather: 0.43 ms, no math: 0.75 ms, neither: 0.26 ms) show
// the gather does not overlap with the rest at 2 warps per scheduler, and 10-12 warps per SM were slower
// (0.91 ms, spills + shared-memory pressure).  A warp-specialised variant (producer warps + mbarrier
// pipeline, profiles/r02d_*) measured 0.80 ms.  k_demux_default therefore stays the default; this kernel
// is selectable (pscl_demux_select_kernel(ctx, 3)) and parity-tested, and is the base for the next round.


#ifndef CLS_THREADS
#define CLS_THREADS 256
#endif
#ifndef CLS_NBUF
#define CLS_NBUF 3       /* genotype-row buffers per warp: CLS_NBUF-1 gathers in flight behind the batch being multiplied */
#endif
#ifndef CLS_PF
#define CLS_PF 8         /* record batches in flight per warp (power of two) */
#endif
#define CLS_PKT_B 272    /* one batch packet = one bulk copy: 16-byte header + 32 records of 8 bytes */
#define CLS_SLOT_B 288
#define CLS_FLAG_M 1u    /* Hankel form (classes M and D) */
#define CLS_FLAG_END 2u  /* last batch of its work item */
#define CLS_FLAG_EXIT 4u
#define CLS_FLAG_D 8u    /* class D: the fold comes from the deep table */

struct ClsArgs {
  const unsigned char* pkt;  // [n_pkt][CLS_PKT_B] batch packets: 16-byte header {flags | n<<8, item, -, -} + 32 records,
                             // S/M record {snp, b0 | b1<<8 | b2<<16 | cnt<<24}, D record {snp, deep row}
  const double* deep;        // [n_deep][6] folded per-read factors f0..f4 (max 1) of the class-D pairs
  const double* gpM;         // [V][RM]   genotype rows, padded to 16 B; first value -1 = SNP without GP
  const double* gpS;         // [V][2 nv] (S_j, M_j);                   first value -1 = SNP without GP
  const double* fold_tab;    // [3*64][PSCL_FOLD_ROW]
  const uint4* desc;         // [n_work] {first packet, end packet, item, -} of a work item
  double* partial;           // [items][nv*nv*2], row (item - item_base)
  int* counter;
  int32_t item_base, n_work;
  int32_t n_snps;  // rows of gpM / gpS
};

__host__ __device__ constexpr int cls_pow2ceil(int x) { return x <= 1 ? 1 : x <= 2 ? 2 : x <= 4 ? 4 : x <= 8 ? 8 : x <= 16 ? 16 : 32; }

template <int NV>
struct ClsCfg {
  static constexpr int THREADS = CLS_THREADS;
  static constexpr int ND = NV * (NV - 1) / 2;
  static constexpr int NE = NV + ND + 2;        // singlets | doublets (k<j) | k=0 column factor | pair normaliser
  static constexpr int E_SG0 = NV + ND, E_MX = NV + ND + 1;
  static constexpr int RM = (3 * NV + 1) & ~1;  // doubles per class-M row
  static constexpr int RS = 2 * NV;             // doubles per class-S row
  static constexpr int ROWB_M = RM * 8, ROWB_S = RS * 8;
  static constexpr int STRIDE = ((ROWB_M / 16) | 1) * 16;  // odd multiple of 16 B: conflict-free LDS.128 per lane
  static constexpr int NCH_M = ROWB_M / 16, NCH_S = ROWB_S / 16;
  static constexpr int LPR_M = cls_pow2ceil(NCH_M), LPR_S = cls_pow2ceil(NCH_S);
  static constexpr int OFF_TABM = 0;
  static constexpr int OFF_TABS = OFF_TABM + 3 * 64 * PSCL_FOLD_ROW * 8;
  static constexpr int OFF_ROWS = OFF_TABS + 256 * 4 * 8;  // tabS is indexed by a raw record byte
  // per warp: CLS_NBUF genotype-row buffers | packet ring (CLS_PF slots of 288 B) | mbarriers
  static constexpr int OFF_RING = CLS_NBUF * 32 * STRIDE;
  static constexpr int OFF_BAR = OFF_RING + CLS_PF * CLS_SLOT_B;
  static constexpr int WARP_B = (OFF_BAR + CLS_PF * 8 + 127) & ~127;
  static constexpr size_t SMEM = (size_t)OFF_ROWS + (size_t)(THREADS / 32) * WARP_B;
  static_assert(NCH_M <= 32 && NE <= 40 && (32 * STRIDE) % 16 == 0, "row / accumulator geometry");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// L2 residency: the genotype tables (tens of MB, gathered ~200 times per row) are kept with evict_last,
// the packet stream (read once) passes with evict_first so that it does not push table lines out
__device__ __forceinline__ uint64_t cls_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t cls_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void cls_cp_async16(uint32_t dst, const void* src, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(policy) : "memory");
}
__device__ __forceinline__ void cls_prefetch_l2_keep(const void* p) { asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p)); }
__device__ __forceinline__ void cls_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cls_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// ---- record ring: TMA bulk copies completing on an mbarrier (independent of the cp.async groups of the rows)
__device__ __forceinline__ void cls_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cls_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cls_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void cls_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ double cls_pmax(double x, double y) { return x > y ? x : y; }  // positive, non-NaN operands

#define cls_transpose_prod pscl_transpose_prod  /* common.cuh */

struct ClsBatch {
  uint2 rec;       // this lane's record (a harmless one beyond n)
  uint32_t nf;     // flags | n << 8
  int32_t item;
};
__device__ __forceinline__ int cls_grab(int* counter, int lane) {  // inline PTX: the result is not needed right away
  int w = 0;
  if (lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(w) : "l"(counter) : "memory");
  return w;
}

template <int NV>
#ifdef CLS_MAXNREG
__global__ void __launch_bounds__(CLS_THREADS) __maxnreg__(CLS_MAXNREG) k_demux_cls(ClsArgs a) {
#else
__global__ void __launch_bounds__(CLS_THREADS, 1) k_demux_cls(ClsArgs a) {
#endif
  using Cfg = ClsCfg<NV>;
  constexpr int NE = Cfg::NE, ND = Cfg::ND, E_SG0 = Cfg::E_SG0, E_MX = Cfg::E_MX;
  extern __shared__ __align__(128) unsigned char cls_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    double* tabM = reinterpret_cast<double*>(cls_smem + Cfg::OFF_TABM);
    double* tabS = reinterpret_cast<double*>(cls_smem + Cfg::OFF_TABS);
    for (int i = tid; i < 3 * 64 * PSCL_FOLD_ROW; i += Cfg::THREADS) tabM[i] = a.fold_tab[i];
    for (int c = tid; c < 256; c += Cfg::THREADS) {
      const int cc = c < 3 * 64 ? c : 2 * 64;  // bytes beyond the table (class-D records) read the ones row
      const double t0 = a.fold_tab[cc * PSCL_FOLD_ROW], t4 = a.fold_tab[cc * PSCL_FOLD_ROW + 4];
      const double mx = cls_pmax(t0, t4), b = (t4 - t0) * 0.25;
      tabS[c * 4 + 0] = fma(1e-10, mx, t0);  // h[0] = f[0] + 1e-10*mx (:704-725 without the division)
      tabS[c * 4 + 1] = b;
      tabS[c * 4 + 2] = mx;
      tabS[c * 4 + 3] = b + b;
    }
  }
  {  // warm the L2 with the two genotype tables (evict_last): no demand miss tails in the first batches
    const size_t linesM = ((size_t)a.n_snps * Cfg::ROWB_M + 127) / 128, linesS = ((size_t)a.n_snps * Cfg::ROWB_S + 127) / 128;
    for (size_t i = (size_t)blockIdx.x * Cfg::THREADS + tid; i < linesM + linesS; i += (size_t)gridDim.x * Cfg::THREADS)
      cls_prefetch_l2_keep(i < linesM ? reinterpret_cast<const char*>(a.gpM) + i * 128 : reinterpret_cast<const char*>(a.gpS) + (i - linesM) * 128);
  }
  __syncthreads();
  const uint64_t pol_keep = cls_policy_evict_last(), pol_stream = cls_policy_evict_first();
  const double* const tabM = reinterpret_cast<const double*>(cls_smem + Cfg::OFF_TABM);
  const double4* const tabS = reinterpret_cast<const double4*>(cls_smem + Cfg::OFF_TABS);
  unsigned char* const rows0 = cls_smem + Cfg::OFF_ROWS + (size_t)warp * Cfg::WARP_B;  // [2][32][STRIDE]
  const uint32_t rows0_u32 = (uint32_t)__cvta_generic_to_shared(rows0);

  // gather geometry: LPR lanes per row, 32/LPR rows per instruction, LPR instructions per batch
  const int pieceM = lane % Cfg::LPR_M, rsubM = lane / Cfg::LPR_M;
  const int pieceS = lane % Cfg::LPR_S, rsubS = lane / Cfg::LPR_S;
  const char* const srcM = reinterpret_cast<const char*>(a.gpM) + pieceM * 16;
  const char* const srcS = reinterpret_cast<const char*>(a.gpS) + pieceS * 16;
  const uint32_t dstM = rsubM * Cfg::STRIDE + pieceM * 16, dstS = rsubS * Cfg::STRIDE + pieceS * 16;

  unsigned char* const ring = rows0 + Cfg::OFF_RING;  // [CLS_PF][CLS_SLOT_B]
  const uint32_t ring_u32 = rows0_u32 + Cfg::OFF_RING, bar_u32 = rows0_u32 + Cfg::OFF_BAR;
  if (lane == 0) {
    for (int i = 0; i < CLS_PF; ++i) cls_mbar_init(bar_u32 + i * 8, 1);
  }
  __syncwarp();

  // ---- work cursor: packets [pk, pk_end) of the current item; n1 = next item (descriptor in flight),
  //      n2 = the one after (work index in flight in lane 0's register).  A work item is a contiguous
  //      range of ready-made batch packets (k_dmx_pack), so the hot loop cuts nothing itself.
  int c_w = __shfl_sync(0xffffffffu, cls_grab(a.counter, lane), 0);
  int n1_w = __shfl_sync(0xffffffffu, cls_grab(a.counter, lane), 0);
  int n2_raw = cls_grab(a.counter, lane);
  uint4 n1_d = make_uint4(0, 0, 0, 0);
  uint32_t pk = 0, pk_end = 0;
  if (c_w < a.n_work) { const uint4 d = a.desc[c_w]; pk = d.x; pk_end = d.y; }
  if (n1_w < a.n_work) n1_d = a.desc[n1_w];
  uint32_t n_gen = 0, t_exit = 0xffffffffu;  // batches handed to the ring so far; first batch index past the work

  // start the bulk copy of the next packet into ring slot `slot` (every lane is done with its previous content)
  auto prefetch = [&](int slot) {
    while (c_w < a.n_work && pk >= pk_end) {  // item exhausted
      c_w = n1_w; pk = n1_d.x; pk_end = n1_d.y;
      n1_w = __shfl_sync(0xffffffffu, n2_raw, 0);
      if (n1_w < a.n_work) n1_d = a.desc[n1_w];
      n2_raw = cls_grab(a.counter, lane);
    }
    if (c_w < a.n_work) {
      if (lane == 0) {
        cls_mbar_expect_tx(bar_u32 + slot * 8, CLS_PKT_B);
        cls_bulk_g2s(ring_u32 + slot * CLS_SLOT_B, a.pkt + (size_t)pk * CLS_PKT_B, CLS_PKT_B, bar_u32 + slot * 8, pol_stream);
      }
      ++pk;
    } else {
      t_exit = min(t_exit, n_gen);
    }
    ++n_gen;
  };
  // batch t (ring slot t % CLS_PF) out of the ring: waits for its packet
  auto take = [&](uint32_t t, ClsBatch& b) {
    const int slot = t & (CLS_PF - 1);
    b.nf = CLS_FLAG_EXIT;
    b.item = 0;
    b.rec = make_uint2(0u, WS_NONE_CODES);
    if (t < t_exit) {
      cls_mbar_wait(bar_u32 + slot * 8, (t / CLS_PF) & 1u);
      const uint2 h = *reinterpret_cast<const uint2*>(ring + slot * CLS_SLOT_B);
      b.nf = h.x;
      b.item = (int)h.y;
      b.rec = *reinterpret_cast<const uint2*>(ring + slot * CLS_SLOT_B + 16 + lane * 8);
    }
  };
  // cooperative gather of the batch's 32 genotype rows into row buffer `buf`
  auto gather = [&](const ClsBatch& b, int buf) {
    const uint32_t base = rows0_u32 + buf * (32 * Cfg::STRIDE);
    const int snp = (int)b.rec.x;
    if (b.nf & CLS_FLAG_M) {
#pragma unroll
      for (int i = 0; i < Cfg::LPR_M; ++i) {
        const int snp_r = __shfl_sync(0xffffffffu, snp, i * (32 / Cfg::LPR_M) + rsubM);
        if (pieceM < Cfg::NCH_M)
          cls_cp_async16(base + dstM + i * (32 / Cfg::LPR_M) * Cfg::STRIDE, srcM + (size_t)(uint32_t)snp_r * Cfg::ROWB_M, pol_keep);
      }
    } else if (!(b.nf & CLS_FLAG_EXIT)) {
#pragma unroll
      for (int i = 0; i < Cfg::LPR_S; ++i) {
        const int snp_r = __shfl_sync(0xffffffffu, snp, i * (32 / Cfg::LPR_S) + rsubS);
        if (pieceS < Cfg::NCH_S)
          cls_cp_async16(base + dstS + i * (32 / Cfg::LPR_S) * Cfg::STRIDE, srcS + (size_t)(uint32_t)snp_r * Cfg::ROWB_S, pol_keep);
      }
    }
    cls_cp_async_commit();
  };

  double acc[NE];
  int ex[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) { acc[e] = 1.0; ex[e] = 0; }
  int n_has = 0, since = 0;

  // software pipeline: q[0] = batch being multiplied (rows in buffer t % CLS_NBUF), q[1..] = the next ones
  // (rows in flight / being gathered during this iteration); the packets of the CLS_PF batches after them
  // are in flight
#pragma unroll 1
  for (int s = 0; s < CLS_PF; ++s) prefetch(s);
  __syncwarp();
  constexpr int NB = CLS_NBUF;
  ClsBatch q[NB];
#pragma unroll
  for (int i = 0; i < NB - 1; ++i) { take(i, q[i]); gather(q[i], i); }
  int buf = 0;

#pragma unroll 1
  for (uint32_t t = 0; !(q[0].nf & CLS_FLAG_EXIT); ++t) {
    take(t + NB - 1, q[NB - 1]);
    __syncwarp();  // every lane is done with the row buffer of batch t-1 and with the ring slots read so far
    gather(q[NB - 1], buf == 0 ? NB - 1 : buf - 1);  // buffer (t + NB - 1) % NB
    prefetch(t & (CLS_PF - 1));                      // batch t + CLS_PF
    const ClsBatch bC = q[0];
    const uint32_t flags = bC.nf & 0xffu, nvalid = bC.nf >> 8;
    const uint2 rec = bC.rec;
    const double4 tb = tabS[rec.y & 0xffu];  // class S: {a + 1e-10*mx, b, mx, 2b}
    cls_cp_async_wait<NB - 1>();  // this lane's pieces of the current batch have landed ...
    __syncwarp();                 // ... and so have the other lanes' pieces of this lane's row
    const double2* const r2 = reinterpret_cast<const double2*>(rows0 + (size_t)buf * (32 * Cfg::STRIDE) + (size_t)lane * Cfg::STRIDE);

    if (!(flags & CLS_FLAG_M)) {
      // ---------------- class S: at most one usable base-call; h[i] = a1 + b*i ----------------
      double S[NV], M[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) { double2 v = r2[j]; S[j] = v.x; M[j] = v.y; }
      if (((uint32_t)lane < nvalid) && S[0] != -1.0) {  // -1: SNP without GP (cmd_cram_demuxlet.cpp:733)
        ++n_has;
        const double a1 = tb.x, b = tb.y;
        acc[E_MX] *= tb.z;
        // singlets: llksAB[j][0][0] (:806) = log((sum_l g_j[l] pG0[l]) * (sum_m g_0[m])), pG0[l] = h(2l)
        acc[E_SG0] *= S[0];
        double t[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) { t[j] = a1 * S[j]; acc[j] *= fma(tb.w, M[j], t[j]); }
        // doublets at alpha = 0.5: pG1[l][m] = h(l+m)
#pragma unroll
        for (int j = 1; j < NV; ++j) {
          const double u = fma(b, M[j], t[j]), w = b * S[j];
#pragma unroll
          for (int k = 0; k < j; ++k) acc[NV + j * (j - 1) / 2 + k] *= fma(M[k], w, S[k] * u);
        }
      }
    } else {
      // ---------------- classes M, D: Hankel form (as k_demux_default) --------------------------
      double G[NV][3];
      {
        double flat[Cfg::RM];
#pragma unroll
        for (int i = 0; i < Cfg::RM / 2; ++i) { double2 v = r2[i]; flat[2 * i] = v.x; flat[2 * i + 1] = v.y; }
#pragma unroll
        for (int j = 0; j < NV; ++j) { G[j][0] = flat[3 * j]; G[j][1] = flat[3 * j + 1]; G[j][2] = flat[3 * j + 2]; }
      }
      if (((uint32_t)lane < nvalid) && G[0][0] != -1.0) {
        ++n_has;
        double f0, f1, f2, f3, f4;
        if (flags & CLS_FLAG_D) {  // > 3 usable base-calls: folded at upload time (k_dmx_scatter)
          const double2* dp = reinterpret_cast<const double2*>(a.deep + (size_t)rec.y * 6);
          const double2 d0 = dp[0], d1 = dp[1], d2 = dp[2];
          f0 = d0.x; f1 = d0.y; f2 = d1.x; f3 = d1.y; f4 = d2.x;
        } else {
          const double* t0 = tabM + (rec.y & 0xffu) * PSCL_FOLD_ROW;
          const double* t1 = tabM + ((rec.y >> 8) & 0xffu) * PSCL_FOLD_ROW;
          const double* t2 = tabM + ((rec.y >> 16) & 0xffu) * PSCL_FOLD_ROW;
          f0 = t0[0] * t1[0] * t2[0]; f1 = t0[1] * t1[1] * t2[1]; f2 = t0[2] * t1[2] * t2[2];
          f3 = t0[3] * t1[3] * t2[3]; f4 = t0[4] * t1[4] * t2[4];
        }
        // D2 (:704-725) without the division: pG = (f/mx + 1e-10)/(1+1e-10) = h/(mx*(1+1e-10)) with
        // h = f + 1e-10*mx; mx is accumulated once per pair, (1+1e-10)^n_has is applied at the end
        const double mx = cls_pmax(cls_pmax(cls_pmax(f0, f1), cls_pmax(f2, f3)), f4);
        const double h0 = fma(1e-10, mx, f0), h1 = fma(1e-10, mx, f1), h2 = fma(1e-10, mx, f2),
                     h3 = fma(1e-10, mx, f3), h4 = fma(1e-10, mx, f4);
        acc[E_MX] *= mx;
        acc[E_SG0] *= (G[0][0] + G[0][1] + G[0][2]);
#pragma unroll
        for (int j = 0; j < NV; ++j) acc[j] *= (G[j][0] * h0 + G[j][1] * h2 + G[j][2] * h4);
#pragma unroll
        for (int j = 1; j < NV; ++j) {
          const double v0 = h0 * G[j][0] + h1 * G[j][1] + h2 * G[j][2];
          const double v1 = h1 * G[j][0] + h2 * G[j][1] + h3 * G[j][2];
          const double v2 = h2 * G[j][0] + h3 * G[j][1] + h4 * G[j][2];
#pragma unroll
          for (int k = 0; k < j; ++k)
            acc[NV + j * (j - 1) / 2 + k] *= (G[k][0] * v0 + G[k][1] * v1 + G[k][2] * v2);
        }
      }
    }

    const bool item_end = (flags & CLS_FLAG_END) != 0u;
    if (++since == 8 || item_end) {  // keep the running products inside the double range: a class-M term can be
                                     // as small as (10^-6.3 / 3)^3 ~ 5e-21 (three mismatches at phred 63)
      since = 0;
#pragma unroll
      for (int e = 0; e < NE; ++e) pscl_renorm(acc[e], ex[e]);
    }
    if (item_end) {
      // ---- item epilogue: warp transpose-reduce, then ONE log per accumulator ---------------------
      int nh = n_has;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nh += __shfl_xor_sync(0xffffffffu, nh, o);
      // round 1: accumulators 0..31 (lane L ends up with accumulator L); round 2: 32..39 (lane L with 32 + L/4)
      double m1[32], m2[8];
      int x1[32], x2[8];
#pragma unroll
      for (int e = 0; e < 32; ++e) { m1[e] = e < NE ? acc[e] : 1.0; x1[e] = e < NE ? ex[e] : 0; }
#pragma unroll
      for (int e = 0; e < 8; ++e) { m2[e] = 32 + e < NE ? acc[32 + e] : 1.0; x2[e] = 32 + e < NE ? ex[32 + e] : 0; }
      cls_transpose_prod<32>(m1, x1, lane);
      pscl_renorm(m1[0], x1[0]);
      const double lg1 = pscl_prod_log(m1[0], x1[0]);
      double lg2 = 0.0;
      if (NE > 32) {
        cls_transpose_prod<8>(m2, x2, lane);
        pscl_renorm(m2[0], x2[0]);
        lg2 = pscl_prod_log(m2[0], x2[0]);
      }
      auto pick = [&](int e) { return e < 32 ? __shfl_sync(0xffffffffu, lg1, e) : __shfl_sync(0xffffffffu, lg2, (e - 32) * 4); };
      // log(prod mx * (1+1e-10)^n_has): the pair normaliser of :704-725
      const double corr = pick(E_MX) + (double)nh * log1p(1e-10);
      const double sg0_log = pick(E_SG0);  // log prod (sum_m g_0[m]), the k=0 column factor of :806
      double* out = a.partial + (size_t)(bC.item - a.item_base) * (NV * NV * 2);
      auto store = [&](int e, double lg) {
        if (e < NV) out[(e * NV + 0) * 2 + 0] = lg - corr + sg0_log;
        else if (e < NV + ND) {
          int dd = e - NV, j = 1;
          while ((j + 1) * j / 2 <= dd) ++j;  // dd = j(j-1)/2 + k
          const int k = dd - j * (j - 1) / 2;
          const double x = lg - corr;
          out[(j * NV + k) * 2 + 1] = x;
          out[(k * NV + j) * 2 + 1] = x;
        }
      };
      store(lane, lg1);
      if (NE > 32 && (lane & 3) == 0) store(32 + (lane >> 2), lg2);
#pragma unroll
      for (int e = 0; e < NE; ++e) { acc[e] = 1.0; ex[e] = 0; }
      n_has = 0;
    }
#pragma unroll
    for (int i = 0; i < NB - 1; ++i) q[i] = q[i + 1];
    buf = buf == NB - 1 ? 0 : buf + 1;
  }
  cls_cp_async_wait<0>();
}

// ---- class-stream build (once per pileup image): k_dmx_classify lives in demux_poly.inl (k_demux_poly shares it) ----
// one warp per work item: records of a cell are written class S first, then M, then D, each in the
// original (ascending SNP) order; class-D pairs are folded here (cmd_cram_demuxlet.cpp:660-700)
__global__ void k_dmx_scatter(const int64_t* __restrict__ cell_ptr, const int32_t* __restrict__ item_cell,
                              const int64_t* __restrict__ item_pbeg, const int64_t* __restrict__ item_pend,
                              int32_t n_items, const uint2* __restrict__ rec_tmp, const unsigned long long* __restrict__ scan,
                              const uint32_t* __restrict__ pair_rd, const uint8_t* __restrict__ rd_aq,
                              const double* __restrict__ fold_tab, uint2* __restrict__ rec, double* __restrict__ deep) {
  const int item = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (item >= n_items) return;
  const int c = item_cell[item];
  const int64_t c0 = cell_ptr[c], c1 = cell_ptr[c + 1];
  const unsigned long long s0 = scan[c0], s1 = scan[c1];
  const uint32_t n_m = (uint32_t)(s1 - s0), n_d = (uint32_t)((s1 >> 32) - (s0 >> 32));
  const int64_t n_s = (c1 - c0) - n_m - n_d;
  for (int64_t p = item_pbeg[item] + lane; p < item_pend[item]; p += 32) {
    const unsigned long long sp = scan[p], kp = scan[p + 1] - sp;
    const uint32_t rank_m = (uint32_t)sp - (uint32_t)s0, rank_d = (uint32_t)(sp >> 32) - (uint32_t)(s0 >> 32);
    uint2 r = rec_tmp[p];
    int64_t pos;
    if (kp == 0ull) pos = c0 + ((p - c0) - rank_m - rank_d);
    else if (kp == 1ull) pos = c0 + n_s + rank_m;
    else {
      pos = c0 + n_s + n_m + rank_d;
      const uint32_t row = (uint32_t)(sp >> 32);
      double f0 = 1.0, f1 = 1.0, f2 = 1.0, f3 = 1.0, f4 = 1.0;
      uint32_t k = 0;
      for (uint32_t q = pair_rd[p]; q < pair_rd[p + 1]; ++q) {
        const double* t = fold_tab + (uint32_t)rd_aq[q] * PSCL_FOLD_ROW;  // allele-2 rows are all ones
        f0 *= t[0]; f1 *= t[1]; f2 *= t[2]; f3 *= t[3]; f4 *= t[4];
        if ((++k & 7u) == 0u) {  // deep pileups: rescale by the running max like :692-699
          const double ri = 1.0 / fmax(fmax(fmax(f0, f1), fmax(f2, f3)), f4);
          f0 *= ri; f1 *= ri; f2 *= ri; f3 *= ri; f4 *= ri;
        }
      }
      const double ri = 1.0 / fmax(fmax(fmax(f0, f1), fmax(f2, f3)), f4);
      double* d = deep + (size_t)row * 6;
      d[0] = f0 * ri; d[1] = f1 * ri; d[2] = f2 * ri; d[3] = f3 * ri; d[4] = f4 * ri; d[5] = 0.0;
      r.y = row;
    }
    rec[pos] = r;
  }
}

// (dmx_item_ranges lives in demux_poly.inl)
// number of 32-record packets of every work item (a packet never mixes classes)
__global__ void k_dmx_count_packets(const int64_t* __restrict__ cell_ptr, const int32_t* __restrict__ item_cell,
                                    const int64_t* __restrict__ item_pbeg, const int64_t* __restrict__ item_pend, int32_t n_items,
                                    const unsigned long long* __restrict__ scan, uint32_t* __restrict__ npk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_items) return;
  uint32_t n = 0;
  if (i < n_items) {
    const DmxItemRanges r = dmx_item_ranges(cell_ptr, item_cell, item_pbeg, item_pend, scan, i);
#pragma unroll
    for (int c = 0; c < 3; ++c) if (r.hi[c] > r.lo[c]) n += (r.hi[c] - r.lo[c] + 31u) / 32u;
  }
  npk[i] = n;
}
// one warp per work item: writes the item's packets (header + 32 records, idle slots filled with a
// harmless record) and its descriptor {first packet, end packet, item} in natural and in work order
__global__ void k_dmx_pack(const int64_t* __restrict__ cell_ptr, const int32_t* __restrict__ item_cell,
                           const int64_t* __restrict__ item_pbeg, const int64_t* __restrict__ item_pend,
                           const int32_t* __restrict__ item_order, int32_t n_items, const unsigned long long* __restrict__ scan,
                           const uint32_t* __restrict__ pk_off, const uint2* __restrict__ rec, unsigned char* __restrict__ pkt,
                           uint4* __restrict__ desc_nat, uint4* __restrict__ desc_sorted) {
  const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= n_items) return;
  if (lane == 0) {
    const int it = item_order[w];
    desc_nat[w] = make_uint4(pk_off[w], pk_off[w + 1], (uint32_t)w, 0u);
    desc_sorted[w] = make_uint4(pk_off[it], pk_off[it + 1], (uint32_t)it, 0u);
  }
  const DmxItemRanges r = dmx_item_ranges(cell_ptr, item_cell, item_pbeg, item_pend, scan, w);
  uint32_t q = pk_off[w];
  const uint32_t q_end = pk_off[w + 1];
  for (int c = 0; c < 3; ++c) {
    const uint32_t flags = c == 0 ? 0u : c == 1 ? CLS_FLAG_M : (CLS_FLAG_M | CLS_FLAG_D);
    for (uint32_t pos = r.lo[c]; pos < r.hi[c]; pos += 32, ++q) {
      const uint32_t n = min(32u, r.hi[c] - pos);
      unsigned char* dst = pkt + (size_t)q * CLS_PKT_B;
      if (lane == 0)
        *reinterpret_cast<uint4*>(dst) = make_uint4(flags | (q + 1 == q_end ? CLS_FLAG_END : 0u) | (n << 8), (uint32_t)w, 0u, 0u);
      reinterpret_cast<uint2*>(dst + 16)[lane] = (uint32_t)lane < n ? rec[pos + lane] : make_uint2(0u, WS_NONE_CODES);
    }
  }
}

// padded genotype rows and their (S, M) moments
__global__ void k_dmx_geno_tables(const double* __restrict__ gp, const uint8_t* __restrict__ has_gp, int32_t V, int32_t nv,
                                  int32_t RM, double* __restrict__ gpM, double* __restrict__ gpS) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)V * nv) return;
  const int64_t v = i / nv;
  const int j = (int)(i - v * nv);
  const double g0 = gp[i * 3], g1 = gp[i * 3 + 1], g2 = gp[i * 3 + 2];
  const bool absent = has_gp && !has_gp[v] && j == 0;  // the kernel tests the row's first value for -1
  double* m = gpM + v * RM + 3 * j;
  m[0] = absent ? -1.0 : g0; m[1] = g1; m[2] = g2;
  if (j == nv - 1 && RM > 3 * nv) m[3] = 0.0;
  gpS[(v * nv + j) * 2] = absent ? -1.0 : g0 + g1 + g2;
  gpS[(v * nv + j) * 2 + 1] = g1 + 2.0 * g2;
}

static int dmx_build_classes(pscl_ctx* ctx, pscl_plp* p) {
  if (p->dmx_pkt) return PSCL_OK;
  const int64_t P = p->P;
  const int32_t NI = p->n_items;
  if (P + 1 > INT32_MAX) return pscl_fail(ctx, PSCL_EINVAL, "k_demux_cls: a device pileup image holds < 2^31 pairs");
  uint2 *rec_tmp = nullptr, *rec = nullptr;
  unsigned long long *key = nullptr, *scan = nullptr;
  uint32_t *npk = nullptr, *pk_off = nullptr;
  void *tmp = nullptr, *tmp2 = nullptr;
  size_t tmp_bytes = 0, tmp2_bytes = 0;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** d, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(d, bytes ? bytes : 16); };
  alloc((void**)&rec_tmp, sizeof(uint2) * P);
  alloc((void**)&rec, sizeof(uint2) * P);
  alloc((void**)&key, sizeof(unsigned long long) * (P + 1));
  alloc((void**)&scan, sizeof(unsigned long long) * (P + 2));
  alloc((void**)&npk, sizeof(uint32_t) * (NI + 1));
  alloc((void**)&pk_off, sizeof(uint32_t) * (NI + 1));
  alloc((void**)&p->dmx_desc_nat, sizeof(uint4) * NI);
  alloc((void**)&p->dmx_desc_sorted, sizeof(uint4) * NI);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, key, scan, (int)(P + 1), ctx->stream);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tmp2_bytes, npk, pk_off, NI + 1, ctx->stream);
  alloc(&tmp, tmp_bytes);
  alloc(&tmp2, tmp2_bytes);
  if (e == cudaSuccess) e = cudaMemsetAsync(key, 0, sizeof(unsigned long long) * (P + 1), ctx->stream);
  if (e == cudaSuccess && P > 0) {
    k_dmx_classify<<<(unsigned)((P + 255) / 256), 256, 0, ctx->stream>>>(p->pair_snp, p->pair_rd, p->rd_aq, P, rec_tmp, key);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, key, scan, (int)(P + 1), ctx->stream);  // scan[P] = class totals
    ctx->launches++;
  }
  if (e == cudaSuccess) {
    k_dmx_count_packets<<<(unsigned)((NI + 1 + 255) / 256), 256, 0, ctx->stream>>>(p->cell_ptr, p->item_cell, p->item_pbeg, p->item_pend, NI, scan, npk);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp2, tmp2_bytes, npk, pk_off, NI + 1, ctx->stream);  // pk_off[NI] = packets
  ctx->launches += 2;
  unsigned long long totals = 0;
  uint32_t n_pkt = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&totals, scan + P, sizeof(totals), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&n_pkt, pk_off + NI, sizeof(n_pkt), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  const size_t n_deep = (size_t)(totals >> 32);
  alloc((void**)&p->dmx_deep, sizeof(double) * 6 * n_deep);
  alloc((void**)&p->dmx_pkt, (size_t)CLS_PKT_B * n_pkt);
  if (e == cudaSuccess && NI > 0) {
    k_dmx_scatter<<<(unsigned)(((int64_t)NI * 32 + 255) / 256), 256, 0, ctx->stream>>>(
        p->cell_ptr, p->item_cell, p->item_pbeg, p->item_pend, NI, rec_tmp, scan, p->pair_rd, p->rd_aq, ctx->fold_tab, rec, p->dmx_deep);
    k_dmx_pack<<<(unsigned)(((int64_t)NI * 32 + 255) / 256), 256, 0, ctx->stream>>>(
        p->cell_ptr, p->item_cell, p->item_pbeg, p->item_pend, p->item_order, NI, scan, pk_off, rec, p->dmx_pkt, p->dmx_desc_nat,
        p->dmx_desc_sorted);
    ctx->launches += 2;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(rec_tmp); cudaFree(rec); cudaFree(key); cudaFree(scan); cudaFree(npk); cudaFree(pk_off); cudaFree(tmp); cudaFree(tmp2);
  if (e != cudaSuccess) {
    cudaFree(p->dmx_pkt); cudaFree(p->dmx_deep); cudaFree(p->dmx_desc_nat); cudaFree(p->dmx_desc_sorted);
    p->dmx_pkt = nullptr; p->dmx_deep = nullptr; p->dmx_desc_nat = nullptr; p->dmx_desc_sorted = nullptr;
    return pscl_fail(ctx, e == cudaErrorMemoryAllocation ? PSCL_ENOMEM : PSCL_ECUDA, "demuxlet class-stream build failed: %s", cudaGetErrorString(e));
  }
  return PSCL_OK;
}

static int dmx_build_geno_tables(pscl_ctx* ctx) {
  if (ctx->gpM) return PSCL_OK;
  const int nv = ctx->nv, RM = (3 * nv + 1) & ~1;
  const int64_t V = ctx->geno_V;
  PSCL_CUDA(ctx, cudaMalloc((void**)&ctx->gpM, sizeof(double) * (size_t)std::max<int64_t>(V, 1) * RM));
  PSCL_CUDA(ctx, cudaMalloc((void**)&ctx->gpS, sizeof(double) * (size_t)std::max<int64_t>(V, 1) * 2 * nv));
  if (V > 0) {
    k_dmx_geno_tables<<<(unsigned)((V * nv + 255) / 256), 256, 0, ctx->stream>>>(ctx->gp, ctx->has_gp, (int32_t)V, nv, RM, ctx->gpM, ctx->gpS);
    ctx->launches++;
    PSCL_CUDA(ctx, cudaGetLastError());
  }
  return PSCL_OK;
}

template <int NV>
static cudaError_t launch_cls(pscl_ctx* ctx, const ClsArgs& a) {
  using Cfg = ClsCfg<NV>;
  static bool attr_set[64] = {false};
  if (!attr_set[ctx->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_demux_cls<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
    attr_set[ctx->device & 63] = true;
  }
  int grid = std::min(ctx->sm_count, (a.n_work + 7) / 8);
  if (grid < 1) grid = 1;
  k_demux_cls<NV><<<grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(a);
  return cudaGetLastError();
}
