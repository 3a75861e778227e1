// demux_ab.inl — k_demux_ab: the class-split demuxlet kernel (demux_cls.inl) with every batch of 32 pairs
// shared by TWO warps (part of popscle_b200.cu; default alpha grid {0, 0.5}, 2 <= nv <= 8).
//
// Why: k_demux_default / k_demux_cls keep nv + nv(nv-1)/2 + 2 running products per lane (38 at nv = 8),
// i.e. ~250 registers and 8 warps per SM, and at 8 warps per SM the per-pair gather of a genotype row out
// of L2 neither reaches its own throughput (tools/gather_bench.cu: 0.33 ms with 8 warps per SM, 0.24 ms
// with 16) nor overlaps with the arithmetic (profiles/r02_gather_bench.txt).  Here warp A of a pair owns
// the singlets and some doublet rows, warp B the other rows (ws_rows_a balances them), so each lane holds
// ~20 products in ~120 registers and an SM runs 16 warps.  The two warps read the same packet (one TMA
// bulk copy, issued by A) and the same gathered rows (issued alternately: warp t & 1 gathers batch t),
// one 64-thread named barrier per batch keeps them in step; exponents of the products live in shared memory.

#define AB_THREADS 512
#define AB_PF 8   /* packets in flight per warp pair (power of two) */

__host__ __device__ constexpr uint32_t ab_rows_a(int nv) {  // doublet rows j owned by warp A: greedy balance of
  int la = 4 * nv + 4, lb = 0;                               // the class-M instruction counts (row j: 4j+9)
  uint32_t m = 0;
  for (int j = nv - 1; j >= 1; --j) {
    int c = 4 * j + 9;
    if (la <= lb) { la += c; m |= 1u << j; } else lb += c;
  }
  return m;
}
__host__ __device__ constexpr bool ab_owns(int nv, int role, int j) { return (((ab_rows_a(nv) >> j) & 1u) != 0u) == (role == 0); }
// slot of an accumulator inside its role: 0 = pair normaliser, role A: 1 = k=0 column factor, 2.. = singlets;
// then the owned doublet rows in (j,k) order
__host__ __device__ constexpr int ab_dbl_base(int nv, int role) { return role == 0 ? nv + 2 : 1; }
__host__ __device__ constexpr int ab_dbl_slot(int nv, int role, int j, int k) {
  int s = ab_dbl_base(nv, role);
  for (int jj = 1; jj < j; ++jj) if (ab_owns(nv, role, jj)) s += jj;
  return s + k;
}
__host__ __device__ constexpr int ab_nslot(int nv, int role) { return ab_dbl_slot(nv, role, nv, 0); }

template <int NV>
struct AbCfg {
  using C = ClsCfg<NV>;
  static constexpr int NSLOT = ab_nslot(NV, 0) > ab_nslot(NV, 1) ? ab_nslot(NV, 0) : ab_nslot(NV, 1);
  // per warp pair: two genotype-row buffers | packet ring | mbarriers
  static constexpr int OFF_RING = 2 * 32 * C::STRIDE;
  static constexpr int OFF_BAR = OFF_RING + AB_PF * CLS_SLOT_B;
  static constexpr int PAIR_B = (OFF_BAR + AB_PF * 8 + 127) & ~127;
  static constexpr int OFF_PAIRS = C::OFF_ROWS;
  static constexpr int OFF_EX = OFF_PAIRS + (AB_THREADS / 64) * PAIR_B;  // int [NSLOT][AB_THREADS]
  static constexpr size_t SMEM = (size_t)OFF_EX + (size_t)NSLOT * AB_THREADS * 4;
  static_assert(NSLOT <= 32 && SMEM <= 227 * 1024, "slots / shared memory budget");
};

__device__ __forceinline__ void ab_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void ab_mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.shared::cta.b64 t, [%0];\n\t}" ::"r"(bar) : "memory");
}

template <int NV, int ROLE>
__device__ __forceinline__ void ab_warp(const ClsArgs& a, unsigned char* smem, const int pw, const int lane, const int tid) {
  using Cfg = ClsCfg<NV>;
  using Ab = AbCfg<NV>;
  constexpr int NE = Cfg::NE, E_SG0 = Cfg::E_SG0, E_MX = Cfg::E_MX;
  constexpr int NS = ab_nslot(NV, ROLE), DB = ab_dbl_base(NV, ROLE);
#define AB_OWN_ROW(j) ab_owns(NV, ROLE, (j))
  const double* const tabM = reinterpret_cast<const double*>(smem + Cfg::OFF_TABM);
  const double4* const tabS = reinterpret_cast<const double4*>(smem + Cfg::OFF_TABS);
  unsigned char* const rows0 = smem + Ab::OFF_PAIRS + (size_t)pw * Ab::PAIR_B;  // [2][32][STRIDE] | ring | barriers
  const uint32_t rows0_u32 = (uint32_t)__cvta_generic_to_shared(rows0);
  unsigned char* const ring = rows0 + Ab::OFF_RING;
  const uint32_t ring_u32 = rows0_u32 + Ab::OFF_RING, bar_u32 = rows0_u32 + Ab::OFF_BAR;
  int* const s_ex = reinterpret_cast<int*>(smem + Ab::OFF_EX);  // [slot][tid]
  const uint64_t pol_keep = cls_policy_evict_last(), pol_stream = cls_policy_evict_first();

  const int pieceM = lane % Cfg::LPR_M, rsubM = lane / Cfg::LPR_M;
  const int pieceS = lane % Cfg::LPR_S, rsubS = lane / Cfg::LPR_S;
  const char* const srcM = reinterpret_cast<const char*>(a.gpM) + pieceM * 16;
  const char* const srcS = reinterpret_cast<const char*>(a.gpS) + pieceS * 16;
  const uint32_t dstM = rsubM * Cfg::STRIDE + pieceM * 16, dstS = rsubS * Cfg::STRIDE + pieceS * 16;

  // ---- work cursor (warp A only): packets [pk, pk_end) of the current item, next item's descriptor in flight
  int c_w = 0, n1_w = 0, n2_raw = 0;
  uint4 n1_d = make_uint4(0, 0, 0, 0);
  uint32_t pk = 0, pk_end = 0;
  if (ROLE == 0) {
    c_w = __shfl_sync(0xffffffffu, cls_grab(a.counter, lane), 0);
    n1_w = __shfl_sync(0xffffffffu, cls_grab(a.counter, lane), 0);
    n2_raw = cls_grab(a.counter, lane);
    if (c_w < a.n_work) { const uint4 d = a.desc[c_w]; pk = d.x; pk_end = d.y; }
    if (n1_w < a.n_work) n1_d = a.desc[n1_w];
  }
  // warp A: start the bulk copy of the next packet into ring slot `slot`, or park an EXIT header there
  auto prefetch = [&](int slot) {
    while (c_w < a.n_work && pk >= pk_end) {  // item exhausted
      c_w = n1_w; pk = n1_d.x; pk_end = n1_d.y;
      n1_w = __shfl_sync(0xffffffffu, n2_raw, 0);
      if (n1_w < a.n_work) n1_d = a.desc[n1_w];
      n2_raw = cls_grab(a.counter, lane);
    }
    if (lane == 0) {
      if (c_w < a.n_work) {
        cls_mbar_expect_tx(bar_u32 + slot * 8, CLS_PKT_B);
        cls_bulk_g2s(ring_u32 + slot * CLS_SLOT_B, a.pkt + (size_t)pk * CLS_PKT_B, CLS_PKT_B, bar_u32 + slot * 8, pol_stream);
      } else {  // out of work: both warps find an EXIT header behind a completed barrier phase
        *reinterpret_cast<uint2*>(ring + slot * CLS_SLOT_B) = make_uint2(CLS_FLAG_EXIT, 0u);
        ab_mbar_arrive(bar_u32 + slot * 8);
      }
    }
    if (c_w < a.n_work) ++pk;
  };
  auto take = [&](uint32_t t, ClsBatch& b) {
    const int slot = t & (AB_PF - 1);
    cls_mbar_wait(bar_u32 + slot * 8, (t / AB_PF) & 1u);
    const uint2 h = *reinterpret_cast<const uint2*>(ring + slot * CLS_SLOT_B);
    b.nf = h.x;
    b.item = (int)h.y;
    b.rec = make_uint2(0u, WS_NONE_CODES);
    if (!(h.x & CLS_FLAG_EXIT)) b.rec = *reinterpret_cast<const uint2*>(ring + slot * CLS_SLOT_B + 16 + lane * 8);
  };
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
#pragma unroll
  for (int e = 0; e < NE; ++e) acc[e] = 1.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) s_ex[q * AB_THREADS + tid] = 0;
  int n_has = 0, since = 0;
  // renormalise every owned product, its exponent goes to shared memory
  auto renorm_all = [&]() {
    { int x = 0; pscl_renorm(acc[E_MX], x); s_ex[0 * AB_THREADS + tid] += x; }
    if (ROLE == 0) {
      { int x = 0; pscl_renorm(acc[E_SG0], x); s_ex[1 * AB_THREADS + tid] += x; }
#pragma unroll
      for (int j = 0; j < NV; ++j) { int x = 0; pscl_renorm(acc[j], x); s_ex[(2 + j) * AB_THREADS + tid] += x; }
    }
#pragma unroll
    for (int j = 1; j < NV; ++j) {
      if (AB_OWN_ROW(j)) {
#pragma unroll
        for (int k = 0; k < j; ++k) { int x = 0; pscl_renorm(acc[NV + j * (j - 1) / 2 + k], x); s_ex[ab_dbl_slot(NV, ROLE, j, k) * AB_THREADS + tid] += x; }
      }
    }
  };

  if (ROLE == 0) {
#pragma unroll 1
    for (int s = 0; s < AB_PF; ++s) prefetch(s);
  }
  ClsBatch bC, bB;
  take(0, bC);
  if (ROLE == 0) gather(bC, 0);

#pragma unroll 1
  for (uint32_t t = 0; !(bC.nf & CLS_FLAG_EXIT); ++t) {
    const int buf = t & 1;
    if (ROLE == (int)(t & 1u)) cls_cp_async_wait<0>();  // the rows of batch t were this warp's gather: they have landed
    ab_bar_sync(pw + 1);  // rows of batch t visible to both warps; both are done with batch t-1 (its rows, its ring slot)
    take(t + 1, bB);
    if (ROLE == (int)((t + 1u) & 1u)) gather(bB, buf ^ 1);
    if (ROLE == 0) prefetch(t & (AB_PF - 1));  // batch t + AB_PF into the slot of batch t (both warps took it before the barrier)
    const uint32_t flags = bC.nf & 0xffu, nvalid = bC.nf >> 8;
    const uint2 rec = bC.rec;
    const double2* const r2 = reinterpret_cast<const double2*>(rows0 + (size_t)buf * (32 * Cfg::STRIDE) + (size_t)lane * Cfg::STRIDE);

    if (!(flags & CLS_FLAG_M)) {
      // ---------------- class S: at most one usable base-call; h[i] = a1 + b*i ----------------
      double S[NV], M[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) { double2 v = r2[j]; S[j] = v.x; M[j] = v.y; }
      const double4 tb = tabS[rec.y & 0xffu];  // {a + 1e-10*mx, b, mx, 2b}
      if (((uint32_t)lane < nvalid) && S[0] != -1.0) {  // -1: SNP without GP (cmd_cram_demuxlet.cpp:733)
        ++n_has;
        const double a1 = tb.x, b = tb.y;
        acc[E_MX] *= tb.z;
        if (ROLE == 0) {
          acc[E_SG0] *= S[0];
#pragma unroll
          for (int j = 0; j < NV; ++j) acc[j] *= fma(tb.w, M[j], a1 * S[j]);
        }
#pragma unroll
        for (int j = 1; j < NV; ++j) {
          if (AB_OWN_ROW(j)) {
            const double u = fma(b, M[j], a1 * S[j]), w = b * S[j];
#pragma unroll
            for (int k = 0; k < j; ++k) acc[NV + j * (j - 1) / 2 + k] *= fma(M[k], w, S[k] * u);
          }
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
        const double mx = cls_pmax(cls_pmax(cls_pmax(f0, f1), cls_pmax(f2, f3)), f4);
        const double h0 = fma(1e-10, mx, f0), h1 = fma(1e-10, mx, f1), h2 = fma(1e-10, mx, f2),
                     h3 = fma(1e-10, mx, f3), h4 = fma(1e-10, mx, f4);
        acc[E_MX] *= mx;
        if (ROLE == 0) {
          acc[E_SG0] *= (G[0][0] + G[0][1] + G[0][2]);
#pragma unroll
          for (int j = 0; j < NV; ++j) acc[j] *= (G[j][0] * h0 + G[j][1] * h2 + G[j][2] * h4);
        }
#pragma unroll
        for (int j = 1; j < NV; ++j) {
          if (AB_OWN_ROW(j)) {
            const double v0 = h0 * G[j][0] + h1 * G[j][1] + h2 * G[j][2];
            const double v1 = h1 * G[j][0] + h2 * G[j][1] + h3 * G[j][2];
            const double v2 = h2 * G[j][0] + h3 * G[j][1] + h4 * G[j][2];
#pragma unroll
            for (int k = 0; k < j; ++k)
              acc[NV + j * (j - 1) / 2 + k] *= (G[k][0] * v0 + G[k][1] * v1 + G[k][2] * v2);
          }
        }
      }
    }

    const bool item_end = (flags & CLS_FLAG_END) != 0u;
    if (++since == 8 || item_end) { since = 0; renorm_all(); }
    if (item_end) {
      // ---- item epilogue: warp transpose-reduce of this role's slots, ONE log per accumulator ----------
      int nh = n_has;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nh += __shfl_xor_sync(0xffffffffu, nh, o);
      double m1[32];
      int x1[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) { m1[q] = 1.0; x1[q] = q < NS ? s_ex[q * AB_THREADS + tid] : 0; }
      m1[0] = acc[E_MX];
      if (ROLE == 0) {
        m1[1] = acc[E_SG0];
#pragma unroll
        for (int j = 0; j < NV; ++j) m1[2 + j] = acc[j];
      }
#pragma unroll
      for (int j = 1; j < NV; ++j) {
        if (AB_OWN_ROW(j)) {
#pragma unroll
          for (int k = 0; k < j; ++k) m1[ab_dbl_slot(NV, ROLE, j, k)] = acc[NV + j * (j - 1) / 2 + k];
        }
      }
      cls_transpose_prod<32>(m1, x1, lane);  // lane L now holds slot L
      pscl_renorm(m1[0], x1[0]);
      const double lg = pscl_prod_log(m1[0], x1[0]);
      // log(prod mx * (1+1e-10)^n_has): the pair normaliser of :704-725 (identical in both warps)
      const double corr = __shfl_sync(0xffffffffu, lg, 0) + (double)nh * log1p(1e-10);
      const double sg0_log = (ROLE == 0) ? __shfl_sync(0xffffffffu, lg, 1) : 0.0;  // the k=0 column factor of :806
      double* out = a.partial + (size_t)(bC.item - a.item_base) * (NV * NV * 2);
      if (ROLE == 0 && lane >= 2 && lane < 2 + NV) out[((lane - 2) * NV + 0) * 2 + 0] = lg - corr + sg0_log;
      if (lane >= DB && lane < NS) {
        int d = lane - DB, jj = 0, kk = 0;
#pragma unroll 1
        for (int j = 1; j < NV; ++j) {
          if (ab_owns(NV, ROLE, j)) {
            if (d < j) { jj = j; kk = d; break; }
            d -= j;
          }
        }
        const double x = lg - corr;
        out[(jj * NV + kk) * 2 + 1] = x;
        out[(kk * NV + jj) * 2 + 1] = x;
      }
#pragma unroll
      for (int e = 0; e < NE; ++e) acc[e] = 1.0;
#pragma unroll
      for (int q = 0; q < NS; ++q) s_ex[q * AB_THREADS + tid] = 0;
      n_has = 0;
    }
    bC = bB;
  }
  cls_cp_async_wait<0>();
#undef AB_OWN_ROW
}

template <int NV>
__global__ void __launch_bounds__(AB_THREADS, 1) k_demux_ab(ClsArgs a) {
  using Cfg = ClsCfg<NV>;
  using Ab = AbCfg<NV>;
  extern __shared__ __align__(128) unsigned char ab_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    double* tabM = reinterpret_cast<double*>(ab_smem + Cfg::OFF_TABM);
    double* tabS = reinterpret_cast<double*>(ab_smem + Cfg::OFF_TABS);
    for (int i = tid; i < 3 * 64 * PSCL_FOLD_ROW; i += AB_THREADS) tabM[i] = a.fold_tab[i];
    for (int c = tid; c < 256; c += AB_THREADS) {
      const int cc = c < 3 * 64 ? c : 2 * 64;  // bytes beyond the table (class-D records) read the ones row
      const double t0 = a.fold_tab[cc * PSCL_FOLD_ROW], t4 = a.fold_tab[cc * PSCL_FOLD_ROW + 4];
      const double mx = cls_pmax(t0, t4), b = (t4 - t0) * 0.25;
      tabS[c * 4 + 0] = fma(1e-10, mx, t0);
      tabS[c * 4 + 1] = b;
      tabS[c * 4 + 2] = mx;
      tabS[c * 4 + 3] = b + b;
    }
    if ((warp & 1) == 0 && lane == 0) {  // the pair's mbarriers
      const uint32_t bar = (uint32_t)__cvta_generic_to_shared(ab_smem + Ab::OFF_PAIRS + (size_t)(warp >> 1) * Ab::PAIR_B) + Ab::OFF_BAR;
      for (int i = 0; i < AB_PF; ++i) cls_mbar_init(bar + i * 8, 1);
    }
    // warm the L2 with the two genotype tables (evict_last)
    const size_t linesM = ((size_t)a.n_snps * Cfg::ROWB_M + 127) / 128, linesS = ((size_t)a.n_snps * Cfg::ROWB_S + 127) / 128;
    for (size_t i = (size_t)blockIdx.x * AB_THREADS + tid; i < linesM + linesS; i += (size_t)gridDim.x * AB_THREADS)
      cls_prefetch_l2_keep(i < linesM ? reinterpret_cast<const char*>(a.gpM) + i * 128 : reinterpret_cast<const char*>(a.gpS) + (i - linesM) * 128);
  }
  __syncthreads();
  if (warp & 1) ab_warp<NV, 1>(a, ab_smem, warp >> 1, lane, tid);
  else ab_warp<NV, 0>(a, ab_smem, warp >> 1, lane, tid);
}

template <int NV>
static cudaError_t launch_ab(pscl_ctx* ctx, const ClsArgs& a) {
  using Ab = AbCfg<NV>;
  static bool attr_set[64] = {false};
  if (!attr_set[ctx->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_demux_ab<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Ab::SMEM);
    if (e != cudaSuccess) return e;
    attr_set[ctx->device & 63] = true;
  }
  int grid = std::min(ctx->sm_count, (a.n_work + 7) / 8);
  if (grid < 1) grid = 1;
  k_demux_ab<NV><<<grid, AB_THREADS, Ab::SMEM, ctx->stream>>>(a);
  return cudaGetLastError();
}
