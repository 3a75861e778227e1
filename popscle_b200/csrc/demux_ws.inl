// demux_ws.inl — warp-specialised demuxlet kernel for the default alpha grid {0, 0.5}, 2 <= nv <= 8
// (part of popscle_b200.cu; replaces cmd_cram_demuxlet.cpp:655-747 like k_demux_default, which is
// kept as the A/B baseline).
//
// Why a second formulation.  ncu of k_demux_default<8> (profiles/r01a_*) shows an FP64-issue-bound
// kernel at 2 warps per scheduler whose integer phases (index chasing, row gather) and FP64 phases
// do not overlap.  This kernel attacks both terms:
//
//  1. Fewer FP64 instructions.  For a (cell,SNP) pair with at most one usable base-call — three
//     quarters of all pairs in droplet data — the per-read factor pR(1-p)+pA*p (:685) is LINEAR in
//     the mixing fraction p = (l+m)/4, so with h[i] = a + b*i
//         sum_{l,m} g_j[l] g_k[m] h[l+m] = S_k (a S_j + b M_j) + M_k (b S_j),
//         S_j = sum_l g_j[l],  M_j = sum_l l g_j[l]
//     i.e. 2 FP64 instructions per doublet instead of 3 (+3 instead of 9 per sample), on a
//     (S,M) genotype table of 16 B per sample instead of 24 B.  S_j is carried exactly (it is 1 only
//     up to the float32 rounding of the VCF values), so nothing is approximated.  Pairs with >= 2
//     usable base-calls ("class M": 2-3, folded from a shared-memory table; "class D": > 3, folded
//     once per upload into a 5-value table) keep the Hankel form of k_demux_default.  The pileup
//     image is re-ordered once per upload into 8-byte records, class S first, then M, then D inside
//     each cell (k_dmx_classify / k_dmx_scatter), so a warp never mixes the forms and the hot loop
//     never chases a read offset.
//  2. Overlap.  Each CTA runs 4 channels of {1 producer warp, 2 consumer warps}.  The producer
//     walks the record stream (cp.async into a shared-memory ring, 8 batches ahead), gathers the 32
//     genotype rows of a batch with cp.async straight into a shared-memory stage and signals an
//     mbarrier (cp.async.mbarrier.arrive.noinc); the two consumers split the nv + nv(nv-1)/2 running
//     products between them (role A: singlets + some doublet rows, role B: the other rows), so each
//     holds ~20 accumulators, and fetch the next batch's header while they multiply the current one.
//     setmaxnreg moves registers from the producer warpgroup to the consumer warpgroups.  All three
//     roles are rolled loops: together they have to live in the SM's 32 KB instruction cache (an
//     unrolled first version spent half its time in stall_no_inst).
//
// Arithmetic conventions (closed-form pG, mantissa/exponent running products, one log per
// accumulator per work item, alpha = 0.5 mirror) are those of k_demux_default; see demux.inl.

#include <cub/device/device_scan.cuh>

#define WS_NSTAGE 4 /* stages per channel (power of two) */
#define WS_PF 8     /* record batches in flight per producer (power of two) */
#define WS_FLAG_M 1u    /* Hankel form (classes M and D) */
#define WS_FLAG_END 2u  /* last batch of its work item */
#define WS_FLAG_EXIT 4u
#define WS_FLAG_D 8u    /* class D: the fold comes from the deep table */
#define WS_NONE_CODES 0x00808080u /* three "no base-call" codes: allele 2, qual 0 = the all-ones row of both fold tables */

struct WsArgs {
  const uint2* rec;          // [P] class-ordered records: S/M {snp, b0 | b1<<8 | b2<<16 | cnt<<24}, D {snp, deep row}
  const double* deep;        // [n_deep][6] folded per-read factors f0..f4 (max 1) of the class-D pairs
  const double* gpM;         // [V][RM]   genotype rows, padded to 16 B; first value -1 = SNP without GP
  const double* gpS;         // [V][2 nv] (S_j, M_j);                   first value -1 = SNP without GP
  const double* fold_tab;    // [3*64][PSCL_FOLD_ROW]
  const uint4* desc;         // [n_work] {begin, end, first M, first D} record positions of a work item
  const int32_t* desc_item;  // [n_work] item id, or null (item = item_base + w)
  double* partial;           // [items][nv*nv*2], row (item - item_base)
  int* counter;
  int32_t item_base, n_work;
};

__host__ __device__ constexpr int ws_pow2ceil(int x) { return x <= 1 ? 1 : x <= 2 ? 2 : x <= 4 ? 4 : x <= 8 ? 8 : x <= 16 ? 16 : 32; }
// doublet rows j (accumulators (j,k<j)) owned by role A; greedy balance of the class-M instruction
// counts (row j: 4j+9; role A starts with the singlets' 4nv+4)
__host__ __device__ constexpr uint32_t ws_rows_a(int nv) {
  int la = 4 * nv + 4, lb = 0;
  uint32_t m = 0;
  for (int j = nv - 1; j >= 1; --j) {
    int c = 4 * j + 9;
    if (la <= lb) { la += c; m |= 1u << j; } else lb += c;
  }
  return m;
}
__host__ __device__ constexpr bool ws_owns(int nv, int role, int j) { return (((ws_rows_a(nv) >> j) & 1u) != 0u) == (role == 0); }
// epilogue slot of an accumulator inside its role: 0 = pair normaliser, role A: 1 = k=0 column factor,
// 2.. = singlets; then the owned doublet rows in (j,k) order
__host__ __device__ constexpr int ws_dbl_base(int nv, int role) { return role == 0 ? nv + 2 : 1; }
__host__ __device__ constexpr int ws_dbl_slot(int nv, int role, int j, int k) {
  int s = ws_dbl_base(nv, role);
  for (int jj = 1; jj < j; ++jj) if (ws_owns(nv, role, jj)) s += jj;
  return s + k;
}
__host__ __device__ constexpr int ws_nslot(int nv, int role) { return ws_dbl_slot(nv, role, nv, 0); }

template <int NV>
struct WsCfg {
  static constexpr int ND = NV * (NV - 1) / 2;
  static constexpr int NE = NV + ND + 2;
  static constexpr int E_SG0 = NV + ND, E_MX = NV + ND + 1;
  static constexpr int RM = (3 * NV + 1) & ~1;  // doubles per class-M row
  static constexpr int RS = 2 * NV;             // doubles per class-S row
  static constexpr int ROWB_M = RM * 8, ROWB_S = RS * 8;
  static constexpr int STRIDE = ((ROWB_M / 16) | 1) * 16;  // odd multiple of 16 B: conflict-free LDS.128 per lane
  static constexpr int NCH_M = ROWB_M / 16, NCH_S = ROWB_S / 16;
  static constexpr int LPR_M = ws_pow2ceil(NCH_M), LPR_S = ws_pow2ceil(NCH_S);
  static constexpr int OFF_REC = 32 * STRIDE, OFF_HDR = OFF_REC + 32 * 8;
  static constexpr int STAGE_B = OFF_HDR + 16;
  static constexpr int NSLOT = ws_nslot(NV, 0) > ws_nslot(NV, 1) ? ws_nslot(NV, 0) : ws_nslot(NV, 1);
  static constexpr int RING_B = WS_PF * (256 + 16);  // per channel: record ring + batch descriptor ring
  static constexpr int SCR_B = NSLOT * 33 * 12;      // per consumer warp: [NSLOT][33] mantissas + exponents
  static constexpr int OFF_FULL = 0, OFF_EMPTY = 128, OFF_TABM = 256;
  static constexpr int OFF_TABS = OFF_TABM + 3 * 64 * PSCL_FOLD_ROW * 8;
  static constexpr int OFF_STAGE = OFF_TABS + 256 * 4 * 8;  // tabS is indexed by a raw record byte
  static constexpr int OFF_RING = OFF_STAGE + 4 * WS_NSTAGE * STAGE_B;
  static constexpr int OFF_SCR = OFF_RING + 4 * RING_B;
  static constexpr size_t SMEM = (size_t)OFF_SCR + (size_t)8 * ((SCR_B + 15) & ~15);
  static_assert(WS_NSTAGE * 4 * 8 <= 128 && (WS_NSTAGE & (WS_NSTAGE - 1)) == 0 && (WS_PF & (WS_PF - 1)) == 0, "barrier block / rings");
  static_assert(NCH_M <= 32 && STAGE_B % 16 == 0 && NSLOT <= 32, "stage geometry");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ws_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.shared::cta.b64 t, [%0];\n\t}" ::"r"(bar) : "memory");
}
// blocking wait: try_wait suspends the warp in hardware (up to the hint, woken by the completing
// arrive), so waiting warps do not eat issue slots of the working ones
__device__ __forceinline__ void ws_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
  } while (!ok);
}
__device__ __forceinline__ bool ws_mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void ws_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ws_cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ws_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ws_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ws_cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ double ws_pmax(double x, double y) { return x > y ? x : y; }  // positive, non-NaN operands

// ---- producer --------------------------------------------------------------------------------------
// One rolled loop: batch t's records were copied into the shared-memory ring WS_PF iterations earlier
// by cp.async, its descriptor sits beside them.
template <int NV>
__device__ __forceinline__ void ws_producer(const WsArgs& a, unsigned char* smem, const int ch, const int lane) {
  using Cfg = WsCfg<NV>;
  const uint32_t full0 = ws_smem_u32(smem + Cfg::OFF_FULL) + ch * WS_NSTAGE * 8;
  const uint32_t empty0 = ws_smem_u32(smem + Cfg::OFF_EMPTY) + ch * WS_NSTAGE * 8;
  unsigned char* const stage0 = smem + Cfg::OFF_STAGE + (size_t)ch * WS_NSTAGE * Cfg::STAGE_B;
  unsigned char* const ring = smem + Cfg::OFF_RING + (size_t)ch * Cfg::RING_B;  // [WS_PF][32] uint2
  uint4* const dring = reinterpret_cast<uint4*>(ring + WS_PF * 256);            // [WS_PF] {flags | n<<8, -, item, -}

  // gather geometry: LPR lanes per row, 32/LPR rows per instruction, LPR instructions per batch
  const int pieceM = lane % Cfg::LPR_M, rsubM = lane / Cfg::LPR_M;
  const int pieceS = lane % Cfg::LPR_S, rsubS = lane / Cfg::LPR_S;
  const char* const srcM = reinterpret_cast<const char*>(a.gpM) + pieceM * 16;
  const char* const srcS = reinterpret_cast<const char*>(a.gpS) + pieceS * 16;
  const uint32_t dstM = rsubM * Cfg::STRIDE + pieceM * 16, dstS = rsubS * Cfg::STRIDE + pieceS * 16;

  // ---- item cursor: c = item being cut into batches, n1 = next (descriptor in flight),
  //      n2 = the one after (work index in flight in lane 0's register)
  auto grab = [&]() { int w = 0; if (lane == 0) w = atomicAdd(a.counter, 1); return w; };
  auto load_desc = [&](int w, uint4& d, int& item) {
    if (w < a.n_work) {
      d = a.desc[w];
      item = a.desc_item ? a.desc_item[w] : a.item_base + w;
    }
  };
  int c_w = __shfl_sync(0xffffffffu, grab(), 0);
  int n1_w = __shfl_sync(0xffffffffu, grab(), 0);
  int n2_raw = grab();
  uint4 c_d = make_uint4(0, 0, 0, 0), n1_d = make_uint4(0, 0, 0, 0);
  int c_item = 0, n1_item = 0;
  load_desc(c_w, c_d, c_item);
  load_desc(n1_w, n1_d, n1_item);
  uint32_t pos = c_d.x;

  // cut the next batch off the cursor, park its descriptor and start the copy of its records.  The
  // records of an item are one contiguous range [begin, end) whose classes are S | M | D, so the
  // class of a batch is the class of its first record and a batch ends at the next class boundary.
  auto prefetch = [&](int slot) {
    uint32_t flags = WS_FLAG_EXIT, n = 0, idx0 = 0;
    int item = 0;
    while (c_w < a.n_work && pos >= c_d.y) {  // item exhausted
      c_w = n1_w; c_d = n1_d; c_item = n1_item;
      n1_w = __shfl_sync(0xffffffffu, n2_raw, 0);
      load_desc(n1_w, n1_d, n1_item);
      n2_raw = grab();
      pos = c_d.x;
    }
    if (c_w < a.n_work) {
      uint32_t bound = c_d.y;
      if (pos < c_d.z) { flags = 0u; bound = min(bound, c_d.z); }
      else if (pos < c_d.w) { flags = WS_FLAG_M; bound = min(bound, c_d.w); }
      else flags = WS_FLAG_M | WS_FLAG_D;
      idx0 = pos;
      n = min(32u, bound - pos);
      pos += n;
      if (pos >= c_d.y) flags |= WS_FLAG_END;
      item = c_item;
    }
    if (lane == 0) dring[slot] = make_uint4(flags | (n << 8), 0u, (uint32_t)item, 0u);
    unsigned char* dst = ring + slot * 256 + lane * 8;
    if ((uint32_t)lane < n) ws_cp_async8(ws_smem_u32(dst), a.rec + idx0 + lane);
    else *reinterpret_cast<uint2*>(dst) = make_uint2(0u, WS_NONE_CODES);  // SNP 0, no base-calls: a harmless row
    ws_cp_async_commit();
  };
#pragma unroll 1
  for (int s = 0; s < WS_PF; ++s) prefetch(s);

#pragma unroll 1
  for (uint32_t t = 0;; ++t) {
    const int slot = t & (WS_PF - 1), stage = t & (WS_NSTAGE - 1);
    ws_cp_async_wait<WS_PF - 1>();  // the group that carries batch t's records has landed
    __syncwarp();
    const uint4 d = dring[slot];
    const uint2 rec = *reinterpret_cast<const uint2*>(ring + slot * 256 + lane * 8);
    ws_mbar_wait(empty0 + stage * 8, ((t / WS_NSTAGE) & 1u) ^ 1u);  // a fresh mbarrier passes a parity-1 wait
    unsigned char* st = stage0 + (size_t)stage * Cfg::STAGE_B;
    const uint32_t st_u32 = ws_smem_u32(st);
    *reinterpret_cast<uint2*>(st + Cfg::OFF_REC + lane * 8) = rec;
    const int snp = (int)rec.x;
    if (d.x & WS_FLAG_M) {
#pragma unroll
      for (int i = 0; i < Cfg::LPR_M; ++i) {
        const int snp_r = __shfl_sync(0xffffffffu, snp, i * (32 / Cfg::LPR_M) + rsubM);
        if (pieceM < Cfg::NCH_M)
          ws_cp_async16(st_u32 + dstM + i * (32 / Cfg::LPR_M) * Cfg::STRIDE, srcM + (size_t)(uint32_t)snp_r * Cfg::ROWB_M);
      }
    } else {
#pragma unroll
      for (int i = 0; i < Cfg::LPR_S; ++i) {
        const int snp_r = __shfl_sync(0xffffffffu, snp, i * (32 / Cfg::LPR_S) + rsubS);
        if (pieceS < Cfg::NCH_S)
          ws_cp_async16(st_u32 + dstS + i * (32 / Cfg::LPR_S) * Cfg::STRIDE, srcS + (size_t)(uint32_t)snp_r * Cfg::ROWB_S);
      }
    }
    ws_cp_async_mbar_arrive_noinc(full0 + stage * 8);
    if (lane == 0) *reinterpret_cast<uint4*>(st + Cfg::OFF_HDR) = d;
    __syncwarp();
    if (lane == 0) ws_mbar_arrive(full0 + stage * 8);
    if (d.x & WS_FLAG_EXIT) break;
    prefetch(slot);  // batch t + WS_PF
  }
  ws_cp_async_wait<0>();
}

// ---- consumer --------------------------------------------------------------------------------------
template <int NV, int ROLE>
__device__ __forceinline__ void ws_consumer(const WsArgs& a, unsigned char* smem, const int ch, const int lane) {
  using Cfg = WsCfg<NV>;
  constexpr int NE = Cfg::NE, E_SG0 = Cfg::E_SG0, E_MX = Cfg::E_MX;
#define WS_OWN_ROW(j) ws_owns(NV, ROLE, (j))
  const uint32_t full0 = ws_smem_u32(smem + Cfg::OFF_FULL) + ch * WS_NSTAGE * 8;
  const uint32_t empty0 = ws_smem_u32(smem + Cfg::OFF_EMPTY) + ch * WS_NSTAGE * 8;
  const unsigned char* const stage0 = smem + Cfg::OFF_STAGE + (size_t)ch * WS_NSTAGE * Cfg::STAGE_B;
  const double* const tabM = reinterpret_cast<const double*>(smem + Cfg::OFF_TABM);
  const double4* const tabS = reinterpret_cast<const double4*>(smem + Cfg::OFF_TABS);
  double* const scr_m = reinterpret_cast<double*>(smem + Cfg::OFF_SCR + (size_t)(ch * 2 + ROLE) * ((Cfg::SCR_B + 15) & ~15));
  int* const scr_x = reinterpret_cast<int*>(scr_m + Cfg::NSLOT * 33);

  double acc[NE];
  int ex[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) { acc[e] = 1.0; ex[e] = 0; }
  int n_has = 0, since = 0;

  // header, this lane's record and its class-S table entry of the batch about to be processed; those
  // of batch t+1 are fetched while batch t is being multiplied (if the producer is ahead, as it should)
  uint4 hdr;
  uint2 rec;
  double4 tb;
  auto fetch_head = [&](int stage, uint4& h, uint2& r, double4& b) {
    const unsigned char* st = stage0 + (size_t)stage * Cfg::STAGE_B;
    h = *reinterpret_cast<const uint4*>(st + Cfg::OFF_HDR);  // {flags | n<<8, -, item, -}
    r = *reinterpret_cast<const uint2*>(st + Cfg::OFF_REC + lane * 8);
    b = tabS[r.y & 0xffu];                                   // {a + 1e-10*mx, b, mx, 2b}
  };
  ws_mbar_wait(full0, 0u);
  fetch_head(0, hdr, rec, tb);

#pragma unroll 1
  for (uint32_t t = 0;; ++t) {
    const int stage = t & (WS_NSTAGE - 1);
    const uint32_t flags = hdr.x & 0xffu, nvalid = hdr.x >> 8;
    if (flags & WS_FLAG_EXIT) break;
    const double2* const r2 = reinterpret_cast<const double2*>(stage0 + (size_t)stage * Cfg::STAGE_B + (size_t)lane * Cfg::STRIDE);
    const int stage1 = (t + 1) & (WS_NSTAGE - 1);
    const uint32_t parity1 = ((t + 1) / WS_NSTAGE) & 1u;
    uint4 hdr1 = hdr;
    uint2 rec1 = rec;
    double4 tb1 = tb;
    bool ready1;

    if (!(flags & WS_FLAG_M)) {
      // ---------------- class S: at most one usable base-call; h[i] = a1 + b*i ----------------
      double S[NV], M[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) { double2 v = r2[j]; S[j] = v.x; M[j] = v.y; }
      ready1 = __all_sync(0xffffffffu, ws_mbar_test(full0 + stage1 * 8, parity1));
      if (ready1) fetch_head(stage1, hdr1, rec1, tb1);
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
          if (WS_OWN_ROW(j)) {
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
      ready1 = __all_sync(0xffffffffu, ws_mbar_test(full0 + stage1 * 8, parity1));
      if (ready1) fetch_head(stage1, hdr1, rec1, tb1);
      if (((uint32_t)lane < nvalid) && G[0][0] != -1.0) {
        ++n_has;
        double f0, f1, f2, f3, f4;
        if (flags & WS_FLAG_D) {  // > 3 usable base-calls: folded at upload time (k_dmx_scatter)
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
        const double mx = ws_pmax(ws_pmax(ws_pmax(f0, f1), ws_pmax(f2, f3)), f4);
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
          if (WS_OWN_ROW(j)) {
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
    __syncwarp();  // every lane has consumed its row of this stage
    if (lane == 0) ws_mbar_arrive(empty0 + stage * 8);

    const bool item_end = (flags & WS_FLAG_END) != 0u;
    if (++since == 16 || item_end) {  // keep the running products inside the double range (16 terms >= ~1e-10 each)
      since = 0;
      pscl_renorm(acc[E_MX], ex[E_MX]);
      if (ROLE == 0) {
        pscl_renorm(acc[E_SG0], ex[E_SG0]);
#pragma unroll
        for (int j = 0; j < NV; ++j) pscl_renorm(acc[j], ex[j]);
      }
#pragma unroll
      for (int j = 1; j < NV; ++j) {
        if (WS_OWN_ROW(j)) {
#pragma unroll
          for (int k = 0; k < j; ++k) pscl_renorm(acc[NV + j * (j - 1) / 2 + k], ex[NV + j * (j - 1) / 2 + k]);
        }
      }
    }
    if (item_end) {
      // ---- item epilogue: the lanes' (mantissa in [1,2), exponent) pairs are transposed through
      // shared memory so that lane s multiplies the 32 factors of accumulator slot s and takes its ONE log
      scr_m[0 * 33 + lane] = acc[E_MX]; scr_x[0 * 33 + lane] = ex[E_MX];
      if (ROLE == 0) {
        scr_m[1 * 33 + lane] = acc[E_SG0]; scr_x[1 * 33 + lane] = ex[E_SG0];
#pragma unroll
        for (int j = 0; j < NV; ++j) { scr_m[(2 + j) * 33 + lane] = acc[j]; scr_x[(2 + j) * 33 + lane] = ex[j]; }
      }
#pragma unroll
      for (int j = 1; j < NV; ++j) {
        if (WS_OWN_ROW(j)) {
#pragma unroll
          for (int k = 0; k < j; ++k) {
            scr_m[ws_dbl_slot(NV, ROLE, j, k) * 33 + lane] = acc[NV + j * (j - 1) / 2 + k];
            scr_x[ws_dbl_slot(NV, ROLE, j, k) * 33 + lane] = ex[NV + j * (j - 1) / 2 + k];
          }
        }
      }
      int nh = n_has;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nh += __shfl_xor_sync(0xffffffffu, nh, o);
      __syncwarp();
      constexpr int NS = ws_nslot(NV, ROLE), DB = ws_dbl_base(NV, ROLE);
      double pm0 = 1.0, pm1 = 1.0;
      int px = 0;
      if (lane < NS) {
#pragma unroll 4
        for (int i = 0; i < 32; i += 2) {
          pm0 *= scr_m[lane * 33 + i];
          pm1 *= scr_m[lane * 33 + i + 1];
          px += scr_x[lane * 33 + i] + scr_x[lane * 33 + i + 1];
        }
        pm0 *= pm1;  // 32 mantissas in [1,2): < 2^32
        pscl_renorm(pm0, px);
      }
      const double lg = pscl_prod_log(pm0, px);
      // log(prod mx * (1+1e-10)^n_has): the pair normaliser of :704-725
      const double corr = __shfl_sync(0xffffffffu, lg, 0) + (double)nh * log1p(1e-10);
      const double sg0_log = (ROLE == 0) ? __shfl_sync(0xffffffffu, lg, 1) : 0.0;  // the k=0 column factor of :806
      double* out = a.partial + (size_t)((int)hdr.z - a.item_base) * (NV * NV * 2);
      if (ROLE == 0 && lane >= 2 && lane < 2 + NV) out[((lane - 2) * NV + 0) * 2 + 0] = lg - corr + sg0_log;
      if (lane >= DB && lane < NS) {
        int d = lane - DB, jj = 0, kk = 0;
#pragma unroll 1
        for (int j = 1; j < NV; ++j) {
          if (ws_owns(NV, ROLE, j)) {
            if (d < j) { jj = j; kk = d; break; }
            d -= j;
          }
        }
        const double x = lg - corr;
        out[(jj * NV + kk) * 2 + 1] = x;
        out[(kk * NV + jj) * 2 + 1] = x;
      }
      __syncwarp();
#pragma unroll
      for (int e = 0; e < NE; ++e) { acc[e] = 1.0; ex[e] = 0; }
      n_has = 0;
    }
    if (!ready1) {
      ws_mbar_wait(full0 + stage1 * 8, parity1);
      fetch_head(stage1, hdr1, rec1, tb1);
    }
    hdr = hdr1; rec = rec1; tb = tb1;
  }
#undef WS_OWN_ROW
}

template <int NV>
__global__ void __launch_bounds__(384, 1) k_demux_ws(WsArgs a) {
  using Cfg = WsCfg<NV>;
  extern __shared__ __align__(128) unsigned char ws_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < 4 * WS_NSTAGE; ++i) {
      ws_mbar_init(ws_smem_u32(ws_smem + Cfg::OFF_FULL) + i * 8, 33);  // 32 cp.async arrivals + the header's
      ws_mbar_init(ws_smem_u32(ws_smem + Cfg::OFF_EMPTY) + i * 8, 2);  // the two consumer warps
    }
  }
  {
    double* tabM = reinterpret_cast<double*>(ws_smem + Cfg::OFF_TABM);
    double* tabS = reinterpret_cast<double*>(ws_smem + Cfg::OFF_TABS);
    for (int i = tid; i < 3 * 64 * PSCL_FOLD_ROW; i += 384) tabM[i] = a.fold_tab[i];
    for (int c = tid; c < 256; c += 384) {
      const int cc = c < 3 * 64 ? c : 2 * 64;  // bytes beyond the table (class-D records) read the ones row
      const double t0 = a.fold_tab[cc * PSCL_FOLD_ROW], t4 = a.fold_tab[cc * PSCL_FOLD_ROW + 4];
      const double mx = ws_pmax(t0, t4), b = (t4 - t0) * 0.25;
      tabS[c * 4 + 0] = fma(1e-10, mx, t0);
      tabS[c * 4 + 1] = b;
      tabS[c * 4 + 2] = mx;
      tabS[c * 4 + 3] = b + b;
    }
  }
  __syncthreads();
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    ws_producer<NV>(a, ws_smem, warp, lane);
  } else if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    ws_consumer<NV, 0>(a, ws_smem, warp - 4, lane);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    ws_consumer<NV, 1>(a, ws_smem, warp - 8, lane);
  }
}

// ---- class-stream build (once per pileup image) ------------------------------------------------------
// record of every pair in original order + class key for the scan: low word counts class M
// (2-3 usable base-calls), high word class D (> 3)
__global__ void k_dmx_classify(const int32_t* __restrict__ pair_snp, const uint32_t* __restrict__ pair_rd,
                               const uint8_t* __restrict__ rd_aq, int64_t P, uint2* __restrict__ rec_tmp,
                               unsigned long long* __restrict__ key) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const uint32_t r0 = pair_rd[p], r1 = pair_rd[p + 1];
  uint32_t cnt = 0, codes = WS_NONE_CODES;
  for (uint32_t r = r0; r < r1; ++r) {
    const uint32_t aq = rd_aq[r];
    if ((aq >> 6) == 2u) continue;  // cmd_cram_demuxlet.cpp:664
    if (cnt < 3u) codes = (codes & ~(0xffu << (8 * cnt))) | (aq << (8 * cnt));
    ++cnt;
  }
  rec_tmp[p] = make_uint2((uint32_t)pair_snp[p], codes | (min(cnt, 255u) << 24));
  key[p] = cnt <= 1u ? 0ull : cnt <= 3u ? 1ull : (1ull << 32);
}

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

// record ranges of every work item, in natural item order and in the kernels' work order
__global__ void k_dmx_desc(const int64_t* __restrict__ cell_ptr, const int32_t* __restrict__ item_cell,
                           const int64_t* __restrict__ item_pbeg, const int64_t* __restrict__ item_pend,
                           const int32_t* __restrict__ item_order, int32_t n_items,
                           const unsigned long long* __restrict__ scan, uint4* __restrict__ desc_nat,
                           uint4* __restrict__ desc_sorted) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_items) return;
  auto make = [&](int item) {
    const int c = item_cell[item];
    const int64_t c0 = cell_ptr[c], c1 = cell_ptr[c + 1];
    const unsigned long long s0 = scan[c0], s1 = scan[c1];
    const int64_t n_m = (uint32_t)(s1 - s0), n_d = (uint32_t)((s1 >> 32) - (s0 >> 32)), n_s = (c1 - c0) - n_m - n_d;
    return make_uint4((uint32_t)item_pbeg[item], (uint32_t)item_pend[item], (uint32_t)(c0 + n_s), (uint32_t)(c0 + n_s + n_m));
  };
  desc_nat[w] = make(w);
  desc_sorted[w] = make(item_order[w]);
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
  if (p->dmx_rec) return PSCL_OK;
  const int64_t P = p->P;
  if (P + 1 > INT32_MAX) return pscl_fail(ctx, PSCL_EINVAL, "k_demux_ws: a device pileup image holds < 2^31 pairs");
  uint2* rec_tmp = nullptr;
  unsigned long long *key = nullptr, *scan = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** d, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(d, bytes ? bytes : 16); };
  alloc((void**)&rec_tmp, sizeof(uint2) * P);
  alloc((void**)&key, sizeof(unsigned long long) * (P + 1));
  alloc((void**)&scan, sizeof(unsigned long long) * (P + 2));
  alloc((void**)&p->dmx_rec, sizeof(uint2) * P);
  alloc((void**)&p->dmx_desc_nat, sizeof(uint4) * p->n_items);
  alloc((void**)&p->dmx_desc_sorted, sizeof(uint4) * p->n_items);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, key, scan, (int)(P + 1), ctx->stream);
  alloc(&tmp, tmp_bytes);
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
  unsigned long long totals = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&totals, scan + P, sizeof(totals), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  const size_t n_deep = (size_t)(totals >> 32);
  alloc((void**)&p->dmx_deep, sizeof(double) * 6 * n_deep);
  if (e == cudaSuccess && p->n_items > 0) {
    k_dmx_scatter<<<(unsigned)(((int64_t)p->n_items * 32 + 255) / 256), 256, 0, ctx->stream>>>(
        p->cell_ptr, p->item_cell, p->item_pbeg, p->item_pend, p->n_items, rec_tmp, scan, p->pair_rd, p->rd_aq, ctx->fold_tab,
        p->dmx_rec, p->dmx_deep);
    k_dmx_desc<<<(unsigned)((p->n_items + 255) / 256), 256, 0, ctx->stream>>>(
        p->cell_ptr, p->item_cell, p->item_pbeg, p->item_pend, p->item_order, p->n_items, scan, p->dmx_desc_nat, p->dmx_desc_sorted);
    ctx->launches += 2;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(rec_tmp); cudaFree(key); cudaFree(scan); cudaFree(tmp);
  if (e != cudaSuccess) {
    cudaFree(p->dmx_rec); cudaFree(p->dmx_deep); cudaFree(p->dmx_desc_nat); cudaFree(p->dmx_desc_sorted);
    p->dmx_rec = nullptr; p->dmx_deep = nullptr; p->dmx_desc_nat = nullptr; p->dmx_desc_sorted = nullptr;
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
static cudaError_t launch_ws(pscl_ctx* ctx, const WsArgs& a) {
  using Cfg = WsCfg<NV>;
  static bool attr_set[64] = {false};
  if (!attr_set[ctx->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_demux_ws<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
    attr_set[ctx->device & 63] = true;
  }
  int grid = std::min(ctx->sm_count, (a.n_work + 3) / 4);
  if (grid < 1) grid = 1;
  k_demux_ws<NV><<<grid, 384, Cfg::SMEM, ctx->stream>>>(a);
  return cudaGetLastError();
}
