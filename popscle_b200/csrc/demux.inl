// demux.inl — demuxlet per-barcode likelihood grid + per-cell epilogue (part of popscle_b200.cu)
//
// Replaces cmd_cram_demuxlet.cpp:636-991.  Two accumulation kernels write per-work-item partial
// LLK grids laid out like the reference's llksAB ([j][k][n], :620/:746); one epilogue kernel sums
// the items of a cell in fixed order and restates the prior / logAdd / best-next / SNG-DBL-AMB
// logic (:788-991).
//
//   k_demux_default<NV, DELTA, DICT> : the default alpha grid {0, 0.5} (:85-89) and 2 <= nv <= 8.  One warp
//       per work item, ONE LANE PER (cell,SNP) PAIR: the nv singlet + nv(nv-1)/2 doublet running
//       products live in that lane's registers.  Inputs are software-pipelined: indices two
//       iterations ahead (registers), read bytes and the genotype row one iteration ahead
//       (cp.async into a conflict-free shared-memory row per lane).
//       DICT : the genotype table is made of <= 256 distinct triples (hard calls): 8 bytes of codes per
//              pair instead of the row, the triples in 16 conflict-free shared-memory copies (bit-identical).
//       DELTA: the staged pscl_demux_run: SNP ids come from 16-bit gaps that are still crossing PCIe in
//              slices when the kernel starts; a warp waits on its slice's flag and scans the gaps itself.
//   k_demux_poly (demux_poly.inl), k_demux_cls / k_demux_ab (demux_cls.inl, demux_ab.inl): see those files.
//   k_demux_general<EPT> : any nv / alpha grid.  One CTA per (work item, tile of 256*EPT grid
//       entries); 8 pairs per round are folded by one warp each, staged in shared memory, and
//       every thread updates its EPT register accumulators.
//
// Arithmetic notes (all FP64):
//  * pG (:655-725) is evaluated in closed form: the reference divides by the running joint
//    maximum after every read and once more after adding 1e-10, which equals
//    (raw/max(raw) + 1e-10)/(1 + 1e-10) up to rounding.
//  * log(sum) per term (:746) becomes a running product with a separately tracked exponent and a
//    single log per accumulator per work item.
//  * sum_lm g_j[l] g_k[m] pG[n][l][m] is evaluated as g_j . (pG[n] g_k).
//  * at alpha == 0.5, pG[l][m] depends on l+m only, so LLK[j][k] == LLK[k][j] mathematically; the
//    default kernel computes k<j once and mirrors it (the reference's choice between (j,k) and
//    (k,j) is rounding noise; SURVEY.md Appendix C compares that pair unordered).

#include <cuda_pipeline_primitives.h>

// ------------------------------------------------------------------------------------------------
// default-grid kernel
// ------------------------------------------------------------------------------------------------
// Per-read factors of the 5 distinct mixing fractions p = 0, .25, .5, .75, 1 that the {0, 0.5} grid
// produces (p = 0.5*l + (m-l)*0.5*alpha, :673): fold_tab[allele][qual][i] = pR*(1-p_i) + pA*p_i with
// pR/pA from :666-667.  Row allele==2 (skipped read, :664) and the "no read" row are all ones, so
// the fold is branch-free.  Filled on the host in the reference's own expression order.
#define PSCL_FOLD_ROW 6  /* 5 factors padded to 48 B */
#define PSCL_FOLD_ONES (2 * 64)
#define PSCL_DICT_N 256 /* entries of the genotype dictionary (8-bit codes) */
#define PSCL_CLS_SLOTS 12 /* genotype-class variant: 9 doublet + 3 singlet products per pair */
#ifndef PSCL_RENORM_EVERY
#define PSCL_RENORM_EVERY 16 /* pairs between two exponent extractions of the running products (power of two) */
#endif

// max of positive, non-NaN doubles: DSETP + 2 SEL instead of fmax()'s NaN-aware sequence
__device__ __forceinline__ double dmx_pmax(double x, double y) { return x > y ? x : y; }

template <int NV>
struct DefaultCfg {
  static constexpr int ND = NV * (NV - 1) / 2;            // doublet accumulators (k < j)
  static constexpr int NE = NV + ND + 2;                  // + singlet column scale + pair normaliser
  static constexpr int ROW_D = NV * 3;                    // doubles per genotype row
  static constexpr bool V16 = (NV % 2 == 0);              // row is a multiple of 16 B
  // smem row stride in doubles: odd multiple of 16 B (even NV) / odd number of doubles (odd NV)
  static constexpr int STRIDE_D = V16 ? ((ROW_D / 2) | 1) * 2 : ROW_D;
  // cooperative gather geometry: CH-byte pieces, LPR (power of two) lanes per row, RPI rows per instruction
  static constexpr int CH = V16 ? 16 : 8;
  static constexpr int NCH = ROW_D * 8 / CH;
  static constexpr int LPR = NCH <= 1 ? 1 : NCH <= 2 ? 2 : NCH <= 4 ? 4 : NCH <= 8 ? 8 : NCH <= 16 ? 16 : 32;
  static constexpr int RPI = 32 / LPR;
  static_assert(NCH <= 32, "genotype row too long for the cooperative gather");
  static_assert(NE <= 40, "the item epilogue reduces at most 32 + 8 accumulators");
  // fold table | [2][threads] genotype rows | [NE][threads] exponents
  __host__ __device__ static constexpr size_t smem_rows(int nt) {
    return sizeof(double) * 3 * 64 * PSCL_FOLD_ROW + (size_t)2 * nt * STRIDE_D * sizeof(double) + (size_t)NE * nt * sizeof(int);
  }
  // dictionary variants: fold table | [NE][threads] exponents | 256 triples x 16 copies
  __host__ __device__ static constexpr size_t smem_dict(int nt) {
    return sizeof(double) * 3 * 64 * PSCL_FOLD_ROW + (size_t)NE * nt * sizeof(int) + (size_t)PSCL_DICT_N * 3 * sizeof(double) * 16;
  }
  // genotype-class variant: the same + [12][threads] per-pair products of the SNP's (at most three) distinct triples
  __host__ __device__ static constexpr size_t smem_cls(int nt) { return smem_dict(nt) + (size_t)PSCL_CLS_SLOTS * nt * sizeof(double); }
};

struct DemuxArgs {
  const int32_t* pair_snp;
  const uint32_t* pair_rd;
  const uint8_t* rd_aq;
  const double* gp;
  const uint8_t* has_gp;
  const double* phred_err;
  const double* fold_tab;     // [3][64][PSCL_FOLD_ROW]
  const int32_t* item_order;  // nullable
  const int64_t* item_pbeg;
  const int64_t* item_pend;
  double* partial;            // [n_work][nv*nv*nalpha]
  int* counter;
  int32_t item_base, n_work;
  int32_t nv, nalpha;
  // staged run (k_demux_default<NV, true>): SNP ids come from the ABI-3 gaps, which are still crossing PCIe in slices of
  // whole cells when the kernel starts; flags[k] is written behind slice k by the same copy queue
  const uint16_t* delta = nullptr;     // [P] gap to the previous pair's SNP id (ignored at a cell's first pair)
  const uint8_t* delta8 = nullptr;     // ABI 6 (DELTA == 2): the same in 8 bits, 255 = the next entry of gap_big
  const uint32_t* gap_big = nullptr;
  const int64_t* cell_gap_ptr = nullptr;  // [C+1] first large gap of every cell
  int64_t n_gap_big = 0;
  const int32_t* first = nullptr;      // [C] SNP id of each cell's first pair
  const int64_t* cell_ptr = nullptr;   // [C+1]
  const int32_t* item_cell = nullptr;  // [n_items]
  const int* flags = nullptr;          // [n_stages]
  int* bad = nullptr;                  // 2: SNP id out of range, 4: a slice never arrived
  int32_t n_snps = 0, n_stages = 0;
  long long spin_limit = 0;            // clock64 ticks a warp waits for a slice before it gives up (bad = 4)
  int32_t stage_cell[PSCL_MAX_STAGES + 1] = {0};
  // dictionary-coded genotypes (k_demux_default<NV, *, true>): when every (SNP, sample) triple of the table is one of
  // <= 256 distinct triples (hard calls: 3 per combination of genotype counts), a pair needs 8 bytes of codes instead of
  // a 24*nv-byte row, and the triples themselves sit in shared memory
  const unsigned long long* gp_code = nullptr;  // [V] 8-bit dictionary index per sample
  const double* gp_dict = nullptr;              // [256][3]
  // genotype classes (k_demux_default<NV, 0, 2>): a SNP of a hard-call table holds at most three distinct triples (one per
  // genotype: (1-err) onehot + err avg with avg common to the SNP's samples), so gp_code then points at, per SNP, the
  // dictionary indices of those triples in order of first appearance (bytes 0-2) and each sample's class, 2 bits from bit 32
  int32_t cls = 0;
};

// NT threads per CTA, one CTA per SM.  (Splitting a work item's running products between two warps, so that 12 or 16
// warps fit an SM at 158 / 128 registers, was measured at 0.67 / 0.71 ms against 0.53 ms for this form: the duplicated
// fold and the extra work items cost more than the added warps hide; profiles/r2a_variants.txt.)
template <int NV, int DELTA, int DICT, int NT>
__global__ void __launch_bounds__(NT, 1) k_demux_default(DemuxArgs a) {
  using Cfg = DefaultCfg<NV>;
  constexpr int ND = Cfg::ND, SD = Cfg::STRIDE_D, NAM = Cfg::NE;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_tab = reinterpret_cast<double*>(smem_raw);        // [3*64][PSCL_FOLD_ROW]
  double* s_g = s_tab + 3 * 64 * PSCL_FOLD_ROW;               // rows: [2][NT][SD] genotype rows
  int* s_exp = reinterpret_cast<int*>(s_g + (DICT ? 0 : 2 * NT * SD));  // [NAM][NT]
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 3 * 64 * PSCL_FOLD_ROW; i += NT) s_tab[i] = a.fold_tab[i];
  // DICT: the dictionary, 16 copies interleaved ([256*3][16] doubles): lane l reads copy l%16, which lives in its own
  // pair of banks, so the 32 lanes' lookups of 32 different triples never collide (a single copy cost ~3x the wavefronts)
  double* const s_dict = reinterpret_cast<double*>(s_exp + NAM * NT);
  if constexpr (DICT != 0) {
    for (int i = tid; i < PSCL_DICT_N * 3 * 16; i += NT) s_dict[i] = a.gp_dict[i >> 4];
  }
  double* const s_cls = s_dict + PSCL_DICT_N * 3 * 16;        // DICT == 2: [PSCL_CLS_SLOTS][NT], column tid is this lane's
  __syncthreads();
  double* const g_row0 = s_g + (size_t)tid * SD;              // buffer 1 is NT*SD doubles further

  // the next work index is fetched while the current item is processed (inline PTX: the compiler would
  // otherwise turn the lane-0 atomicAdd into a warp-aggregated one whose result is needed at once)
  auto grab = [&]() { int w = 0; if (lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(w) : "l"(a.counter) : "memory"); return w; };

  // ---- one work item ------------------------------------------------------------------------------------------------
  // accumulators: [0, NV) singlets | [NV, NV+ND) doublets (j,k<j) at NV + j(j-1)/2 + k | k=0 column factor | pair normaliser
  unsigned seen_slices = 0;  // DELTA: slices of the staged image whose flag this warp has already seen set
  auto run = [&](const int item) {
    constexpr int NACC = Cfg::NE, E_SG0 = NV + ND, E_MX = NV + ND + 1;

    const int64_t pb = a.item_pbeg[item], pe = a.item_pend[item];
    const int niter = (int)((pe - pb + 31) >> 5);
    int snp_run = 0;       // DELTA: SNP id of the pair before the next 32
    int64_t cell_pb = -1;  // DELTA: first pair of the item's cell
    int64_t big_run = 0;   // DELTA == 2: next unread entry of gap_big
    if constexpr (DELTA != 0) {
      const int c = a.item_cell[item];
      cell_pb = a.cell_ptr[c];
      int k = 0;
      for (int i = 1; i < a.n_stages; ++i) k += (c >= a.stage_cell[i]) ? 1 : 0;
      // Wait for the slice (bounded: a copy that never lands must not hang the device).  The flag is polled with a relaxed
      // system-scope load, once per slice and warp (seen_slices).  What must hold is that the gaps are read after the flag
      // was seen set: they are read with ld.global.cg (L2, the point of coherence the copy engine writes to; nothing of the
      // slice can sit in this SM's L1 before), by loads issued after the branch on the flag's value has resolved.
      // (Timing note: under ncu this kernel reports ~0.87 ms on configs[1] against 0.64 ms with the wait compiled out — the
      // difference is the kernel really waiting for the last slices to cross PCIe, not a cost of the polling.)
      if (!((seen_slices >> k) & 1u)) {
        if (lane == 0) {
          const long long t_start = clock64();
          for (;;) {
            int f;
            asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(f) : "l"(a.flags + k) : "memory");
            if (f) break;
            __nanosleep(200);
            if (clock64() - t_start > a.spin_limit) { atomicExch(a.bad, 4); break; }
          }
        }
        __syncwarp();
        seen_slices |= 1u << k;
      }
      snp_run = a.first[c];
      if constexpr (DELTA == 2) big_run = a.cell_gap_ptr[c];
      if (pb > cell_pb) {  // a later work item of a large cell: id of the pair before it
        int sum = 0;
        if constexpr (DELTA == 2) {  // 8-bit gaps: the markers before pb own the next entries of gap_big, in order
          int nmk = 0;
          for (int64_t q = cell_pb + 1 + lane; q < pb; q += 32) { const int d = (int)__ldcg(a.delta8 + q); if (d == 255) ++nmk; else sum += d; }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) nmk += __shfl_xor_sync(0xffffffffu, nmk, o);
          for (int k = lane; k < nmk; k += 32) sum += (big_run + k < a.n_gap_big) ? (int)a.gap_big[big_run + k] : 0;
          big_run += nmk;
        } else {
          for (int64_t q = cell_pb + 1 + lane; q < pb; q += 32) sum += (int)__ldcg(a.delta + q);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        snp_run += sum;
      }
    }

    double acc[NACC];
#pragma unroll
    for (int e = 0; e < NACC; ++e) { acc[e] = 1.0; s_exp[e * NT + tid] = 0; }
    int n_has = 0;

    // ---- software pipeline --------------------------------------------------------------------
    // stage A (two iterations ahead): snp, r0, r1
    // stage B (one ahead): first 3 read bytes (kept in separate registers so nothing waits on them
    //                      before use), has_gp, genotype row via cp.async into this lane's smem row
    int32_t snpA = 0; uint32_t r0A = 0, r1A = 0; bool okA = false;
    uint32_t r0B = 0, r1B = 0, b0B = 0, b1B = 0, b2B = 0, hasB = 0;
    unsigned long long codeB = 0;
    // DELTA: two more stages in front, one loadA call apart each, so that no step waits on a load issued in the same call:
    //   raw    the 8- or 16-bit gap of iteration it+2 is loaded;
    //   marker (8-bit form) the raw gaps of iteration it+1 are tested for the marker 255 and the marked lanes load their
    //          large gap (ballot/popc rank into gap_big);
    //   scan   the resolved gaps of iteration it become SNP ids (warp scan on top of the running id).
    int rawR = 0, gapM = 0;
    auto raw_gap = [&](int k) {  // gap byte / halfword of iteration k's pair (0 beyond the item and at a cell's first pair)
      const int64_t p = pb + ((int64_t)k << 5) + lane;
      if (k >= niter || p >= pe || p == cell_pb) return 0;
      if constexpr (DELTA == 1) return (int)__ldcg(a.delta + p);
      else return (int)__ldcg(a.delta8 + p);
    };
    auto resolve = [&](int raw) {  // marker -> large gap (the load it issues is consumed by the NEXT call's scan)
      int d = raw;
      if constexpr (DELTA == 2) {
        const bool mk = d == 255;
        const unsigned m = __ballot_sync(0xffffffffu, mk);
        if (mk) {
          const int64_t k = big_run + __popc(m & ((1u << lane) - 1u));
          if (k < a.n_gap_big) d = (int32_t)a.gap_big[k];
          else { d = 0; atomicExch(a.bad, 2); }  // more markers than large gaps: malformed input
        }
        big_run += __popc(m);
      }
      return d;
    };
    auto loadA = [&](int it) {
      int64_t p = pb + ((int64_t)it << 5) + lane;
      okA = (it < niter) && (p < pe);
      if (okA) {
        if constexpr (DELTA == 0) snpA = a.pair_snp[p];
        r0A = a.pair_rd[p]; r1A = a.pair_rd[p + 1];
      }
      if constexpr (DELTA != 0) {
        int d = gapM;  // resolved gap of this iteration's pair
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, d, o); if (lane >= o) d += t; }
        d += snp_run;
        snp_run = __shfl_sync(0xffffffffu, d, 31);
        snpA = d;
        if (okA && (unsigned)d >= (unsigned)a.n_snps) { okA = false; atomicExch(a.bad, 2); }
        gapM = resolve(rawR);      // iteration it+1
        rawR = raw_gap(it + 2);    // iteration it+2
      }
    };
    auto issueB = [&](int buf) {  // consumes stage A
      r0B = r0A; r1B = r1A; hasB = 0; b0B = b1B = b2B = PSCL_FOLD_ONES << 6;  // placeholder: "no read"
      if (okA) {
        hasB = 1;
        if (a.has_gp) hasB = a.has_gp[snpA];
        const uint32_t n = r1B - r0B;
        if (n > 0) b0B = a.rd_aq[r0B];
        if (n > 1) b1B = a.rd_aq[r0B + 1];
        if (n > 2) b2B = a.rd_aq[r0B + 2];
        if constexpr (DICT != 0) codeB = a.gp_code[snpA];
      }
      if constexpr (DICT != 0) return;  // no genotype rows to fetch
      // Cooperative row gather: the warp's 32 genotype rows are fetched LPR lanes per row, so one
      // LDGSTS instruction touches 32/LPR whole rows (a few 128-B lines) instead of 32 scattered
      // 16-B pieces of 32 different rows (one L1 tag lookup each).  A denser mapping (piece
      // f = 32 i + lane: 12 instead of 16 instructions per 192-byte-row batch) was measured and is
      // NOT faster (0.749 vs 0.740 ms): the cost follows the rows touched, not the instructions.
      const unsigned okmask = __ballot_sync(0xffffffffu, okA);
      double* dst_base = s_g + ((size_t)buf * NT + (tid & ~31)) * SD;
#pragma unroll
      for (int i = 0; i < Cfg::LPR; ++i) {
        const int row = i * Cfg::RPI + (lane / Cfg::LPR), piece = lane % Cfg::LPR;
        const int snp_r = __shfl_sync(0xffffffffu, snpA, row);
        if (((okmask >> row) & 1u) && piece < Cfg::NCH) {
          const char* src = reinterpret_cast<const char*>(a.gp + (size_t)snp_r * Cfg::ROW_D) + piece * Cfg::CH;
          char* dst = reinterpret_cast<char*>(dst_base + (size_t)row * SD) + piece * Cfg::CH;
          __pipeline_memcpy_async(dst, src, Cfg::CH);
        }
      }
      __pipeline_commit();
    };
    if constexpr (DELTA != 0) {  // prime the two front stages (two exposed loads per item)
      gapM = resolve(raw_gap(0));
      rawR = raw_gap(1);
    }
    loadA(0);
    issueB(0);
    loadA(1);

    for (int it = 0; it < niter; ++it) {
      const int buf = it & 1;
      const uint32_t r0 = r0B, r1 = r1B, b0 = b0B, b1 = b1B, b2 = b2B;
      const unsigned long long code = codeB;
      const bool has = hasB != 0;
      if constexpr (DICT == 0) __syncwarp();  // every lane is done reading buffer buf^1 (iteration it-1) before it is refilled
      issueB(buf ^ 1);
      loadA(it + 2);
      if constexpr (DICT == 0) {
        __pipeline_wait_prior(1);  // this lane's pieces of iteration `it` have landed ...
        __syncwarp();              // ... and so have the other lanes' pieces of this lane's row
      }

      if (has) {
        ++n_has;
        // ---- D1: fold the reads (cmd_cram_demuxlet.cpp:660-700), branch-free for the first 3 ----
        // byte = allele<<6 | qual -> table row allele*64+qual; the placeholder maps to the ones row
        const double* t0 = s_tab + (b0 < 256u ? b0 : (uint32_t)PSCL_FOLD_ONES) * PSCL_FOLD_ROW;
        const double* t1 = s_tab + (b1 < 256u ? b1 : (uint32_t)PSCL_FOLD_ONES) * PSCL_FOLD_ROW;
        const double* t2 = s_tab + (b2 < 256u ? b2 : (uint32_t)PSCL_FOLD_ONES) * PSCL_FOLD_ROW;
        double f0 = t0[0] * t1[0] * t2[0], f1 = t0[1] * t1[1] * t2[1], f2 = t0[2] * t1[2] * t2[2],
               f3 = t0[3] * t1[3] * t2[3], f4 = t0[4] * t1[4] * t2[4];
        const uint32_t nrd = r1 - r0;
        for (uint32_t r = 3; r < nrd; ++r) {  // rare: > 3 base-calls on one (cell,SNP)
          const double* t = s_tab + (uint32_t)a.rd_aq[r0 + r] * PSCL_FOLD_ROW;
          f0 *= t[0]; f1 *= t[1]; f2 *= t[2]; f3 *= t[3]; f4 *= t[4];
          if ((r & 7u) == 7u) {  // deep pileups: rescale by the running max like :692-699
            const double ri = 1.0 / dmx_pmax(dmx_pmax(dmx_pmax(f0, f1), dmx_pmax(f2, f3)), f4);
            f0 *= ri; f1 *= ri; f2 *= ri; f3 *= ri; f4 *= ri;
          }
        }
        // ---- D2 (:704-725) without the division: pG = (f/mx + 1e-10)/(1+1e-10) = h/(mx*(1+1e-10))
        // with h = f + 1e-10*mx; the common factor mx is accumulated once per pair (acc[E_MX]) and
        // (1+1e-10)^n_has is applied at the end.
        const double mx = dmx_pmax(dmx_pmax(dmx_pmax(f0, f1), dmx_pmax(f2, f3)), f4);
        const double h0 = fma(1e-10, mx, f0), h1 = fma(1e-10, mx, f1), h2 = fma(1e-10, mx, f2),
                     h3 = fma(1e-10, mx, f3), h4 = fma(1e-10, mx, f4);
        acc[E_MX] *= mx;

        // ---- D3: genotype row from this lane's shared-memory row ------------------------------
        if constexpr (DICT == 2) {
          // Genotype classes: the SNP's samples share at most three triples T[c], so the pair's singlet factor takes three
          // values S[c] = T[c].(h0,h2,h4) and its doublet factor nine, D[cj][ck] = T[ck].(H T[cj]) — 63 FP64 operations
          // instead of 171, the same expressions on the same doubles as below (bit-identical records).  They go to this
          // lane's column of s_cls and every accumulator picks its factor by its samples' classes: one add and one 64-bit
          // shared-memory load per accumulator in place of three FP64 operations.
          double T[3][3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const double* d = s_dict + (uint32_t)((code >> (8 * c)) & 255ull) * 48 + (lane & 15);
            T[c][0] = d[0]; T[c][1] = d[16]; T[c][2] = d[32];
          }
          acc[E_SG0] *= (T[0][0] + T[0][1] + T[0][2]);  // sample 0 opens class 0
          double* const sc = s_cls + tid;
          double S0 = 0.0;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const double S = (T[c][0] * h0 + T[c][1] * h2 + T[c][2] * h4);
            if (c == 0) S0 = S;
            sc[(9 + c) * NT] = S;
          }
#pragma unroll
          for (int cj = 0; cj < 3; ++cj) {
            const double v0 = h0 * T[cj][0] + h1 * T[cj][1] + h2 * T[cj][2];
            const double v1 = h1 * T[cj][0] + h2 * T[cj][1] + h3 * T[cj][2];
            const double v2 = h2 * T[cj][0] + h3 * T[cj][1] + h4 * T[cj][2];
#pragma unroll
            for (int ck = 0; ck < 3; ++ck) sc[(cj * 3 + ck) * NT] = (T[ck][0] * v0 + T[ck][1] * v1 + T[ck][2] * v2);
          }
          const uint32_t cb = (uint32_t)(code >> 32);
          uint32_t off[NV];  // byte offset of sample j's class between two slots of s_cls
#pragma unroll
          for (int j = 0; j < NV; ++j) off[j] = ((cb >> (2 * j)) & 3u) * (uint32_t)(NT * sizeof(double));
          const char* const scb = reinterpret_cast<const char*>(sc);
          acc[0] *= S0;
#pragma unroll
          for (int j = 1; j < NV; ++j) acc[j] *= *reinterpret_cast<const double*>(scb + 9 * NT * sizeof(double) + off[j]);
#pragma unroll
          for (int j = 1; j < NV; ++j) {
            const char* const rowb = scb + 3 * off[j];
#pragma unroll
            for (int k = 0; k < j; ++k) acc[NV + j * (j - 1) / 2 + k] *= *reinterpret_cast<const double*>(rowb + off[k]);
          }
        } else {
        double G[NV][3];
        if constexpr (DICT != 0) {
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            const double* d = s_dict + (uint32_t)((code >> (8 * j)) & 255ull) * 48 + (lane & 15);
            G[j][0] = d[0]; G[j][1] = d[16]; G[j][2] = d[32];
          }
        } else {
          const double* row = g_row0 + (size_t)buf * NT * SD;
          if constexpr (Cfg::V16) {
            const double2* r2 = reinterpret_cast<const double2*>(row);
            double flat[Cfg::ROW_D];
#pragma unroll
            for (int i = 0; i < Cfg::ROW_D / 2; ++i) { double2 t = r2[i]; flat[2 * i] = t.x; flat[2 * i + 1] = t.y; }
#pragma unroll
            for (int j = 0; j < NV; ++j) { G[j][0] = flat[3 * j]; G[j][1] = flat[3 * j + 1]; G[j][2] = flat[3 * j + 2]; }
          } else {
#pragma unroll
            for (int j = 0; j < NV; ++j) { G[j][0] = row[3 * j]; G[j][1] = row[3 * j + 1]; G[j][2] = row[3 * j + 2]; }
          }
        }
        // singlets: llksAB[j][0][0] (:806) = log((sum_l g_j[l] pG0[l]) * (sum_m g_0[m])), pG0[l] = h(2l);
        // the (sum_m g_0[m]) column factor is common to all j and accumulated once
        acc[E_SG0] *= (G[0][0] + G[0][1] + G[0][2]);
#pragma unroll
        for (int j = 0; j < NV; ++j) acc[j] *= (G[j][0] * h0 + G[j][1] * h2 + G[j][2] * h4);
        // doublets at alpha = 0.5: pG1[l][m] = h(l+m)
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
      // Exponents move to shared memory every PSCL_RENORM_EVERY pairs.  A term is >= min h >= 1e-10 * mx (rows sum to one)
      // and mx >= f(p = 0.5) >= 0.25^3 for the three folded base-calls (deeper pairs are rescaled), i.e. >= 1.5e-12:
      // 16 of them stay above 1e-190, far inside a double's exponent range.
      if ((it & (PSCL_RENORM_EVERY - 1)) == PSCL_RENORM_EVERY - 1) {
#pragma unroll
        for (int e = 0; e < NACC; ++e) { int ex = 0; pscl_renorm(acc[e], ex); s_exp[e * NT + tid] += ex; }
      }
    }
    if constexpr (DICT == 0) __pipeline_wait_prior(0);

    // ---- item epilogue ---------------------------------------------------------------------------
    // Lane products are multiplied across the warp by a transpose-reduction (recursive halving: 31 + 7 exchanged
    // values per lane instead of 5 butterfly steps per accumulator; mantissas are in [1,2) after renorm, so 32 of
    // them cannot overflow), then ONE log per accumulator per item, taken by the lane that ends up holding it.
    // Two groups: GA in {8, 16, 32} accumulators, then GB in {0, 8} (NACC <= 40); after a transpose of N the product
    // of element e sits in lane e * 32 / N.
    constexpr int GA = NACC <= 8 ? 8 : NACC <= 24 ? 16 : 32, GB = NACC > GA ? 8 : 0;
    static_assert(NACC <= GA + GB, "epilogue groups");
    double lg1, lg2 = 0.0;
    {
      double m1[GA]; int x1[GA];
#pragma unroll
      for (int e = 0; e < GA; ++e) {
        if (e < NACC) { m1[e] = acc[e]; x1[e] = s_exp[e * NT + tid]; pscl_renorm(m1[e], x1[e]); }
        else { m1[e] = 1.0; x1[e] = 0; }
      }
      pscl_transpose_prod<GA>(m1, x1, lane);
      pscl_renorm(m1[0], x1[0]);
      lg1 = pscl_prod_log(m1[0], x1[0]);
    }
    if constexpr (GB > 0) {
      double m2[8]; int x2[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (GA + e < NACC) { m2[e] = acc[GA + e]; x2[e] = s_exp[(GA + e) * NT + tid]; pscl_renorm(m2[e], x2[e]); }
        else { m2[e] = 1.0; x2[e] = 0; }
      }
      pscl_transpose_prod<8>(m2, x2, lane);
      pscl_renorm(m2[0], x2[0]);
      lg2 = pscl_prod_log(m2[0], x2[0]);
    }
    auto pick = [&](int e) { return e < GA ? __shfl_sync(0xffffffffu, lg1, e * (32 / GA)) : __shfl_sync(0xffffffffu, lg2, (e - GA) * 4); };
    int nh = n_has;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nh += __shfl_xor_sync(0xffffffffu, nh, o);
    // log(prod mx * (1+1e-10)^n_has): the pair normaliser of :704-725
    const double corr = pick(E_MX) + (double)nh * log1p(1e-10);
    const double sg0_log = pick(E_SG0);  // log prod (sum_m g_0[m]), the k=0 column factor of :806
    double* out = a.partial + (size_t)(item - a.item_base) * (NV * NV * 2);
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
    if (lane % (32 / GA) == 0) store(lane / (32 / GA), lg1);
    if (GB > 0 && (lane & 3) == 0) store(GA + (lane >> 2), lg2);
    __syncwarp();
  };

  int w_next = grab();
  for (;;) {
    const int w = __shfl_sync(0xffffffffu, w_next, 0);
    if (w >= a.n_work) break;
    w_next = grab();
    run(a.item_order ? a.item_order[w] : a.item_base + w);
  }
}

// ------------------------------------------------------------------------------------------------
// general kernel
// ------------------------------------------------------------------------------------------------
struct GeneralArgs {
  DemuxArgs d;
  int32_t E;        // nv + (nalpha-1)*nv*nv entries
  int32_t PB;       // pair slots per round (<= 8)
  int32_t slot_d;   // doubles per slot
  int32_t nvv_max;  // v doubles per slot
};

template <int EPT>
__global__ void __launch_bounds__(256) k_demux_general(GeneralArgs ga) {
  const DemuxArgs& a = ga.d;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_err = reinterpret_cast<double*>(smem_raw);  // [256]
  double* s_slots = s_err + 256;                        // [PB][slot_d] : pg | g | v
  int* s_has = reinterpret_cast<int*>(s_slots + (size_t)ga.PB * ga.slot_d);  // [PB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nv = a.nv, na = a.nalpha, nv2 = nv * nv;
  s_err[tid] = a.phred_err[tid];

  const int w = blockIdx.x;
  const int item = a.item_order ? a.item_order[w] : a.item_base + w;
  const int64_t pb = a.item_pbeg[item], pe = a.item_pend[item];
  const int e_lo = blockIdx.y * (256 * EPT);
  const int e_hi = min(ga.E, e_lo + 256 * EPT) - 1;
  auto plane_of = [&](int e) { return e < nv ? 0 : 1 + (e - nv) / nv2; };
  const int nlo = plane_of(e_lo), nhi = plane_of(e_hi);
  const int nvv = (nhi - nlo + 1) * nv * 3;
  const int pg_d = na * 9, g_d = nv * 3;

  double acc[EPT];
  int ex[EPT], goff[EPT], voff[EPT], gidx[EPT];
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    int e = e_lo + i * 256 + tid;
    acc[i] = 1.0; ex[i] = 0; gidx[i] = -1; goff[i] = 0; voff[i] = 0;
    if (e <= e_hi) {
      int n, j, k;
      if (e < nv) { n = 0; j = e; k = 0; }
      else { int d = e - nv; n = 1 + d / nv2; j = (d / nv) % nv; k = d % nv; }
      goff[i] = j * 3;
      voff[i] = ((n - nlo) * nv + k) * 3;
      gidx[i] = (j * nv + k) * na + n;
    }
  }
  // per-lane alpha weights for the fold: value i = lane + 32*t  ->  (n,l,m)
  constexpr int VPL = (PSCL_MAX_ALPHA * 9 + 31) / 32;
  double pw[VPL];
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    int i = lane + 32 * t;
    int n = i / 9, l = (i % 9) / 3, m = i % 3;
    pw[t] = (i < pg_d) ? 0.5 * l + (m - l) * 0.5 * c_alpha[n] : 0.0;  // :673
  }
  const double invD = 1.0 / (1.0 + 1e-10);
  __syncthreads();

  for (int64_t base = pb; base < pe; base += ga.PB) {
    // ---- phase A: warp s folds pair base+s and stages its genotype row ----------------------
    if (warp < ga.PB) {
      const int64_t p = base + warp;
      double* slot = s_slots + (size_t)warp * ga.slot_d;
      bool has = false;
      if (p < pe) {
        const int32_t snp = a.pair_snp[p];
        has = a.has_gp ? (a.has_gp[snp] != 0) : true;
        if (has) {
          const uint32_t r0 = a.pair_rd[p], r1 = a.pair_rd[p + 1];
          double val[VPL];
#pragma unroll
          for (int t = 0; t < VPL; ++t) val[t] = 1.0;
          uint32_t cnt = 0;
          for (uint32_t r = r0; r < r1; ++r) {
            uint32_t aq = a.rd_aq[r];
            uint32_t al = aq >> 6;
            if (al == 2) continue;
            double err = s_err[aq & 63u];
            double mat = 1.0 - err, e3 = err / 3.0;
            double pR = (al == 0) ? mat : e3, pA = (al == 1) ? mat : e3;
#pragma unroll
            for (int t = 0; t < VPL; ++t) val[t] *= (pR * (1.0 - pw[t]) + pA * pw[t]);  // :685
            if ((++cnt & 7u) == 0u) {
              double mx = 0.0;
#pragma unroll
              for (int t = 0; t < VPL; ++t) if (lane + 32 * t < pg_d) mx = fmax(mx, val[t]);
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
              double ri = 1.0 / mx;
#pragma unroll
              for (int t = 0; t < VPL; ++t) val[t] *= ri;
            }
          }
          double mx = 0.0;
#pragma unroll
          for (int t = 0; t < VPL; ++t) if (lane + 32 * t < pg_d) mx = fmax(mx, val[t]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          const double ri = 1.0 / mx;
#pragma unroll
          for (int t = 0; t < VPL; ++t)
            if (lane + 32 * t < pg_d) slot[lane + 32 * t] = fma(val[t], ri, 1e-10) * invD;  // :704-725
          const double* src = a.gp + (size_t)snp * g_d;
          for (int i = lane; i < g_d; i += 32) slot[pg_d + i] = src[i];
        }
      }
      if (lane == 0) s_has[warp] = has ? 1 : 0;
    }
    __syncthreads();
    // ---- phase B: v[n][k][l] = sum_m pG[n][l][m] g_k[m] for the planes of this tile -----------
    for (int idx = tid; idx < ga.PB * nvv; idx += 256) {
      int s = idx / nvv, r = idx - s * nvv;
      if (!s_has[s]) continue;
      const double* slot = s_slots + (size_t)s * ga.slot_d;
      int nn = r / g_d + nlo, kl = r % g_d, k = kl / 3, l = kl % 3;
      const double* pg = slot + nn * 9 + l * 3;
      const double* g = slot + pg_d + k * 3;
      s_slots[(size_t)s * ga.slot_d + pg_d + g_d + r] = pg[0] * g[0] + pg[1] * g[1] + pg[2] * g[2];
    }
    __syncthreads();
    // ---- phase C: every thread updates its EPT accumulators -----------------------------------
    for (int s = 0; s < ga.PB; ++s) {
      if (!s_has[s]) continue;
      const double* g = s_slots + (size_t)s * ga.slot_d + pg_d;
      const double* v = g + g_d;
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        if (gidx[i] >= 0) {
          const double* gj = g + goff[i];
          const double* vk = v + voff[i];
          acc[i] *= (gj[0] * vk[0] + gj[1] * vk[1] + gj[2] * vk[2]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < EPT; ++i) pscl_renorm(acc[i], ex[i]);
    __syncthreads();
  }
  double* out = a.partial + (size_t)(item - a.item_base) * ((size_t)nv2 * na);
#pragma unroll
  for (int i = 0; i < EPT; ++i)
    if (gidx[i] >= 0) out[gidx[i]] = pscl_prod_log(acc[i], ex[i]);
}

// ------------------------------------------------------------------------------------------------
// per-cell epilogue (cmd_cram_demuxlet.cpp:788-991)
// ------------------------------------------------------------------------------------------------
struct Top2 {
  double v1, v2;
  int i1, i2;
};
// total order of the reference's scans (:827-837, :883-906): larger value first, earlier scan
// index first among equals; entries never beat the -1e300 initial value unless strictly larger.
__device__ __forceinline__ bool top_before(double va, int ia, double vb, int ib) {
  return (va > vb) || (va == vb && ia < ib);
}
__device__ __forceinline__ void top2_push(Top2& t, double v, int i) {
  if (!(v > -1e300)) return;
  if (top_before(v, i, t.v1, t.i1)) { t.v2 = t.v1; t.i2 = t.i1; t.v1 = v; t.i1 = i; }
  else if (top_before(v, i, t.v2, t.i2)) { t.v2 = v; t.i2 = i; }
}
__device__ __forceinline__ Top2 top2_shfl_merge(Top2 t, int o) {
  double ov1 = __shfl_xor_sync(0xffffffffu, t.v1, o), ov2 = __shfl_xor_sync(0xffffffffu, t.v2, o);
  int oi1 = __shfl_xor_sync(0xffffffffu, t.i1, o), oi2 = __shfl_xor_sync(0xffffffffu, t.i2, o);
  top2_push(t, ov1, oi1);
  top2_push(t, ov2, oi2);
  return t;
}

struct EpiArgs {
  const int64_t* cell_ptr;
  const int32_t* cell_item_ptr;
  double* partial;          // [items][G]; the first item's row is overwritten with the cell sum
  pscl_demux_cell* cells;   // [cells in batch], indexed from out_base
  double* grid;             // nullable: [cells][G]
  int32_t cell_begin;       // first cell of this batch
  int32_t out_base;         // index of cell_begin inside cells[] / grid[]
  int32_t item_base;        // first item of this batch
  int32_t nv, nalpha;
  unsigned long long inv_nv, inv_na;  // ceil(2^40 / d): idx / d == (idx * inv) >> 40, exact while idx * d < 2^40
  double doublet_prior;
};
__device__ __forceinline__ int epi_div(int x, unsigned long long inv) { return (int)(((unsigned long long)(unsigned)x * inv) >> 40); }

__global__ void __launch_bounds__(128) k_demux_epilogue(EpiArgs a) {
  const int c = a.cell_begin + blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nv = a.nv, na = a.nalpha, G = nv * nv * na;
  const int ia = a.cell_item_ptr[c], ib = a.cell_item_ptr[c + 1];
  double* sum_row = a.partial + (size_t)(ia - a.item_base) * G;
  double* grid_row = a.grid ? a.grid + (size_t)(a.out_base + blockIdx.x) * G : nullptr;
  const double lsp = log((1.0 - a.doublet_prior) / nv);                       // :793
  const double ldp1 = log(a.doublet_prior / nv / (nv - 1.) / (na - 1.));      // :794
  const double ldp2 = log(a.doublet_prior / nv / (nv - 1.) / (na - 1.) * 2);  // :795

  Top2 sng = {-1e300, -1e300, 0x7fffffff, 0x7fffffff}, dbl = sng;
  double mx_all = -1e-300, mx_sng = -1e-300;  // the (sic) -1e-300 start of :791 is a term of both sums
  for (int idx = tid; idx < G; idx += 128) {
    const int jk = epi_div(idx, a.inv_na), n = idx - jk * na, j = epi_div(jk, a.inv_nv), k = jk - j * nv;
    const bool is_s = (n == 0 && k == 0), is_d = (n >= 1 && j != k);
    double x = __longlong_as_double(0x7ff8000000000000ll);
    if (is_s || is_d) {
      x = 0.0;  // memset(llksAB,0) :643 — a cell without items keeps zeros
      for (int it = ia; it < ib; ++it) x += a.partial[(size_t)(it - a.item_base) * G + idx];
      if (ib > ia) sum_row[idx] = x;
      if (is_s) {
        top2_push(sng, x, j);
        double t = x + lsp;
        mx_all = fmax(mx_all, t); mx_sng = fmax(mx_sng, t);
      } else {
        top2_push(dbl, x, idx);
        if (c_alpha[n] == 0.5) { if (k < j) mx_all = fmax(mx_all, x + ldp2); }  // :812-815
        else mx_all = fmax(mx_all, x + ldp1);
      }
    }
    if (grid_row) grid_row[idx] = x;
  }
  __shared__ Top2 s_sng[4], s_dbl[4];
  __shared__ double s_mx[2][4], s_sum[2][4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sng = top2_shfl_merge(sng, o);
    dbl = top2_shfl_merge(dbl, o);
    mx_all = fmax(mx_all, __shfl_xor_sync(0xffffffffu, mx_all, o));
    mx_sng = fmax(mx_sng, __shfl_xor_sync(0xffffffffu, mx_sng, o));
  }
  if (lane == 0) { s_sng[warp] = sng; s_dbl[warp] = dbl; s_mx[0][warp] = mx_all; s_mx[1][warp] = mx_sng; }
  __syncthreads();
  mx_all = fmax(fmax(s_mx[0][0], s_mx[0][1]), fmax(s_mx[0][2], s_mx[0][3]));
  mx_sng = fmax(fmax(s_mx[1][0], s_mx[1][1]), fmax(s_mx[1][2], s_mx[1][3]));
  // second pass: sum of exp (the logAdd chains of :804-821 evaluated as max + log(sum exp))
  double se_all = 0.0, se_sng = 0.0;
  for (int idx = tid; idx < G; idx += 128) {
    const int jk = epi_div(idx, a.inv_na), n = idx - jk * na, j = epi_div(jk, a.inv_nv), k = jk - j * nv;
    if (n == 0 && k == 0) {
      double x = (ib > ia) ? sum_row[idx] : 0.0;
      se_all += exp(x + lsp - mx_all);
      se_sng += exp(x + lsp - mx_sng);
    } else if (n >= 1 && j != k) {
      double x = (ib > ia) ? sum_row[idx] : 0.0;
      if (c_alpha[n] == 0.5) { if (k < j) se_all += exp(x + ldp2 - mx_all); }
      else se_all += exp(x + ldp1 - mx_all);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se_all += __shfl_xor_sync(0xffffffffu, se_all, o);
    se_sng += __shfl_xor_sync(0xffffffffu, se_sng, o);
  }
  if (lane == 0) { s_sum[0][warp] = se_all; s_sum[1][warp] = se_sng; }
  __syncthreads();
  if (tid != 0) return;
  for (int wv = 1; wv < 4; ++wv) {
    top2_push(sng, s_sng[wv].v1, s_sng[wv].i1); top2_push(sng, s_sng[wv].v2, s_sng[wv].i2);
    top2_push(dbl, s_dbl[wv].v1, s_dbl[wv].i1); top2_push(dbl, s_dbl[wv].v2, s_dbl[wv].i2);
  }
  se_all = s_sum[0][0] + s_sum[0][1] + s_sum[0][2] + s_sum[0][3] + exp(-1e-300 - mx_all);
  se_sng = s_sum[1][0] + s_sum[1][1] + s_sum[1][2] + s_sum[1][3] + exp(-1e-300 - mx_sng);
  const double sumLLK = mx_all + log(se_all), sngLLK = mx_sng + log(se_sng);

  const int sBest = (sng.i1 == 0x7fffffff) ? -1 : sng.i1, sNext = (sng.i2 == 0x7fffffff) ? -1 : sng.i2;
  const double sngBestLLK = sng.v1, sngNextLLK = sng.v2, dblBestLLK = dbl.v1, dblNextLLK = dbl.v2;
  int dBest1 = -1, dBest2 = -1, dBestA = -1, dNext1 = -1, dNext2 = -1, dNextA = -1;
  if (dbl.i1 != 0x7fffffff) { dBestA = dbl.i1 % na; dBest2 = (dbl.i1 / na) % nv; dBest1 = dbl.i1 / na / nv; }
  if (dbl.i2 != 0x7fffffff) { dNextA = dbl.i2 % na; dNext2 = (dbl.i2 / na) % nv; dNext1 = dbl.i2 / na / nv; }

  pscl_demux_cell o;
  memset(&o, 0, sizeof(o));
  o.n_snps = (int32_t)(a.cell_ptr[c + 1] - a.cell_ptr[c]);
  if (dblBestLLK > sngBestLLK + 2) {  // :925-946
    o.type = PSCL_DBL;
    o.best_pp = exp(dblBestLLK + ((dBestA >= 0 && c_alpha[dBestA] == 0.5) ? ldp2 : ldp1) - sumLLK);
    o.best_j = dBest1; o.best_k = dBest2; o.best_llk = dblBestLLK; o.best_a = dBestA;
    if (dblNextLLK > sngBestLLK + 2) { o.next_j = dNext1; o.next_k = dNext2; o.next_llk = dblNextLLK; o.next_a = dNextA; }
    else { o.next_j = o.next_k = sBest; o.next_llk = sngBestLLK; o.next_a = 0; }
  } else {
    o.type = (sngBestLLK > sngNextLLK + 2) ? PSCL_SNG : PSCL_AMB;  // :947 / :968
    o.best_pp = sngBestLLK + lsp - sumLLK;                         // no exp (:949, :970)
    o.best_j = o.best_k = sBest; o.best_llk = sngBestLLK; o.best_a = 0;
    if (dblBestLLK > sngNextLLK + 2) { o.next_j = dBest1; o.next_k = dBest2; o.next_llk = dblBestLLK; o.next_a = dBestA; }
    else { o.next_j = o.next_k = sNext; o.next_llk = sngNextLLK; o.next_a = 0; }
  }
  o.sng_best = sBest; o.sng_next = sNext;
  o.dbl_best_j = dBest1; o.dbl_best_k = dBest2; o.dbl_best_a = dBestA;
  o.dbl_next_j = dNext1; o.dbl_next_k = dNext2; o.dbl_next_a = dNextA;
  o.sng_pp = exp(sngLLK - sumLLK);               // :990
  o.sng_only_pp = exp(sngBestLLK + lsp - sngLLK);  // :991
  o.sng_best_llk = sngBestLLK; o.sng_next_llk = sngNextLLK;
  o.dbl_best_llk = dblBestLLK; o.dbl_next_llk = dblNextLLK;
  o.sum_llk = sumLLK; o.sng_llk = sngLLK;
  a.cells[a.out_base + blockIdx.x] = o;
}

// Small grids (config 2: 128 entries per cell): one WARP per cell, eight cells per CTA, no block barriers —
// 10k single-cell CTAs were launch-rate bound (119 us per step against 0.74 ms for the accumulation kernel).
__global__ void __launch_bounds__(256) k_demux_epilogue_w(EpiArgs a, int n_cells) {
  const int lane = threadIdx.x & 31, cw = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (cw >= n_cells) return;
  const int c = a.cell_begin + cw;
  const int nv = a.nv, na = a.nalpha, G = nv * nv * na;
  const int ia = a.cell_item_ptr[c], ib = a.cell_item_ptr[c + 1];
  double* sum_row = a.partial + (size_t)(ia - a.item_base) * G;
  double* grid_row = a.grid ? a.grid + (size_t)(a.out_base + cw) * G : nullptr;
  const double lsp = log((1.0 - a.doublet_prior) / nv);                       // :793
  const double ldp1 = log(a.doublet_prior / nv / (nv - 1.) / (na - 1.));      // :794
  const double ldp2 = log(a.doublet_prior / nv / (nv - 1.) / (na - 1.) * 2);  // :795

  Top2 sng = {-1e300, -1e300, 0x7fffffff, 0x7fffffff}, dbl = sng;
  double mx_all = -1e-300, mx_sng = -1e-300;  // the (sic) -1e-300 start of :791 is a term of both sums
  for (int idx = lane; idx < G; idx += 32) {
    const int jk = epi_div(idx, a.inv_na), n = idx - jk * na, j = epi_div(jk, a.inv_nv), k = jk - j * nv;
    const bool is_s = (n == 0 && k == 0), is_d = (n >= 1 && j != k);
    double x = __longlong_as_double(0x7ff8000000000000ll);
    if (is_s || is_d) {
      x = 0.0;  // memset(llksAB,0) :643 — a cell without items keeps zeros
      for (int it = ia; it < ib; ++it) x += a.partial[(size_t)(it - a.item_base) * G + idx];
      if (ib > ia) sum_row[idx] = x;
      if (is_s) {
        top2_push(sng, x, j);
        double t = x + lsp;
        mx_all = fmax(mx_all, t); mx_sng = fmax(mx_sng, t);
      } else {
        top2_push(dbl, x, idx);
        if (c_alpha[n] == 0.5) { if (k < j) mx_all = fmax(mx_all, x + ldp2); }  // :812-815
        else mx_all = fmax(mx_all, x + ldp1);
      }
    }
    if (grid_row) grid_row[idx] = x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sng = top2_shfl_merge(sng, o);
    dbl = top2_shfl_merge(dbl, o);
    mx_all = fmax(mx_all, __shfl_xor_sync(0xffffffffu, mx_all, o));
    mx_sng = fmax(mx_sng, __shfl_xor_sync(0xffffffffu, mx_sng, o));
  }
  // second pass: sum of exp (the logAdd chains of :804-821 evaluated as max + log(sum exp))
  double se_all = 0.0, se_sng = 0.0;
  for (int idx = lane; idx < G; idx += 32) {
    const int jk = epi_div(idx, a.inv_na), n = idx - jk * na, j = epi_div(jk, a.inv_nv), k = jk - j * nv;
    if (n == 0 && k == 0) {
      double x = (ib > ia) ? sum_row[idx] : 0.0;
      se_all += exp(x + lsp - mx_all);
      se_sng += exp(x + lsp - mx_sng);
    } else if (n >= 1 && j != k) {
      double x = (ib > ia) ? sum_row[idx] : 0.0;
      if (c_alpha[n] == 0.5) { if (k < j) se_all += exp(x + ldp2 - mx_all); }
      else se_all += exp(x + ldp1 - mx_all);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se_all += __shfl_xor_sync(0xffffffffu, se_all, o);
    se_sng += __shfl_xor_sync(0xffffffffu, se_sng, o);
  }
  if (lane != 0) return;
  se_all += exp(-1e-300 - mx_all);
  se_sng += exp(-1e-300 - mx_sng);
  const double sumLLK = mx_all + log(se_all), sngLLK = mx_sng + log(se_sng);

  const int sBest = (sng.i1 == 0x7fffffff) ? -1 : sng.i1, sNext = (sng.i2 == 0x7fffffff) ? -1 : sng.i2;
  const double sngBestLLK = sng.v1, sngNextLLK = sng.v2, dblBestLLK = dbl.v1, dblNextLLK = dbl.v2;
  int dBest1 = -1, dBest2 = -1, dBestA = -1, dNext1 = -1, dNext2 = -1, dNextA = -1;
  if (dbl.i1 != 0x7fffffff) { dBestA = dbl.i1 % na; dBest2 = (dbl.i1 / na) % nv; dBest1 = dbl.i1 / na / nv; }
  if (dbl.i2 != 0x7fffffff) { dNextA = dbl.i2 % na; dNext2 = (dbl.i2 / na) % nv; dNext1 = dbl.i2 / na / nv; }

  pscl_demux_cell o;
  memset(&o, 0, sizeof(o));
  o.n_snps = (int32_t)(a.cell_ptr[c + 1] - a.cell_ptr[c]);
  if (dblBestLLK > sngBestLLK + 2) {  // :925-946
    o.type = PSCL_DBL;
    o.best_pp = exp(dblBestLLK + ((dBestA >= 0 && c_alpha[dBestA] == 0.5) ? ldp2 : ldp1) - sumLLK);
    o.best_j = dBest1; o.best_k = dBest2; o.best_llk = dblBestLLK; o.best_a = dBestA;
    if (dblNextLLK > sngBestLLK + 2) { o.next_j = dNext1; o.next_k = dNext2; o.next_llk = dblNextLLK; o.next_a = dNextA; }
    else { o.next_j = o.next_k = sBest; o.next_llk = sngBestLLK; o.next_a = 0; }
  } else {
    o.type = (sngBestLLK > sngNextLLK + 2) ? PSCL_SNG : PSCL_AMB;  // :947 / :968
    o.best_pp = sngBestLLK + lsp - sumLLK;                         // no exp (:949, :970)
    o.best_j = o.best_k = sBest; o.best_llk = sngBestLLK; o.best_a = 0;
    if (dblBestLLK > sngNextLLK + 2) { o.next_j = dBest1; o.next_k = dBest2; o.next_llk = dblBestLLK; o.next_a = dBestA; }
    else { o.next_j = o.next_k = sNext; o.next_llk = sngNextLLK; o.next_a = 0; }
  }
  o.sng_best = sBest; o.sng_next = sNext;
  o.dbl_best_j = dBest1; o.dbl_best_k = dBest2; o.dbl_best_a = dBestA;
  o.dbl_next_j = dNext1; o.dbl_next_k = dNext2; o.dbl_next_a = dNextA;
  o.sng_pp = exp(sngLLK - sumLLK);               // :990
  o.sng_only_pp = exp(sngBestLLK + lsp - sngLLK);  // :991
  o.sng_best_llk = sngBestLLK; o.sng_next_llk = sngNextLLK;
  o.dbl_best_llk = dblBestLLK; o.dbl_next_llk = dblNextLLK;
  o.sum_llk = sumLLK; o.sng_llk = sngLLK;
  a.cells[a.out_base + cw] = o;
}


#include "demux_poly.inl"
#ifdef PSCL_EXPERIMENTAL  // measured slower than k_demux_default: not part of the product library
#include "experimental/demux_cls.inl"
#include "experimental/demux_ab.inl"
#endif

// ------------------------------------------------------------------------------------------------
// host API
// ------------------------------------------------------------------------------------------------
// ---- genotype dictionary ----------------------------------------------------------------------------------------
// Hard-call genotypes (--field GT: gps = (1-err)*onehot + err*avg with avg from the SNP's genotype counts,
// sc_drop_seq.cpp:285-315) put few distinct triples into the table: 3 per combination of counts, 135 for 8 samples.
// Pass 1 lets every triple claim a slot of a 256-entry open-addressing table by a 64-bit hash of its bits; pass 2 writes
// the slot numbers as 8-bit codes and checks each triple bit for bit against its slot, so a hash clash or a 257th triple
// only raises the overflow flag and the row-gather kernel is used instead.
__device__ __forceinline__ unsigned long long dmx_triple_key(const double* t) {
  unsigned long long h = 0x9E3779B97F4A7C15ull;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    h ^= (unsigned long long)__double_as_longlong(t[i]);
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 29;
  }
  return h | 1ull;  // 0 means "free slot"
}
__global__ void k_geno_dict_claim(const double* __restrict__ gp, int32_t V, int32_t nv, unsigned long long* keys, double* dict, int* over) {
  // one thread per SNP: its samples share a few triples, so keys this thread has already placed are skipped
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  unsigned long long seen0 = 0ull, seen1 = 0ull, seen2 = 0ull, seen3 = 0ull;
  for (int j = 0; j < nv; ++j) {
    const double* tp = gp + ((size_t)v * nv + j) * 3;
    const double t[3] = {tp[0], tp[1], tp[2]};
    const unsigned long long key = dmx_triple_key(t);
    if (key == seen0 || key == seen1 || key == seen2 || key == seen3) continue;
    seen3 = seen2; seen2 = seen1; seen1 = seen0; seen0 = key;
    int s = (int)((key >> 32) & (PSCL_DICT_N - 1)), probe = 0;
    for (; probe < PSCL_DICT_N; ++probe, s = (s + 1) & (PSCL_DICT_N - 1)) {
      unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(keys + s);
      if (k == 0ull) {
        k = atomicCAS(keys + s, 0ull, key);
        if (k == 0ull) { dict[3 * s] = t[0]; dict[3 * s + 1] = t[1]; dict[3 * s + 2] = t[2]; break; }
      }
      if (k == key) break;
    }
    if (probe == PSCL_DICT_N) { atomicOr(over, 1); return; }
  }
}
__global__ void k_geno_dict_codes(const double* __restrict__ gp, int32_t V, int32_t nv, const unsigned long long* __restrict__ keys,
                                  const double* __restrict__ dict, unsigned long long* __restrict__ code,
                                  unsigned long long* __restrict__ cls, int* over) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  unsigned long long c = 0;
  bool bad = false;
  for (int j = 0; j < nv; ++j) {
    const double* t = gp + ((size_t)v * nv + j) * 3;
    const unsigned long long key = dmx_triple_key(t);
    int s = (int)((key >> 32) & (PSCL_DICT_N - 1)), probe = 0;
    while (probe < PSCL_DICT_N && keys[s] != key) { s = (s + 1) & (PSCL_DICT_N - 1); ++probe; }
    if (probe == PSCL_DICT_N) { bad = true; break; }
    bad |= __double_as_longlong(dict[3 * s]) != __double_as_longlong(t[0]) || __double_as_longlong(dict[3 * s + 1]) != __double_as_longlong(t[1]) ||
           __double_as_longlong(dict[3 * s + 2]) != __double_as_longlong(t[2]);
    c |= (unsigned long long)s << (8 * j);
  }
  code[v] = c;
  if (bad) atomicOr(over, 1);
  // genotype classes: distinct codes in order of first appearance (a fourth — a missing call among hard ones — only raises
  // bit 1 of the flag: the dictionary kernel takes that table)
  unsigned long long w = 0;
  int d0 = -1, d1 = -1, d2 = -1;
  bool many = false;
  for (int j = 0; j < nv; ++j) {
    const int s = (int)((c >> (8 * j)) & 255ull);
    int k;
    if (s == d0 || d0 < 0) { d0 = s; k = 0; }
    else if (s == d1 || d1 < 0) { d1 = s; k = 1; }
    else if (s == d2 || d2 < 0) { d2 = s; k = 2; }
    else { many = true; k = 0; }
    w |= (unsigned long long)k << (32 + 2 * j);
  }
  w |= (unsigned long long)(d0 < 0 ? 0 : d0) | (unsigned long long)(d1 < 0 ? 0 : d1) << 8 | (unsigned long long)(d2 < 0 ? 0 : d2) << 16;
  cls[v] = w;
  if (many && !bad) atomicOr(over, 2);
}

// ABI 4: genotype table from the reader's raw posteriors, mixed with the genotype error on the device.  One thread per
// SNP walks its samples in order, as sc_drop_seq.cpp:288-315 does; explicit _rn arithmetic keeps the compiler from
// contracting (1-err)*gp + err*avg into an FMA, so the table is bit-identical to the host-mixed one.
__global__ void k_geno_mix(const float* __restrict__ f32, const uint8_t* __restrict__ gt8, const double* __restrict__ err_snp, double err0,
                           int32_t V, int32_t nv, const uint8_t* __restrict__ has_gp, double* __restrict__ gp, int* bad) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  if (has_gp && !has_gp[v]) return;
  double avg0 = 1e-10, avg1 = 1e-10, avg2 = 1e-10;
  double* out = gp + (size_t)v * nv * 3;
  for (int j = 0; j < nv; ++j) {
    double g0, g1, g2;
    if (gt8) {
      const uint8_t c = gt8[(size_t)v * nv + j];
      if (c > 2) { atomicExch(bad, 1); return; }
      g0 = c == 0 ? 1.0 : 0.0; g1 = c == 1 ? 1.0 : 0.0; g2 = c == 2 ? 1.0 : 0.0;
    } else {
      const float* f = f32 + ((size_t)v * nv + j) * 3;
      g0 = (double)f[0]; g1 = (double)f[1]; g2 = (double)f[2];
    }
    out[3 * j] = g0; out[3 * j + 1] = g1; out[3 * j + 2] = g2;
    avg0 = __dadd_rn(avg0, g0); avg1 = __dadd_rn(avg1, g1); avg2 = __dadd_rn(avg2, g2);
  }
  const double sum = __dadd_rn(__dadd_rn(avg0, avg1), avg2);
  avg0 = __ddiv_rn(avg0, sum); avg1 = __ddiv_rn(avg1, sum); avg2 = __ddiv_rn(avg2, sum);
  double err = err_snp ? err_snp[v] : err0;
  if (err > 0.999) err = 0.999;
  if (err < 0) err = 0;
  if (err > 0) {
    const double keep = __dsub_rn(1.0, err);
    for (int j = 0; j < nv; ++j) {
      out[3 * j] = __dadd_rn(__dmul_rn(keep, out[3 * j]), __dmul_rn(err, avg0));
      out[3 * j + 1] = __dadd_rn(__dmul_rn(keep, out[3 * j + 1]), __dmul_rn(err, avg1));
      out[3 * j + 2] = __dadd_rn(__dmul_rn(keep, out[3 * j + 2]), __dmul_rn(err, avg2));
    }
  }
}

// Device room for a genotype table (and its has_gp bytes): kept from one table to the next, grown when needed.
static int demux_geno_reserve(pscl_ctx* ctx, size_t gp_bytes, size_t has_gp_bytes) {
  int rc;
  if ((rc = pscl_reserve(ctx, &ctx->gp, &ctx->gp_cap, gp_bytes)) != PSCL_OK) return rc;
  ctx->has_gp = nullptr;
  if (has_gp_bytes) {
    if ((rc = pscl_reserve(ctx, &ctx->has_gp_buf, &ctx->has_gp_cap, has_gp_bytes)) != PSCL_OK) return rc;
    ctx->has_gp = ctx->has_gp_buf;
  }
  return PSCL_OK;
}

// Tail of every way a genotype table gets onto the device (host copy, on-device mixing, NVLink peer copy): the
// dictionary form of the table for k_demux_default (2 <= nv <= 8) — two small kernels, flag to a pinned word.
static int demux_geno_finish(pscl_ctx* ctx, int32_t n_samples, int32_t n_snps) {
  ctx->nv = n_samples;
  ctx->geno_V = n_snps;
  ctx->dict_built = false;
  if (n_samples <= 8 && n_snps > 0 && ctx->h_dict_over) {
    if (!ctx->gp_dict) {
      PSCL_CUDA(ctx, cudaMalloc((void**)&ctx->gp_dict, sizeof(double) * 3 * PSCL_DICT_N));
      PSCL_CUDA(ctx, cudaMalloc((void**)&ctx->gp_dict_key, sizeof(unsigned long long) * PSCL_DICT_N));
      PSCL_CUDA(ctx, cudaMalloc((void**)&ctx->gp_dict_over, sizeof(int)));
    }
    // (buffers of a context are kept from one table to the next: a run per call must not pay cudaMalloc / cudaFree, which
    // also synchronise the device)
    int rc;
    if ((rc = pscl_reserve(ctx, &ctx->gp_code, &ctx->gp_code_cap, sizeof(unsigned long long) * (size_t)n_snps)) != PSCL_OK) return rc;
    if ((rc = pscl_reserve(ctx, &ctx->gp_cls, &ctx->gp_cls_cap, sizeof(unsigned long long) * (size_t)n_snps)) != PSCL_OK) return rc;
    PSCL_CUDA(ctx, cudaMemsetAsync(ctx->gp_dict, 0, sizeof(double) * 3 * PSCL_DICT_N, ctx->stream));
    PSCL_CUDA(ctx, cudaMemsetAsync(ctx->gp_dict_key, 0, sizeof(unsigned long long) * PSCL_DICT_N, ctx->stream));
    PSCL_CUDA(ctx, cudaMemsetAsync(ctx->gp_dict_over, 0, sizeof(int), ctx->stream));
    k_geno_dict_claim<<<(unsigned)((n_snps + 255) / 256), 256, 0, ctx->stream>>>(ctx->gp, n_snps, n_samples, ctx->gp_dict_key, ctx->gp_dict,
                                                                                 ctx->gp_dict_over);
    k_geno_dict_codes<<<(unsigned)((n_snps + 255) / 256), 256, 0, ctx->stream>>>(ctx->gp, n_snps, n_samples, ctx->gp_dict_key, ctx->gp_dict,
                                                                                 ctx->gp_code, ctx->gp_cls, ctx->gp_dict_over);
    ctx->launches += 2;
    PSCL_CUDA(ctx, cudaGetLastError());
    *ctx->h_dict_over = 1;
    PSCL_CUDA(ctx, cudaMemcpyAsync(ctx->h_dict_over, ctx->gp_dict_over, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->dict_built = true;
  }
  PSCL_CUDA(ctx, cudaEventRecord(ctx->ev_dict, ctx->stream));  // both pinned flag words are valid once this completes
  return PSCL_OK;
}

extern "C" int pscl_demux_set_geno(pscl_ctx* ctx, const pscl_geno* geno, int32_t n_snps) {
  if (!ctx) return PSCL_EINVAL;
  PsclScope scope__(ctx);
  if (!geno || (!geno->gp && !geno->gp_f32 && !geno->gt8) || geno->n_samples < 2 || n_snps < 0)
    return pscl_fail(ctx, PSCL_EINVAL,
                     "pscl_demux_set_geno: need gp and n_samples >= 2 (the reference divides by nv-1, "
                     "cmd_cram_demuxlet.cpp:794)");
  PSCL_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->gpM) { cudaFree(ctx->gpM); ctx->gpM = nullptr; }
  if (ctx->gpS) { cudaFree(ctx->gpS); ctx->gpS = nullptr; }
  size_t bytes = sizeof(double) * (size_t)n_snps * geno->n_samples * 3;
  int rc;
  if ((rc = demux_geno_reserve(ctx, bytes, geno->has_gp ? (size_t)n_snps : 0)) != PSCL_OK) return rc;
  if (geno->has_gp) PSCL_CUDA(ctx, cudaMemcpyAsync(ctx->has_gp, geno->has_gp, n_snps, cudaMemcpyHostToDevice, ctx->stream));
  if (geno->gp) {
    PSCL_CUDA(ctx, cudaMemcpyAsync(ctx->gp, geno->gp, bytes, cudaMemcpyHostToDevice, ctx->stream));
  } else if (n_snps > 0) {  // ABI 4 raw posteriors: copy the small form, mix on the device
    const size_t cells = (size_t)n_snps * geno->n_samples;
    // scratch for the small form: the context's own, kept from call to call (not the block cache: a block handed back there
    // could be given out again, in this same call, to an array the copy stream fills while k_geno_mix still reads it)
    float* d_f32 = nullptr; uint8_t* d_gt8 = nullptr; double* d_err = nullptr;
    int rc2;
    if ((rc2 = pscl_reserve(ctx, &ctx->geno_bad, &ctx->geno_bad_cap, sizeof(int))) != PSCL_OK) return rc2;
    int* const d_bad = ctx->geno_bad;
    PSCL_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    PSCL_CUDA(ctx, cudaMemsetAsync(ctx->gp, 0, bytes, ctx->stream));  // rows of SNPs without GP stay zero
    if ((rc2 = pscl_reserve(ctx, &ctx->geno_raw, &ctx->geno_raw_cap, geno->gt8 ? cells : sizeof(float) * cells * 3)) != PSCL_OK) return rc2;
    if (geno->gt8) {
      d_gt8 = ctx->geno_raw;
      PSCL_CUDA(ctx, cudaMemcpyAsync(d_gt8, geno->gt8, cells, cudaMemcpyHostToDevice, ctx->stream));
    } else {
      d_f32 = reinterpret_cast<float*>(ctx->geno_raw);
      PSCL_CUDA(ctx, cudaMemcpyAsync(d_f32, geno->gp_f32, sizeof(float) * cells * 3, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (geno->geno_err_snp) {
      if ((rc2 = pscl_reserve(ctx, &ctx->geno_err, &ctx->geno_err_cap, sizeof(double) * n_snps)) != PSCL_OK) return rc2;
      d_err = ctx->geno_err;
      PSCL_CUDA(ctx, cudaMemcpyAsync(d_err, geno->geno_err_snp, sizeof(double) * n_snps, cudaMemcpyHostToDevice, ctx->stream));
    }
    k_geno_mix<<<(unsigned)((n_snps + 127) / 128), 128, 0, ctx->stream>>>(d_f32, d_gt8, d_err, geno->geno_err, n_snps, geno->n_samples, ctx->has_gp,
                                                                           ctx->gp, d_bad);
    ctx->launches++;
    PSCL_CUDA(ctx, cudaGetLastError());
    // codes other than 0/1/2 are an input error; the flag travels to a pinned word and pscl_demux_score reads it
    if (ctx->h_geno_bad) {
      *ctx->h_geno_bad = 0;
      PSCL_CUDA(ctx, cudaMemcpyAsync(ctx->h_geno_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    }
  } else if (ctx->h_geno_bad) {
    *ctx->h_geno_bad = 0;
  }
  if (!geno->gp && ctx->h_geno_bad == nullptr) return pscl_fail(ctx, PSCL_ENOMEM, "no pinned flag word for the raw genotype forms");
  return demux_geno_finish(ctx, geno->n_samples, n_snps);
}

extern "C" int pscl_demux_keep_grid(pscl_ctx* ctx, int enable) {
  if (!ctx) return PSCL_EINVAL;
  PsclScope scope__(ctx);
  ctx->keep_grid = enable != 0;
  return PSCL_OK;
}
extern "C" int pscl_demux_force_general(pscl_ctx* ctx, int enable) {
  if (!ctx) return PSCL_EINVAL;
  PsclScope scope__(ctx);
  ctx->force_general = enable != 0;
  return PSCL_OK;
}
extern "C" int pscl_demux_last_kernel(const pscl_ctx* ctx) { return ctx ? ctx->dm_last_kernel : 0; }
extern "C" int pscl_demux_select_kernel(pscl_ctx* ctx, int which) {
  if (!ctx || which < 0 || which > 7) return PSCL_EINVAL;
#ifndef PSCL_EXPERIMENTAL
  if (which == 3 || which == 5)
    return pscl_fail(ctx, PSCL_EINVAL, "k_demux_cls / k_demux_ab are experiments (csrc/experimental/), built only with -DPSCL_EXPERIMENTAL");
#endif
  ctx->demux_kernel = which;
  return PSCL_OK;
}

// Threads per CTA of k_demux_default (compile-time; overridable for A/B builds, tools/build_variant.sh)
#ifndef PSCL_DICT_NT
#define PSCL_DICT_NT 256
#endif
#ifndef PSCL_ROWS_NT
#define PSCL_ROWS_NT 384
#endif
#ifndef PSCL_CLS_NT
#define PSCL_CLS_NT 256
#endif
// whether the automatic choice takes the genotype-class variant of the dictionary kernel when the table allows it
#ifndef PSCL_CLS_DEFAULT
#define PSCL_CLS_DEFAULT 1
#endif

template <int NV, int DELTA, int DICT>
static cudaError_t launch_default_v(pscl_ctx* ctx, const DemuxArgs& a) {
  constexpr int NT = DICT == 2 ? PSCL_CLS_NT : DICT ? PSCL_DICT_NT : PSCL_ROWS_NT;
  using Cfg = DefaultCfg<NV>;
  constexpr size_t SMEM = DICT == 2 ? Cfg::smem_cls(NT) : DICT ? Cfg::smem_dict(NT) : Cfg::smem_rows(NT);
  static_assert(SMEM <= 227 * 1024, "k_demux_default: shared memory budget");
  auto kern = k_demux_default<NV, DELTA, DICT, NT>;
  static bool attr_set[64] = {false};
  if (!attr_set[ctx->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) return e;
    attr_set[ctx->device & 63] = true;
  }
  const int wpc = NT / 32;  // warps (= work items in flight) per CTA
  int grid = ctx->sm_count;
  if (grid * wpc > a.n_work) grid = (a.n_work + wpc - 1) / wpc;
  if (grid < 1) grid = 1;
  kern<<<grid, NT, SMEM, ctx->stream>>>(a);
  return cudaGetLastError();
}

template <int NV>
static cudaError_t launch_default(pscl_ctx* ctx, const DemuxArgs& a) {
  const int delta = a.delta8 ? 2 : a.delta ? 1 : 0;  // staged run on 8-bit (ABI 6) / 16-bit (ABI 3) SNP gaps, or plain SNP ids
  if (a.gp_code && a.cls && delta == 0) return launch_default_v<NV, 0, 2>(ctx, a);
  if (a.gp_code) return delta == 2 ? launch_default_v<NV, 2, 1>(ctx, a) : delta == 1 ? launch_default_v<NV, 1, 1>(ctx, a) : launch_default_v<NV, 0, 1>(ctx, a);
  return delta == 2 ? launch_default_v<NV, 2, 0>(ctx, a) : delta == 1 ? launch_default_v<NV, 1, 0>(ctx, a) : launch_default_v<NV, 0, 0>(ctx, a);
}

// out_origin: the cell whose record is dm_cells[0]; total_cells: how many records the output holds.  pscl_demux_score
// passes (cell_begin, cell_end - cell_begin); the pipelined pscl_demux_run scores slice after slice into one array
// (0, n_cells), each slice in its own item order (slice_order).
static int demux_score_impl(pscl_ctx* ctx, const pscl_plp* plp, const pscl_demux_opts* opts, int32_t cell_begin, int32_t cell_end,
                            int32_t out_origin, int32_t total_cells, bool slice_order, bool alpha_set = false) {
  PsclScope scope__(ctx);
  if (!plp || !opts || !opts->alphas) return pscl_fail(ctx, PSCL_EINVAL, "pscl_demux_score: NULL argument");
  if (!ctx->gp) return pscl_fail(ctx, PSCL_ESTATE, "pscl_demux_score: call pscl_demux_set_geno first");
  if (ctx->geno_V != plp->V) return pscl_fail(ctx, PSCL_EINVAL, "genotype table has %d SNPs, pileup %d", ctx->geno_V, plp->V);
  if (cell_begin < 0 || cell_end > plp->C || cell_begin > cell_end)
    return pscl_fail(ctx, PSCL_EINVAL, "bad cell range [%d,%d) of %d", cell_begin, cell_end, plp->C);
  const int na = opts->n_alpha, nv = ctx->nv;
  if (na < 2 || na > PSCL_MAX_ALPHA)
    return pscl_fail(ctx, PSCL_EINVAL, "n_alpha must be in [2,%d] (the reference divides by nAlpha-1, cmd_cram_demuxlet.cpp:794)", PSCL_MAX_ALPHA);
  if (!(opts->doublet_prior > 0.0 && opts->doublet_prior < 1.0))
    return pscl_fail(ctx, PSCL_EINVAL, "doublet_prior must be in (0,1)");
  PSCL_CUDA(ctx, cudaSetDevice(ctx->device));
  double h_alpha[PSCL_MAX_ALPHA] = {0};
  for (int i = 0; i < na; ++i) h_alpha[i] = opts->alphas[i];
  // (a run has queued this copy ahead of its bulk copies: behind them it would hold the first kernel back until they are through)
  if (!alpha_set) PSCL_CUDA(ctx, cudaMemcpyToSymbolAsync(c_alpha, h_alpha, sizeof(h_alpha), 0, cudaMemcpyHostToDevice, ctx->stream));

  const int ncell = total_cells;
  const size_t G = (size_t)nv * nv * na;
  int rc;
  if ((rc = pscl_reserve(ctx, (pscl_demux_cell**)&ctx->dm_cells, &ctx->dm_cells_cap, sizeof(pscl_demux_cell) * (size_t)ncell)) != PSCL_OK) return rc;
  if (ctx->keep_grid && (rc = pscl_reserve(ctx, &ctx->dm_grid, &ctx->dm_grid_cap, sizeof(double) * G * ncell)) != PSCL_OK) return rc;
  ctx->dm_cell_begin = out_origin; ctx->dm_cell_end = out_origin + total_cells; ctx->dm_nalpha = na;

  const bool use_default = !ctx->force_general && ctx->demux_kernel != 2 && ctx->demux_kernel != 4 && na == 2 && h_alpha[0] == 0.0 &&
                           h_alpha[1] == 0.5 && nv >= 2 && nv <= 8;
  PSCL_CUDA(ctx, cudaEventSynchronize(ctx->ev_dict));  // set_geno's flag words (long done by now in every normal flow)
  if (ctx->h_geno_bad && *ctx->h_geno_bad)
    return pscl_fail(ctx, PSCL_EINVAL, "pscl_geno.gt8 holds a code other than 0/1/2 (missing calls need gp_f32)");
  bool use_dict = false;  // dictionary-coded genotypes: auto (0) and 6 take them when the table allows, 1 keeps the row gather
  if (use_default && ctx->dict_built && (ctx->demux_kernel == 0 || ctx->demux_kernel == 6 || ctx->demux_kernel == 7)) use_dict = (*ctx->h_dict_over & 1) == 0;
  // ... and by genotype classes (7; auto when PSCL_CLS_DEFAULT) when no SNP holds more than three distinct triples and the
  // SNP ids are plain (the staged in-kernel gap decoding keeps the dictionary form)
  const bool use_cls = use_dict && *ctx->h_dict_over == 0 && plp->n_stages <= 1 && ctx->gp_cls != nullptr &&
                       (ctx->demux_kernel == 7 || (ctx->demux_kernel == 0 && PSCL_CLS_DEFAULT && !getenv("PSCL_NO_CLS")));
  // every other shape: the polynomial kernel unless k_demux_general was asked for
  const bool use_poly = !use_default && !ctx->force_general && ctx->demux_kernel != 2;
  if (use_poly) {
    if ((rc = ply_build_stream(ctx, const_cast<pscl_plp*>(plp))) != PSCL_OK) return rc;
    double h_gamma[PSCL_MAX_ALPHA] = {0};
    for (int i = 0; i < na; ++i) h_gamma[i] = 0.5 * h_alpha[i];
    PSCL_CUDA(ctx, cudaMemcpyToSymbolAsync(c_gamma, h_gamma, sizeof(h_gamma), 0, cudaMemcpyHostToDevice, ctx->stream));
  }
#ifdef PSCL_EXPERIMENTAL
  const bool use_ws = use_default && (ctx->demux_kernel == 3 || ctx->demux_kernel == 5);
  const bool use_ab = use_default && ctx->demux_kernel == 5;
  if (use_ws) {
    if ((rc = dmx_build_classes(ctx, const_cast<pscl_plp*>(plp))) != PSCL_OK) return rc;
    if ((rc = dmx_build_geno_tables(ctx)) != PSCL_OK) return rc;
  }
#else
  const bool use_ws = false, use_ab = false;
#endif
  ctx->dm_last_kernel = use_ab ? 5 : use_ws ? 3 : use_default ? (use_cls ? 7 : use_dict ? 6 : 1) : use_poly ? 4 : 2;
  const std::vector<int32_t>& cip = plp->h_cell_item_ptr;
  size_t max_items = ctx->partial_budget_bytes / (G * sizeof(double));
  if (max_items < 1) max_items = 1;

  PSCL_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  float main_ms_known = 0.f;
  bool single_batch = true;
  int c0 = cell_begin;
  while (c0 < cell_end) {
    // batch = as many whole cells as fit the partial-grid budget
    int c1 = c0 + 1;
    while (c1 < cell_end && (size_t)(cip[c1 + 1] - cip[c0]) <= max_items) ++c1;
    const int ib = cip[c0], ie = cip[c1], nwork = ie - ib;
    if (c1 < cell_end || c0 > cell_begin) single_batch = false;
    if ((rc = pscl_reserve(ctx, &ctx->dm_partial, &ctx->dm_partial_cap, sizeof(double) * G * (size_t)(nwork > 0 ? nwork : 1))) != PSCL_OK) return rc;
    if (nwork > 0) {
      DemuxArgs a;
      a.pair_snp = plp->pair_snp; a.pair_rd = plp->pair_rd; a.rd_aq = plp->rd_aq;
      a.gp = ctx->gp; a.has_gp = ctx->has_gp; a.phred_err = ctx->phred_err; a.fold_tab = ctx->fold_tab;
      a.item_order = (ib == 0 && ie == plp->n_items) ? plp->item_order : (slice_order && c0 == cell_begin && c1 == cell_end) ? plp->item_order + ib : nullptr;
      a.item_pbeg = plp->item_pbeg; a.item_pend = plp->item_pend;
      a.partial = ctx->dm_partial; a.counter = ctx->dm_counter;
      a.item_base = ib; a.n_work = nwork; a.nv = nv; a.nalpha = na;
      if (use_dict) { a.gp_code = ctx->gp_code; a.gp_dict = ctx->gp_dict; }
      if (use_cls) { a.gp_code = ctx->gp_cls; a.cls = 1; }
      if (plp->n_stages > 1) {  // staged pscl_demux_run: ids from the gaps, slice by slice as they land
        if (!use_default || use_ws || a.item_order == nullptr)
          return pscl_fail(ctx, PSCL_ESTATE, "a staged pileup image can only be scored whole by k_demux_default");
        a.delta = plp->d_delta; a.delta8 = plp->d_delta8; a.gap_big = plp->d_gap_big; a.cell_gap_ptr = plp->d_cell_gap_ptr;
        a.n_gap_big = plp->n_gap_big;
        a.first = plp->d_first; a.cell_ptr = plp->cell_ptr; a.item_cell = plp->item_cell;
        a.flags = ctx->stage_flags; a.bad = plp->d_bad; a.n_snps = plp->V; a.n_stages = plp->n_stages;
        a.spin_limit = ctx->stage_spin_ticks;
        for (int k = 0; k <= plp->n_stages; ++k) a.stage_cell[k] = plp->stage_cell[k];
      }
      cudaError_t e = cudaSuccess;
#ifdef PSCL_EXPERIMENTAL
      if (use_ws) {
        PSCL_CUDA(ctx, cudaMemsetAsync(ctx->dm_counter, 0, sizeof(int), ctx->stream));
        ClsArgs wa;
        wa.pkt = plp->dmx_pkt; wa.deep = plp->dmx_deep;
        wa.gpM = ctx->gpM; wa.gpS = ctx->gpS; wa.fold_tab = ctx->fold_tab;
        const bool whole = (ib == 0 && ie == plp->n_items);
        wa.desc = whole ? plp->dmx_desc_sorted : plp->dmx_desc_nat + ib;
        wa.partial = ctx->dm_partial; wa.counter = ctx->dm_counter;
        wa.item_base = ib; wa.n_work = nwork; wa.n_snps = ctx->geno_V;
        if (use_ab) {
          switch (nv) {
            case 2: e = launch_ab<2>(ctx, wa); break;
            case 3: e = launch_ab<3>(ctx, wa); break;
            case 4: e = launch_ab<4>(ctx, wa); break;
            case 5: e = launch_ab<5>(ctx, wa); break;
            case 6: e = launch_ab<6>(ctx, wa); break;
            case 7: e = launch_ab<7>(ctx, wa); break;
            case 8: e = launch_ab<8>(ctx, wa); break;
          }
        } else {
          switch (nv) {
            case 2: e = launch_cls<2>(ctx, wa); break;
            case 3: e = launch_cls<3>(ctx, wa); break;
            case 4: e = launch_cls<4>(ctx, wa); break;
            case 5: e = launch_cls<5>(ctx, wa); break;
            case 6: e = launch_cls<6>(ctx, wa); break;
            case 7: e = launch_cls<7>(ctx, wa); break;
            case 8: e = launch_cls<8>(ctx, wa); break;
          }
        }
      } else
#endif
      if (use_default) {
        PSCL_CUDA(ctx, cudaMemsetAsync(ctx->dm_counter, 0, sizeof(int), ctx->stream));
        switch (nv) {
          case 2: e = launch_default<2>(ctx, a); break;
          case 3: e = launch_default<3>(ctx, a); break;
          case 4: e = launch_default<4>(ctx, a); break;
          case 5: e = launch_default<5>(ctx, a); break;
          case 6: e = launch_default<6>(ctx, a); break;
          case 7: e = launch_default<7>(ctx, a); break;
          case 8: e = launch_default<8>(ctx, a); break;
        }
      } else if (use_poly) {
        PolyArgs pa;
        pa.rec = plp->ply_rec; pa.rng = plp->ply_rng;
        pa.order = (ib == 0 && ie == plp->n_items) ? plp->item_order : nullptr;
        pa.gp = ctx->gp; pa.has_gp = ctx->has_gp; pa.pair_rd = plp->pair_rd; pa.rd_aq = plp->rd_aq; pa.phred_err = ctx->phred_err;
        pa.partial = ctx->dm_partial; pa.item_base = ib; pa.nv = nv; pa.na = na; pa.tiles = (nv + PLY_T - 1) / PLY_T;
        e = launch_poly(ctx, pa, nwork);
      } else {
        constexpr int EPT = 8;
        GeneralArgs ga;
        ga.d = a;
        ga.E = nv + (na - 1) * nv * nv;
        int planes = std::min(na, (256 * EPT) / (nv * nv) + 2);
        ga.nvv_max = planes * nv * 3;
        ga.slot_d = na * 9 + nv * 3 + ga.nvv_max;
        size_t slot_bytes = sizeof(double) * ga.slot_d;
        int PB = (int)std::min<size_t>(8, (200 * 1024 - 2048 - 64) / slot_bytes);
        if (PB < 1) return pscl_fail(ctx, PSCL_EINVAL, "n_samples=%d too large for the shared-memory staging of the general kernel", nv);
        ga.PB = PB;
        size_t smem = sizeof(double) * 256 + slot_bytes * PB + sizeof(int) * 8;
        static bool attr_set[64] = {false};
        if (!attr_set[ctx->device & 63]) {
          PSCL_CUDA(ctx, cudaFuncSetAttribute(k_demux_general<EPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
          attr_set[ctx->device & 63] = true;
        }
        dim3 grid((unsigned)nwork, (unsigned)((ga.E + 256 * EPT - 1) / (256 * EPT)));
        k_demux_general<EPT><<<grid, 256, smem, ctx->stream>>>(ga);
        e = cudaGetLastError();
      }
      ctx->launches++;
      if (e != cudaSuccess) return pscl_fail(ctx, PSCL_ECUDA, "demux kernel launch failed: %s", cudaGetErrorString(e));
    }
    if (single_batch) PSCL_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    EpiArgs ea;
    ea.cell_ptr = plp->cell_ptr; ea.cell_item_ptr = plp->cell_item_ptr; ea.partial = ctx->dm_partial;
    ea.cells = (pscl_demux_cell*)ctx->dm_cells; ea.grid = ctx->keep_grid ? ctx->dm_grid : nullptr;
    ea.cell_begin = c0; ea.out_base = c0 - out_origin; ea.item_base = ib; ea.nv = nv; ea.nalpha = na;
    ea.doublet_prior = opts->doublet_prior;
    ea.inv_nv = ((1ull << 40) + nv - 1) / nv; ea.inv_na = ((1ull << 40) + na - 1) / na;
    if (G <= 1024) k_demux_epilogue_w<<<(unsigned)((c1 - c0 + 7) / 8), 256, 0, ctx->stream>>>(ea, c1 - c0);
    else k_demux_epilogue<<<(unsigned)(c1 - c0), 128, 0, ctx->stream>>>(ea);
    ctx->launches++;
    PSCL_CUDA(ctx, cudaGetLastError());
    c0 = c1;
  }
  PSCL_CUDA(ctx, cudaEventRecord(ctx->ev2, ctx->stream));
  ctx->dm_timed = true;
  ctx->dm_single_batch = single_batch;
  (void)main_ms_known;
  return PSCL_OK;
}

extern "C" int pscl_demux_score(pscl_ctx* ctx, const pscl_plp* plp, const pscl_demux_opts* opts,
                                int32_t cell_begin, int32_t cell_end) {
  if (!ctx) return PSCL_EINVAL;
  return demux_score_impl(ctx, plp, opts, cell_begin, cell_end, cell_begin, cell_end - cell_begin, false);
}

extern "C" int pscl_demux_last_kernel_ms(pscl_ctx* ctx, float* ms_main, float* ms_total) {
  if (!ctx) return PSCL_EINVAL;
  PsclScope scope__(ctx);
  if (!ctx->dm_timed) return pscl_fail(ctx, PSCL_ESTATE, "no pscl_demux_score has run");
  PSCL_CUDA(ctx, cudaEventSynchronize(ctx->ev2));
  float t = 0.f, m = -1.f;
  PSCL_CUDA(ctx, cudaEventElapsedTime(&t, ctx->ev0, ctx->ev2));
  if (ctx->dm_single_batch) PSCL_CUDA(ctx, cudaEventElapsedTime(&m, ctx->ev0, ctx->ev1));
  if (ms_main) *ms_main = m;
  if (ms_total) *ms_total = t;
  return PSCL_OK;
}

extern "C" int pscl_demux_fetch(pscl_ctx* ctx, pscl_demux_cell* out, double* llk_grid) {
  if (!ctx) return PSCL_EINVAL;
  PsclScope scope__(ctx);
  if (!ctx->dm_timed) return pscl_fail(ctx, PSCL_ESTATE, "pscl_demux_fetch before pscl_demux_score");
  const int ncell = ctx->dm_cell_end - ctx->dm_cell_begin;
  if (out)
    PSCL_CUDA(ctx, cudaMemcpyAsync(out, ctx->dm_cells, sizeof(pscl_demux_cell) * (size_t)ncell, cudaMemcpyDeviceToHost, ctx->stream));
  if (llk_grid) {
    if (!ctx->keep_grid || !ctx->dm_grid) return pscl_fail(ctx, PSCL_ESTATE, "llk_grid requested but pscl_demux_keep_grid was off");
    size_t G = (size_t)ctx->nv * ctx->nv * ctx->dm_nalpha;
    PSCL_CUDA(ctx, cudaMemcpyAsync(llk_grid, ctx->dm_grid, sizeof(double) * G * ncell, cudaMemcpyDeviceToHost, ctx->stream));
  }
  PSCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PSCL_OK;
}

extern "C" int pscl_demux_run(pscl_ctx* ctx, const pscl_pileup* host, const pscl_geno* geno,
                              const pscl_demux_opts* opts, pscl_demux_cell* out, double* llk_grid) {
  if (!ctx) return PSCL_EINVAL;
  PsclScope scope__(ctx);
  if (!host || !geno || !opts || !out) return pscl_fail(ctx, PSCL_EINVAL, "pscl_demux_run: NULL argument");
  // PSCL_TRACE=1: wall-clock of every phase on stderr (each phase is drained first, so the sum is an upper bound)
  static const bool trace = getenv("PSCL_TRACE") != nullptr;
  auto now = [&]() { if (trace) cudaStreamSynchronize(ctx->stream); return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  // Staged run: with the ABI-3 delta arrays and the default kernel's shape, the SNP gaps (the largest array) cross PCIe in
  // slices of whole cells on a second stream, each followed by a 4-byte flag.  The scoring kernel is launched once, as
  // soon as everything else is on the device; it takes its work items slice by slice, decodes the gaps itself and waits
  // on a slice's flag before touching it, so the scoring hides under the copy except for the last slice.  Cells are
  // independent (cmd_cram_demuxlet.cpp:636) and every cell is summed in the same order, so no record changes.
  // PSCL_STAGES=n overrides the slice count (1 = off).
  int stages = 1;
  if (!llk_grid && !ctx->keep_grid && !ctx->force_general && (ctx->demux_kernel <= 1 || ctx->demux_kernel == 6) && opts->alphas && opts->n_alpha == 2 &&
      opts->alphas[0] == 0.0 && opts->alphas[1] == 0.5 && geno->n_samples >= 2 && geno->n_samples <= 8 &&
      (host->pair_snp_delta16 || host->pair_snp_delta8) && host->cell_first_snp && host->n_pairs < ((int64_t)1 << 30)) {
    if (const char* sv = getenv("PSCL_STAGES")) stages = atoi(sv);
    else if (const char* sl = getenv("PSCL_SLICES")) stages = atoi(sl);  // the pipelined form, whatever the size
    else if (host->n_pairs >= ((int64_t)1 << 22)) stages = (int)std::min<int64_t>(PSCL_MAX_STAGES, host->n_pairs / 1250000);  // ~2.5 MB of gaps per slice
    if (stages < 1) stages = 1;
    if (stages > 1) {
      // a staged image is scored by ONE launch over all work items: fall back to the plain upload when their partial
      // grids would not fit the scratch budget (pscl_demux_score would then have to split the cells into batches)
      size_t items = 0;
      for (int32_t c = 0; c < host->n_cells; ++c) items += (size_t)((host->cell_ptr[c + 1] - host->cell_ptr[c] + PSCL_ITEM_PAIRS - 1) / PSCL_ITEM_PAIRS);
      const size_t G = (size_t)geno->n_samples * geno->n_samples * 2;
      if (items * G * sizeof(double) > ctx->partial_budget_bytes) stages = 1;
    }
  }
  static const bool timeline = getenv("PSCL_TIMELINE") != nullptr;
  ctx->tl_on = timeline;
  if (timeline) {
    for (auto& ev : ctx->tl) if (!ev) cudaEventCreate(&ev);
    cudaEventRecord(ctx->tl[0], ctx->stream);
    t_pscl_alloc_ms = 0.0; t_pscl_alloc_n = 0;
  }
  const auto th0 = std::chrono::steady_clock::now();
  // Pipelined run (the default for that shape): every copy of the call goes to the copy stream in one queue — small arrays,
  // base-call counts, base-calls, then the SNP gaps in `slices` slices of whole cells, an event behind each — while the
  // context's stream builds the genotype tables, scans the counts as soon as they have landed, unpacks the base-calls, and
  // then takes slice after slice: decode its gaps, score its cells (the plain kernel), finish their records.  What is left
  // after the last byte has crossed PCIe is one slice's worth of work.  PSCL_SLICES=n overrides the count (1 = off);
  // PSCL_STAGES selects the older form (one launch whose warps wait on flag words) instead.
  int slices = 1;
  if (stages > 1 && !getenv("PSCL_STAGES")) {
    slices = getenv("PSCL_SLICES") ? stages : std::min(6, stages);
    stages = 1;
  }
  const auto t0 = now();
  const bool deferred = ctx->h_bad != nullptr && ctx->copy_stream != nullptr;
  if (!deferred) slices = 1;
  int rc = pscl_demux_set_geno(ctx, geno, host->n_snps);  // genotypes first: their (small) copies lead the queue
  if (rc != PSCL_OK) return rc;
  bool alpha_set = false;
  if (opts->alphas && opts->n_alpha >= 2 && opts->n_alpha <= PSCL_MAX_ALPHA) {  // ... and the alpha grid (validated by the scoring call)
    double h_alpha[PSCL_MAX_ALPHA] = {0};
    for (int i = 0; i < opts->n_alpha; ++i) h_alpha[i] = opts->alphas[i];
    PSCL_CUDA(ctx, cudaMemcpyToSymbolAsync(c_alpha, h_alpha, sizeof(h_alpha), 0, cudaMemcpyHostToDevice, ctx->stream));
    alpha_set = true;
  }
  if (timeline) cudaEventRecord(ctx->tl[3], ctx->stream);
  const auto t1 = now();
  pscl_plp* plp = nullptr;
  rc = plp_upload_impl(ctx, host, &plp, stages, 0, 0, deferred, nullptr, slices);  // validity flag read below, with the records
  if (rc != PSCL_OK) return rc;
  const auto t2 = now();
  bool keep = ctx->keep_grid;
  int bad = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tl_groups;  // PSCL_TIMELINE: start / end of every group's scoring
  if (plp->n_slices > 1) {
    // the gaps of slice k are decoded as soon as they have landed; the cells are scored in `groups` launches (a launch per
    // slice would leave the persistent kernel's 1184 warps with two or three work items each)
    int groups = 3;
    if (const char* gv = getenv("PSCL_GROUPS")) groups = std::max(1, atoi(gv));
    groups = std::min(groups, plp->n_slices);
    int32_t g0 = 0;  // first cell not scored yet
    for (int k = 0; k < plp->n_slices && rc == PSCL_OK; ++k) {
      const int32_t c0 = plp->stage_cell[k], c1 = plp->stage_cell[k + 1];
      cudaError_t e = cudaStreamWaitEvent(ctx->stream, ctx->slice_ev[k], 0);
      if (e == cudaSuccess && c1 > c0 && plp->sl_full) {  // everything of the slice's cells in one launch: ids, offsets, base-calls
        k_decode_cells<<<(unsigned)(c1 - c0), PSCL_DEC8_NT, 0, ctx->stream>>>(plp->cell_ptr + c0, plp->d_first + c0, plp->d_delta8, plp->d_gap_big, plp->d_cell_gap_ptr + c0,
                                                                         plp->n_gap_big, plp->sl_n2, plp->sl_nbig, plp->sl_nblk, plp->sl_n_big, plp->sl_cell_rd + c0,
                                                                         plp->sl_rpk, plp->sl_rpal, plp->sl_read_bits, plp->N, c1 - c0, plp->V, plp->pair_snp, plp->pair_rd,
                                                                         plp->rd_aq, plp->d_bad);
        ctx->launches++;
        e = cudaGetLastError();
      } else if (e == cudaSuccess && c1 > c0) {
        if (plp->sl_reads && plp->sl_rb[k + 1] > plp->sl_rb[k]) {  // the slice's base-calls
          k_unpack_reads<<<(unsigned)((plp->sl_rb[k + 1] - plp->sl_rb[k] + 255) / 256), 256, 0, ctx->stream>>>(plp->sl_rpk, plp->sl_rpal, plp->sl_read_bits, plp->sl_rb[k],
                                                                                                          plp->sl_rb[k + 1], plp->rd_aq, plp->d_bad);
          ctx->launches++;
        }
        const unsigned grid = (unsigned)(((int64_t)(c1 - c0) * 32 + 255) / 256);
        if (plp->d_delta8) k_decode_snp8<<<(unsigned)(c1 - c0), PSCL_DEC8_NT, 0, ctx->stream>>>(plp->cell_ptr + c0, plp->d_first + c0, plp->d_delta8, plp->d_gap_big, plp->d_cell_gap_ptr + c0,
                                                                        plp->n_gap_big, c1 - c0, plp->V, plp->pair_snp, plp->d_bad);
        else k_decode_snp<<<grid, 256, 0, ctx->stream>>>(plp->cell_ptr + c0, plp->d_first + c0, plp->d_delta, 0, c1 - c0, plp->V, plp->pair_snp, plp->d_bad);
        ctx->launches++;
        e = cudaGetLastError();
      }
      if (e != cudaSuccess) { rc = pscl_fail(ctx, PSCL_ECUDA, "pipelined run: slice %d: %s", k, cudaGetErrorString(e)); break; }
      const bool group_end = (k + 1) * groups / plp->n_slices != k * groups / plp->n_slices || k + 1 == plp->n_slices;
      if (group_end && c1 > g0) {
        if (timeline && k + 1 == plp->n_slices) cudaEventRecord(ctx->tl[6], ctx->stream);
        cudaEvent_t ga = nullptr, gb = nullptr;
        if (timeline && tl_groups.size() < 32) { cudaEventCreate(&ga); cudaEventCreate(&gb); cudaEventRecord(ga, ctx->stream); }
        rc = demux_score_impl(ctx, plp, opts, g0, c1, 0, host->n_cells, true, alpha_set);
        if (ga) { cudaEventRecord(gb, ctx->stream); tl_groups.push_back({ga, gb}); }
        g0 = c1;
      }
    }
    ctx->dm_single_batch = false;
  } else if (plp->n_stages > 1) {
    rc = demux_score_impl(ctx, plp, opts, 0, host->n_cells, 0, host->n_cells, false, alpha_set);  // one launch; its warps wait on the slice flags
  } else {
    if (llk_grid) ctx->keep_grid = true;
    rc = demux_score_impl(ctx, plp, opts, 0, host->n_cells, 0, host->n_cells, false, alpha_set);
  }
  // the image's validity flag (malformed arrays, a slice that never arrived) comes back with the records: one drain per run
  int* const bad_dst = ctx->h_bad ? ctx->h_bad : &bad;
  if (rc == PSCL_OK && cudaMemcpyAsync(bad_dst, plp->d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
    rc = pscl_fail(ctx, PSCL_ECUDA, "run: flag read-back failed");
  const auto t3 = now();
  const auto th1 = std::chrono::steady_clock::now();
  if (rc == PSCL_OK) rc = pscl_demux_fetch(ctx, out, llk_grid);
  if (timeline && rc == PSCL_OK) {
    cudaEventRecord(ctx->tl[5], ctx->stream);
    cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->copy_stream);
    float v[8] = {0};
    cudaEvent_t evs[8] = {ctx->tl[1], ctx->tl[2], ctx->tl[3], ctx->tl[4], ctx->ev1, ctx->ev2, ctx->tl[5], ctx->tl[6]};
    for (int i = 0; i < 8; ++i) if (cudaEventElapsedTime(&v[i], ctx->tl[0], evs[i]) != cudaSuccess) { v[i] = -1.f; cudaGetLastError(); }
    for (auto& g : tl_groups) {
      float a0 = 0.f, a1 = 0.f;
      cudaEventElapsedTime(&a0, ctx->tl[0], g.first); cudaEventElapsedTime(&a1, ctx->tl[0], g.second);
      fprintf(stderr, "[timeline ms] group scored %.3f -> %.3f\n", a0, a1);
      cudaEventDestroy(g.first); cudaEventDestroy(g.second);
    }
    fprintf(stderr, "[timeline ms] arrays landed %.3f | gaps landed %.3f | geno tables %.3f | decoded %.3f | scored %.3f | epilogue %.3f | fetched %.3f | last group starts %.3f | host: %d allocator calls %.3f, enqueue %.3f, whole call %.3f\n",
            v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], t_pscl_alloc_n, t_pscl_alloc_ms, ms(th0, th1), ms(th0, std::chrono::steady_clock::now()));
  }
  else cudaStreamSynchronize(ctx->stream);  // the staging buffer and the caller's arrays are sources of queued copies
  bad = *bad_dst;
  if (rc == PSCL_OK && bad == 4) rc = pscl_fail(ctx, PSCL_ECUDA, "staged run: a slice of pair_snp_delta16 never reached the device");
  else if (rc == PSCL_OK && bad) rc = pscl_fail(ctx, PSCL_EINVAL, "%s", pscl_bad_pileup_msg(bad));
  const auto t4 = now();
  ctx->keep_grid = keep;
  std::string err = ctx->err;
  const int n_st = plp->n_stages;
  pscl_plp_free(ctx, plp);
  ctx->err = err;
  if (trace)
    fprintf(stderr, "[pscl_demux_run] set_geno %.3f ms | upload %.3f | score %.3f (%d slices) | fetch %.3f | free %.3f\n", ms(t0, t1), ms(t1, t2),
            ms(t2, t3), n_st, ms(t3, t4), ms(t4, now()));
  return rc;
}
