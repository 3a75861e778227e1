// demux_poly.inl — demuxlet pair-grid kernel for many samples and/or a dense alpha grid
// (part of popscle_b200.cu; replaces cmd_cram_demuxlet.cpp:655-747 for the shapes k_demux_default
// does not cover: nv > 8 or an alpha grid other than {0, 0.5}; config 4's 64 samples x 21 alphas).
//
// The reference evaluates, per (cell,SNP) pair and per (j,k,n),
//     term = sum_{l,m} g_j[l] g_k[m] pG_n[l][m],   pG_n[l][m] = (F(p)/max + 1e-10)/(1 + 1e-10),
//     F(p) = prod_reads (pR + (pA - pR) p),        p = l/2 + gamma_n (m - l),  gamma_n = alpha_n / 2
// (:666-725; p = 0.5 l + (m-l) 0.5 alpha, :673) — 9 multiply-adds per (j,k,n) after a per-pair table
// fill, i.e. ~7 nv^2 n_alpha flops per pair (626 kflop at config 4).  k_demux_general does exactly
// that and reaches 5 % of the FP64 peak there (tiles of 2048 grid entries, three block-wide barriers
// per 8 pairs, every tile re-folding the reads).
//
// This kernel uses that pG is a POLYNOMIAL in gamma_n whose degree is the number R of usable
// base-calls of the pair (R <= 1 for three quarters of all pairs, <= 3 for 99.6 %).  Expanding F
// around p0 = l/2 (exact at gamma = 0 and well conditioned for alpha <= 1/2: every term keeps its sign)
//     pG(l/2 + gamma (m-l)) = sum_e Phi_e[l] (m-l)^e gamma^e,   Phi_e[l] = F^(e)(l/2) / e! (scaled)
//     term(j,k,n) = sum_e K_e(j,k) gamma_n^e,   K_e = sum_l g_j[l] Phi_e[l] * sum_m g_k[m] (m-l)^e
// so one thread owns one (j,k), computes K_0..K_R once per pair from the three genotype moments of
// k (S, M, Q) and the Phi-weighted moments of j (R=1: 12 FP64 instructions, R=2: 30, R=3: 45), and
// then needs R FMAs + 1 multiply per alpha plane (Horner in gamma_n, gamma_n read from the constant
// bank) instead of 9 + 1.  The running products of the n_alpha planes of its (j,k) live in the
// thread's registers (mantissa + exponent; one log per entry per work item).  Pairs with R > 3
// (0.4 %) take the direct 9-FMA form from a per-pair table in shared memory.
//
// Work decomposition: CTA = (work item, 16 x 16 tile of the nv x nv sample grid), 256 threads; the
// tiles of one item are adjacent in the launch order, so the genotype rows they all read come out of
// L2.  Pairs are taken 8 at a time (warp w stages pair w: 16 + 16 genotype rows of 24 B, prefetched
// through registers one group ahead; classes M and D also get their Phi / pG tables from that warp),
// two block barriers per group.  The pairs of an item are processed in class order S | M | D
// (records of k_dmx_classify below, so R is uniform within a group except M's 2 vs 3).

#ifndef PLY_MINB
#define PLY_MINB 2         /* CTAs per SM the register allocation aims at (NPL <= 21) */
#endif
#ifndef PLY_EXSMEM
#define PLY_EXSMEM 1       /* 1: exponents of the running products in shared memory (touched once per 16 pairs) */
#endif
#define PLY_G 8            /* pairs per group = warps per CTA */
#define PLY_T 16           /* tile edge */
#define PLY_MAX_GRID (PSCL_MAX_ALPHA * 9)

static __constant__ double c_gamma[PSCL_MAX_ALPHA];  // alpha_n / 2

struct PolyArgs {
  const uint2* rec;         // [P] class-ordered records: S/M {snp, b0 | b1<<8 | b2<<16 | cnt<<24}, D {snp, original pair}
  const uint4* rng;         // [n_items] {begin, first M, first D, end} record positions of a work item
  const int32_t* order;     // nullable: work index -> item
  const double* gp;         // [V][nv][3]
  const uint8_t* has_gp;    // nullable
  const uint32_t* pair_rd;  // class D re-reads its base-calls
  const uint8_t* rd_aq;
  const double* phred_err;  // [256]
  double* partial;          // [items][nv*nv*na]
  int32_t item_base, nv, na, tiles;  // tiles per dimension
};

#define WS_NONE_CODES 0x00808080u /* three "no base-call" codes: allele 2, qual 0 = the all-ones row of both fold tables */

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


// class ranges of a work item inside the class-ordered record array: the item's records are the
// contiguous range [ib, ie) of its cell's range, whose classes are S | M | D with boundaries sm, md
struct DmxItemRanges { uint32_t lo[3], hi[3]; };
__device__ __forceinline__ DmxItemRanges dmx_item_ranges(const int64_t* cell_ptr, const int32_t* item_cell, const int64_t* item_pbeg,
                                                         const int64_t* item_pend, const unsigned long long* scan, int item) {
  const int c = item_cell[item];
  const int64_t c0 = cell_ptr[c], c1 = cell_ptr[c + 1];
  const unsigned long long s0 = scan[c0], s1 = scan[c1];
  const int64_t n_m = (uint32_t)(s1 - s0), n_d = (uint32_t)((s1 >> 32) - (s0 >> 32)), n_s = (c1 - c0) - n_m - n_d;
  const uint32_t ib = (uint32_t)item_pbeg[item], ie = (uint32_t)item_pend[item], sm = (uint32_t)(c0 + n_s), md = (uint32_t)(c0 + n_s + n_m);
  DmxItemRanges r;
  r.lo[0] = ib; r.hi[0] = min(ie, sm);
  r.lo[1] = max(ib, sm); r.hi[1] = min(ie, md);
  r.lo[2] = max(ib, md); r.hi[2] = ie;
  return r;
}
// one warp per work item: class-ordered flat records (class D keeps the original pair index) and the
// item's class boundaries
__global__ void k_ply_scatter(const int64_t* __restrict__ cell_ptr, const int32_t* __restrict__ item_cell,
                              const int64_t* __restrict__ item_pbeg, const int64_t* __restrict__ item_pend, int32_t n_items,
                              const uint2* __restrict__ rec_tmp, const unsigned long long* __restrict__ scan,
                              uint2* __restrict__ rec, uint4* __restrict__ rng) {
  const int item = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (item >= n_items) return;
  const int c = item_cell[item];
  const int64_t c0 = cell_ptr[c], c1 = cell_ptr[c + 1];
  const unsigned long long s0 = scan[c0], s1 = scan[c1];
  const uint32_t n_m = (uint32_t)(s1 - s0), n_d = (uint32_t)((s1 >> 32) - (s0 >> 32));
  const int64_t n_s = (c1 - c0) - n_m - n_d;
  if (lane == 0) {
    const DmxItemRanges r = dmx_item_ranges(cell_ptr, item_cell, item_pbeg, item_pend, scan, item);
    rng[item] = make_uint4(r.lo[0], max(r.hi[0], r.lo[0]), max(r.hi[1], max(r.hi[0], r.lo[0])), r.hi[2]);
  }
  for (int64_t p = item_pbeg[item] + lane; p < item_pend[item]; p += 32) {
    const unsigned long long sp = scan[p], kp = scan[p + 1] - sp;
    const uint32_t rank_m = (uint32_t)sp - (uint32_t)s0, rank_d = (uint32_t)(sp >> 32) - (uint32_t)(s0 >> 32);
    uint2 r = rec_tmp[p];
    int64_t pos;
    if (kp == 0ull) pos = c0 + ((p - c0) - rank_m - rank_d);
    else if (kp == 1ull) pos = c0 + n_s + rank_m;
    else { pos = c0 + n_s + n_m + rank_d; r.y = (uint32_t)p; }
    rec[pos] = r;
  }
}

template <int NPL>
__global__ void __launch_bounds__(256, NPL <= 21 ? PLY_MINB : 1) k_demux_poly(PolyArgs a) {
  // ---- shared memory -----------------------------------------------------------------------------
  __shared__ __align__(16) double s_rowJ[PLY_G][PLY_T][3];  // genotype rows of the tile's 16 j ...
  __shared__ __align__(16) double s_rowK[PLY_G][PLY_T][3];  // ... and 16 k samples, per pair of the group
  __shared__ __align__(16) double s_par[PLY_G][12];         // S: {c0, c1/2, c1}; M: Phi_e[l] at [3e + l]
  __shared__ int s_kind[PLY_G];                             // 15 skip | 0 S | 2, 3 M (degree) | 4 D (direct)
#if PLY_EXSMEM
  extern __shared__ int s_ex_dyn[];  // [NPL][256] exponents of the running products
  int (*s_ex)[256] = reinterpret_cast<int (*)[256]>(s_ex_dyn);
#endif
  __shared__ __align__(16) double s_tabS[3 * 64][2];        // per base-call code: {c0, c1} of pG(p) = c0 + c1 p
  __shared__ __align__(16) double s_tabR[3 * 64][2];        // per base-call code: {pR, pA} (:666-667)
  __shared__ __align__(16) double s_dir[PLY_G][PLY_MAX_GRID];  // class D: pG_n[l][m] at [9n + 3l + m]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nv = a.nv, na = a.na;
  for (int c = tid; c < 3 * 64; c += 256) {
    const int al = c >> 6;
    const double err = a.phred_err[c & 63], mat = 1.0 - err, e3 = err / 3.0;
    const double pR = al == 2 ? 1.0 : (al == 0 ? mat : e3), pA = al == 2 ? 1.0 : (al == 1 ? mat : e3);  // skipped read: factor 1
    const double inv = 1.0 / (fmax(pR, pA) * (1.0 + 1e-10));  // F linear: its maximum over the grid sits at p = 0 or p = 1
    s_tabR[c][0] = pR; s_tabR[c][1] = pA;
    s_tabS[c][0] = fma(pR, inv, 1e-10 / (1.0 + 1e-10));
    s_tabS[c][1] = (pA - pR) * inv;
  }

  const int w = blockIdx.y;
  const int item = a.order ? a.order[w] : a.item_base + w;
  const uint4 rg = a.rng[item];  // {begin, first M, first D, end}
  const int tile = blockIdx.x, jt = tile / a.tiles, kt = tile % a.tiles;
  const int tj = tid >> 4, tk = tid & 15;               // this thread's entry inside the tile
  const int J = jt * PLY_T + tj, K = kt * PLY_T + tk;
  // staging role: warp w loads pair w of the group; lanes 0-15 the tile's j rows, 16-31 its k rows
  const int my_sample = (lane < 16 ? jt : kt) * PLY_T + (lane & 15);
  const bool my_valid = my_sample < nv;

  double acc[NPL];
#if PLY_EXSMEM
#pragma unroll
  for (int n = 0; n < NPL; ++n) { acc[n] = 1.0; s_ex[n][tid] = 0; }
#else
  int ex[NPL];
#pragma unroll
  for (int n = 0; n < NPL; ++n) { acc[n] = 1.0; ex[n] = 0; }
#endif

  // grid points this lane evaluates when its warp sets up a class-M/D pair: i = lane + 32 t -> (n, l, m)
  constexpr int VPL = (PLY_MAX_GRID + 31) / 32;

  // ---- group pipeline ------------------------------------------------------------------------------
  // Groups never straddle a class boundary: [begin, first M), [first M, first D), [first D, end) are cut
  // into groups of 8 separately, so a group's class is that of its first record.
  const uint32_t bnd[4] = {rg.x, rg.y, rg.z, rg.w};
  auto group_of = [&](uint32_t pos, uint32_t& n, int& cls) {  // group starting at record `pos`
    cls = pos < bnd[1] ? 0 : pos < bnd[2] ? 1 : 2;
    n = min((uint32_t)PLY_G, bnd[cls + 1] - pos);
  };

  uint32_t pos = rg.x;  // records [pos, pos + n) form the current group
  // registers carrying the NEXT group's data while the current one is multiplied
  uint2 recN = make_uint2(0u, 0u);
  double g0N = 0.0, g1N = 0.0, g2N = 0.0;
  int hasN = 0;
  uint32_t nN = 0;
  int clsN = 0;
  auto issue = [&](uint32_t p0) {  // loads of the group starting at record p0 (this warp: pair p0 + warp)
    nN = 0; clsN = 0; hasN = 0;
    if (p0 < rg.w) {
      group_of(p0, nN, clsN);
      if ((uint32_t)warp < nN) {
        recN = a.rec[p0 + warp];
        hasN = a.has_gp ? (int)a.has_gp[recN.x] : 1;
        if (my_valid) {
          const double* src = a.gp + ((size_t)recN.x * nv + my_sample) * 3;
          g0N = src[0]; g1N = src[1]; g2N = src[2];
        }
      }
    }
  };
  issue(pos);
  __syncthreads();  // tables

  int since = 0;
  while (pos < rg.w) {
    // ---- stage the group whose loads are in registers ---------------------------------------------
    const uint32_t n = nN;
    const int cls = clsN;
    if ((uint32_t)warp < n) {
      double* dst = lane < 16 ? &s_rowJ[warp][lane][0] : &s_rowK[warp][lane - 16][0];
      dst[0] = g0N; dst[1] = g1N; dst[2] = g2N;
      const uint2 rec = recN;
      int kind = 15;
      if (hasN) {
        if (cls == 0) {
          kind = 0;
          if (lane == 0) {
            const double c0 = s_tabS[rec.y & 0xffu][0], c1 = s_tabS[rec.y & 0xffu][1];
            s_par[warp][0] = c0; s_par[warp][1] = 0.5 * c1; s_par[warp][2] = c1;
          }
        } else if (cls == 1) {
          // ---- class M: R = 2 or 3 usable base-calls; joint maximum over the (n,l,m) grid (:692-699) ----
          const int R = (int)(rec.y >> 24);
          kind = R;
          const double r0 = s_tabR[rec.y & 0xffu][0], a0 = s_tabR[rec.y & 0xffu][1];
          const double r1 = s_tabR[(rec.y >> 8) & 0xffu][0], a1 = s_tabR[(rec.y >> 8) & 0xffu][1];
          const double r2 = s_tabR[(rec.y >> 16) & 0xffu][0], a2 = s_tabR[(rec.y >> 16) & 0xffu][1];  // ones when R == 2
          const double b0 = a0 - r0, b1 = a1 - r1, b2 = a2 - r2;
          double mx = 0.0;
#pragma unroll
          for (int t = 0; t < VPL; ++t) {
            const int i = lane + 32 * t;
            if (i < na * 9) {
              const int nn = i / 9, l = (i % 9) / 3, m = i % 3;
              const double p = 0.5 * l + c_gamma[nn] * (m - l);
              mx = fmax(mx, fma(b0, p, r0) * fma(b1, p, r1) * fma(b2, p, r2));
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          if (lane < 3) {  // Phi_e[l] = F^(e)(l/2)/e!, the read factors taken exactly at p = 0, 1/2, 1
            const double inv = 1.0 / (mx * (1.0 + 1e-10));
            const double f0 = lane == 0 ? r0 : lane == 2 ? a0 : 0.5 * (r0 + a0);
            const double f1 = lane == 0 ? r1 : lane == 2 ? a1 : 0.5 * (r1 + a1);
            const double f2 = lane == 0 ? r2 : lane == 2 ? a2 : 0.5 * (r2 + a2);
            s_par[warp][0 + lane] = fma(f0 * f1 * f2, inv, 1e-10 / (1.0 + 1e-10));
            s_par[warp][3 + lane] = (b0 * f1 * f2 + f0 * b1 * f2 + f0 * f1 * b2) * inv;
            s_par[warp][6 + lane] = (b0 * b1 * f2 + b0 * f1 * b2 + f0 * b1 * b2) * inv;
            s_par[warp][9 + lane] = (b0 * b1 * b2) * inv;
          }
        } else {
          // ---- class D: > 3 usable base-calls; the full pG table, folded read by read ------------------
          kind = 4;
          const uint32_t q0 = a.pair_rd[rec.y], q1 = a.pair_rd[rec.y + 1];
          double val[VPL], pw[VPL];
#pragma unroll
          for (int t = 0; t < VPL; ++t) {
            const int i = lane + 32 * t, nn = min(i / 9, PSCL_MAX_ALPHA - 1), l = (i % 9) / 3, m = i % 3;
            pw[t] = 0.5 * l + c_gamma[nn] * (m - l);
            val[t] = 1.0;
          }
          uint32_t cnt = 0;
          for (uint32_t q = q0; q < q1; ++q) {
            const uint32_t aq = a.rd_aq[q];
            const double pR = s_tabR[aq][0], pA = s_tabR[aq][1];
#pragma unroll
            for (int t = 0; t < VPL; ++t) val[t] *= fma(pA - pR, pw[t], pR);
            if ((++cnt & 7u) == 0u) {  // deep pileups: rescale by the running max like :692-699
              double mx = 0.0;
#pragma unroll
              for (int t = 0; t < VPL; ++t) if (lane + 32 * t < na * 9) mx = fmax(mx, val[t]);
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
              const double ri = 1.0 / mx;
#pragma unroll
              for (int t = 0; t < VPL; ++t) val[t] *= ri;
            }
          }
          double mx = 0.0;
#pragma unroll
          for (int t = 0; t < VPL; ++t) if (lane + 32 * t < na * 9) mx = fmax(mx, val[t]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          const double inv = 1.0 / (mx * (1.0 + 1e-10));
#pragma unroll
          for (int t = 0; t < VPL; ++t)
            if (lane + 32 * t < na * 9) s_dir[warp][lane + 32 * t] = fma(val[t], inv, 1e-10 / (1.0 + 1e-10));
        }
      }
      if (lane == 0) s_kind[warp] = kind;
    }
    __syncthreads();
    const uint32_t pos_next = pos + n;
    issue(pos_next);  // the next group's loads fly while this one is multiplied

    // ---- multiply: every thread updates the NPL planes of its (J,K) -------------------------------------
    uint32_t kinds = 0;  // the group's 8 kinds, 4 bits each: no shared-memory round trip inside the pair loop
#pragma unroll
    for (int s = 0; s < PLY_G; ++s) kinds |= (uint32_t)(s_kind[s] & 15) << (4 * s);
    for (uint32_t s = 0; s < n; ++s) {
      const int kind = (int)((kinds >> (4 * s)) & 15u);
      if (kind == 15) continue;
      if (kind == 0 && s + 1 < n && ((kinds >> (4 * s + 4)) & 15u) == 0u) {
        // two class-S pairs at once: (K0a + g K1a)(K0b + g K1b) = c0 + g c1 + g^2 c2, so the pair of pairs costs
        // 2 FMAs + 1 multiply per plane instead of 2 x (1 + 1), and one trip through the loop instead of two
        const double ja0 = s_rowJ[s][tj][0], ja1 = s_rowJ[s][tj][1], ja2 = s_rowJ[s][tj][2];
        const double ka0 = s_rowK[s][tk][0], ka1 = s_rowK[s][tk][1], ka2 = s_rowK[s][tk][2];
        const double jb0 = s_rowJ[s + 1][tj][0], jb1 = s_rowJ[s + 1][tj][1], jb2 = s_rowJ[s + 1][tj][2];
        const double kb0 = s_rowK[s + 1][tk][0], kb1 = s_rowK[s + 1][tk][1], kb2 = s_rowK[s + 1][tk][2];
        const double Sja = ja0 + ja1 + ja2, Mja = fma(2.0, ja2, ja1), Ska = ka0 + ka1 + ka2, Mka = fma(2.0, ka2, ka1);
        const double Sjb = jb0 + jb1 + jb2, Mjb = fma(2.0, jb2, jb1), Skb = kb0 + kb1 + kb2, Mkb = fma(2.0, kb2, kb1);
        const double K0a = fma(s_par[s][1], Mja, s_par[s][0] * Sja) * Ska, K1a = s_par[s][2] * (Sja * Mka - Mja * Ska);
        const double K0b = fma(s_par[s + 1][1], Mjb, s_par[s + 1][0] * Sjb) * Skb, K1b = s_par[s + 1][2] * (Sjb * Mkb - Mjb * Skb);
        const double c0 = K0a * K0b, c1 = fma(K0a, K1b, K1a * K0b), c2 = K1a * K1b;
#pragma unroll
        for (int nn = 0; nn < NPL; ++nn) acc[nn] *= fma(fma(c2, c_gamma[nn], c1), c_gamma[nn], c0);
        ++s;
        continue;
      }
      const double gj0 = s_rowJ[s][tj][0], gj1 = s_rowJ[s][tj][1], gj2 = s_rowJ[s][tj][2];
      const double gk0 = s_rowK[s][tk][0], gk1 = s_rowK[s][tk][1], gk2 = s_rowK[s][tk][2];
      const double Sk = gk0 + gk1 + gk2, Mk = fma(2.0, gk2, gk1);
      if (kind == 0) {
        const double Sj = gj0 + gj1 + gj2, Mj = fma(2.0, gj2, gj1);
        const double c0 = s_par[s][0], c1h = s_par[s][1], c1 = s_par[s][2];
        const double K0 = fma(c1h, Mj, c0 * Sj) * Sk;
        const double K1 = c1 * (Sj * Mk - Mj * Sk);
#pragma unroll
        for (int nn = 0; nn < NPL; ++nn) acc[nn] *= fma(K1, c_gamma[nn], K0);
      } else if (kind == 2 || kind == 3) {
        const double Qk = fma(4.0, gk2, gk1);
        const double* ph = s_par[s];
        const double u00 = gj0 * ph[0], u01 = gj1 * ph[1], u02 = gj2 * ph[2];
        const double u10 = gj0 * ph[3], u11 = gj1 * ph[4], u12 = gj2 * ph[5];
        const double u20 = gj0 * ph[6], u21 = gj1 * ph[7], u22 = gj2 * ph[8];
        const double K0 = (u00 + u01 + u02) * Sk;
        const double K1 = (u10 + u11 + u12) * Mk - fma(2.0, u12, u11) * Sk;
        const double K2 = (u20 + u21 + u22) * Qk - 2.0 * fma(2.0, u22, u21) * Mk + fma(4.0, u22, u21) * Sk;
        if (kind == 2) {
#pragma unroll
          for (int nn = 0; nn < NPL; ++nn) acc[nn] *= fma(fma(K2, c_gamma[nn], K1), c_gamma[nn], K0);
        } else {
          const double u30 = gj0 * ph[9], u31 = gj1 * ph[10], u32 = gj2 * ph[11];
          const double Tk = 3.0 * Qk - 2.0 * Mk;  // sum_m m^3 g_k[m]
          const double K3 = (u30 + u31 + u32) * Tk - 3.0 * fma(2.0, u32, u31) * Qk + 3.0 * fma(4.0, u32, u31) * Mk - fma(8.0, u32, u31) * Sk;
#pragma unroll
          for (int nn = 0; nn < NPL; ++nn) acc[nn] *= fma(fma(fma(K3, c_gamma[nn], K2), c_gamma[nn], K1), c_gamma[nn], K0);
        }
      } else {
        const double w00 = gj0 * gk0, w01 = gj0 * gk1, w02 = gj0 * gk2, w10 = gj1 * gk0, w11 = gj1 * gk1, w12 = gj1 * gk2,
                     w20 = gj2 * gk0, w21 = gj2 * gk1, w22 = gj2 * gk2;
#pragma unroll
        for (int nn = 0; nn < NPL; ++nn) {
          const double* t = &s_dir[s][9 * (nn < PSCL_MAX_ALPHA ? nn : 0)];
          acc[nn] *= w00 * t[0] + w01 * t[1] + w02 * t[2] + w10 * t[3] + w11 * t[4] + w12 * t[5] + w20 * t[6] + w21 * t[7] + w22 * t[8];
        }
      }
    }
    since += (int)n;
    if (since >= 16) {  // keep the running products inside the double range (terms >= ~1e-10 each)
      since = 0;
#if PLY_EXSMEM
#pragma unroll
      for (int nn = 0; nn < NPL; ++nn) { int x = 0; pscl_renorm(acc[nn], x); s_ex[nn][tid] += x; }
#else
#pragma unroll
      for (int nn = 0; nn < NPL; ++nn) pscl_renorm(acc[nn], ex[nn]);
#endif
    }
    pos = pos_next;
    __syncthreads();  // everyone is done with the staged group before it is overwritten
  }
  if (J < nv && K < nv) {
    double* out = a.partial + (size_t)(item - a.item_base) * ((size_t)nv * nv * na) + ((size_t)J * nv + K) * na;
#pragma unroll
    for (int nn = 0; nn < NPL; ++nn) {
      if (nn < na) {
#if PLY_EXSMEM
        int x = s_ex[nn][tid];
#else
        int x = ex[nn];
#endif
        pscl_renorm(acc[nn], x);
        out[nn] = pscl_prod_log(acc[nn], x);
      }
    }
  }
}

// class-ordered record stream for k_demux_poly (built lazily, once per pileup image)
static int ply_build_stream(pscl_ctx* ctx, pscl_plp* p) {
  if (p->ply_rec) return PSCL_OK;
  const int64_t P = p->P;
  const int32_t NI = p->n_items;
  if (P + 1 > INT32_MAX) return pscl_fail(ctx, PSCL_EINVAL, "k_demux_poly: a device pileup image holds < 2^31 pairs");
  uint2* rec_tmp = nullptr;
  unsigned long long *key = nullptr, *scan = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** d, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(d, bytes ? bytes : 16); };
  alloc((void**)&rec_tmp, sizeof(uint2) * P);
  alloc((void**)&key, sizeof(unsigned long long) * (P + 1));
  alloc((void**)&scan, sizeof(unsigned long long) * (P + 2));
  alloc((void**)&p->ply_rec, sizeof(uint2) * P);
  alloc((void**)&p->ply_rng, sizeof(uint4) * NI);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, key, scan, (int64_t)(P + 1), ctx->stream);
  alloc(&tmp, tmp_bytes);
  if (e == cudaSuccess) e = cudaMemsetAsync(key, 0, sizeof(unsigned long long) * (P + 1), ctx->stream);
  if (e == cudaSuccess && P > 0) {
    k_dmx_classify<<<(unsigned)((P + 255) / 256), 256, 0, ctx->stream>>>(p->pair_snp, p->pair_rd, p->rd_aq, P, rec_tmp, key);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, key, scan, (int64_t)(P + 1), ctx->stream);
  if (e == cudaSuccess && NI > 0) {
    k_ply_scatter<<<(unsigned)(((int64_t)NI * 32 + 255) / 256), 256, 0, ctx->stream>>>(
        p->cell_ptr, p->item_cell, p->item_pbeg, p->item_pend, NI, rec_tmp, scan, p->ply_rec, p->ply_rng);
    e = cudaGetLastError();
  }
  ctx->launches += 3;
  cudaFree(rec_tmp); cudaFree(key); cudaFree(scan); cudaFree(tmp);
  if (e != cudaSuccess) {
    cudaFree(p->ply_rec); cudaFree(p->ply_rng);
    p->ply_rec = nullptr; p->ply_rng = nullptr;
    return pscl_fail(ctx, e == cudaErrorMemoryAllocation ? PSCL_ENOMEM : PSCL_ECUDA, "k_demux_poly record stream build failed: %s", cudaGetErrorString(e));
  }
  return PSCL_OK;
}

template <int NPL>
static cudaError_t launch_poly_n(pscl_ctx* ctx, const PolyArgs& a, dim3 grid) {
  const size_t dyn = PLY_EXSMEM ? sizeof(int) * NPL * 256 : 0;
  static bool attr_set[64] = {false};
  if (dyn && !attr_set[ctx->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_demux_poly<NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    attr_set[ctx->device & 63] = true;
  }
  k_demux_poly<NPL><<<grid, 256, dyn, ctx->stream>>>(a);
  return cudaGetLastError();
}
static cudaError_t launch_poly(pscl_ctx* ctx, const PolyArgs& a, int n_work) {
  dim3 grid((unsigned)(a.tiles * a.tiles), (unsigned)n_work);
  if (a.na <= 4) return launch_poly_n<4>(ctx, a, grid);
  if (a.na <= 8) return launch_poly_n<8>(ctx, a, grid);
  if (a.na <= 16) return launch_poly_n<16>(ctx, a, grid);
  if (a.na <= 21) return launch_poly_n<21>(ctx, a, grid);
  return launch_poly_n<32>(ctx, a, grid);
}
