// context.inl — context lifetime and pileup upload (part of the single TU popscle_b200.cu)

#define PSCL_ITEM_PAIRS 2048

extern "C" int pscl_abi_version(void) { return PSCL_ABI_VERSION; }

extern "C" int pscl_create(int device, pscl_ctx** out, char* err, size_t errlen) {
  auto fail = [&](int code, const std::string& m) {
    if (err && errlen) snprintf(err, errlen, "%s", m.c_str());
    return code;
  };
  if (!out) return fail(PSCL_EINVAL, "pscl_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return fail(PSCL_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                 " (popscle_b200 has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(PSCL_ENODEV, "device ordinal out of range");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(PSCL_ECUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(PSCL_ENODEV, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                 "; this library carries sm_100a code only");
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(PSCL_ECUDA, cudaGetErrorString(e));
  pscl_ctx* c = new pscl_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  {  // bound of the staged kernel's wait for a slice, in SM clock ticks (prop.clockRate is in kHz = ticks per ms)
    const char* tv = getenv("PSCL_STAGE_TIMEOUT_MS");
    const long long ms = tv ? atoll(tv) : 2000;
    c->stage_spin_ticks = (ms > 0 ? ms : 1) * (long long)(prop.clockRate > 0 ? prop.clockRate : 1900000);
  }
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    delete c;
    return fail(PSCL_ECUDA, cudaGetErrorString(e));
  }
  cudaEventCreate(&c->ev0);
  cudaEventCreate(&c->ev1);
  cudaEventCreate(&c->ev2);
  cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c->stage_go, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->ev_counts, cudaEventDisableTiming);
  for (auto& ev : c->slice_ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  if (cudaHostAlloc((void**)&c->h_one, sizeof(int), cudaHostAllocDefault) == cudaSuccess) *c->h_one = 1;
  if (cudaHostAlloc((void**)&c->h_dict_over, sizeof(int), cudaHostAllocDefault) == cudaSuccess) *c->h_dict_over = 1;
  if (cudaHostAlloc((void**)&c->h_geno_bad, sizeof(int), cudaHostAllocDefault) == cudaSuccess) *c->h_geno_bad = 0;
  if (cudaHostAlloc((void**)&c->h_bad, sizeof(int), cudaHostAllocDefault) == cudaSuccess) *c->h_bad = 0;
  cudaEventCreateWithFlags(&c->ev_dict, cudaEventDisableTiming);
  {  // keep freed blocks mapped in the device's default pool (see common.cuh: device memory)
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  PsclScope scope__(c);
  // PhredHelper.cpp:30 — same expression, same libm, evaluated on the host
  double tab[256];
  for (int i = 0; i < 256; ++i) tab[i] = (i > 1) ? pow(0.1, i * 0.1) : 0.75;
  // per-read fold factors for the default alpha grid {0, 0.5}: cmd_cram_demuxlet.cpp:666-667,:673,:685
  static double ftab[3 * 64 * 6];
  for (int al = 0; al < 3; ++al)
    for (int q = 0; q < 64; ++q) {
      double* t = ftab + (al * 64 + q) * 6;
      double err = tab[q], mat = 1. - err;
      double pR = (al == 0) ? mat : err / 3.0, pA = (al == 1) ? mat : err / 3.0;
      for (int i = 0; i < 5; ++i) {
        double p = 0.25 * i;
        t[i] = (al == 2) ? 1.0 : (pR * (1.0 - p) + pA * p);
      }
      t[5] = 1.0;
    }
  if (cudaMalloc((void**)&c->fold_tab, sizeof(ftab)) != cudaSuccess ||
      cudaMemcpyAsync(c->fold_tab, ftab, sizeof(ftab), cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
      cudaMalloc((void**)&c->phred_err, sizeof(tab)) != cudaSuccess ||
      cudaMemcpyAsync(c->phred_err, tab, sizeof(tab), cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
      cudaMalloc((void**)&c->dm_counter, 64) != cudaSuccess || cudaMalloc((void**)&c->stage_flags, sizeof(int) * PSCL_MAX_STAGES) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) {
    delete c;
    return fail(PSCL_ENOMEM, "device allocation failed in pscl_create");
  }
  *out = c;
  return PSCL_OK;
}

static void fmx_state_free(pscl_ctx* ctx);

extern "C" void pscl_destroy(pscl_ctx* ctx) {
  if (!ctx) return;
  PsclScope scope__(ctx);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  fmx_state_free(ctx);
  cudaFree(ctx->gp);
  cudaFree(ctx->has_gp_buf);
  cudaFree(ctx->geno_raw); cudaFree(ctx->geno_err); cudaFree(ctx->geno_bad);
  cudaFree(ctx->gpM);
  cudaFree(ctx->gpS);
  cudaFree(ctx->dm_cells);
  cudaFree(ctx->dm_grid);
  cudaFree(ctx->dm_partial);
  cudaFree(ctx->dm_counter);
  cudaFree(ctx->phred_err);
  cudaFree(ctx->fold_tab);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaEventDestroy(ctx->ev2);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->stage_go) cudaEventDestroy(ctx->stage_go);
  if (ctx->ev_up) cudaEventDestroy(ctx->ev_up);
  if (ctx->ev_counts) cudaEventDestroy(ctx->ev_counts);
  for (auto& ev : ctx->slice_ev) if (ev) cudaEventDestroy(ev);
  if (ctx->h_one) cudaFreeHost(ctx->h_one);
  if (ctx->h_dict_over) cudaFreeHost(ctx->h_dict_over);
  if (ctx->h_geno_bad) cudaFreeHost(ctx->h_geno_bad);
  if (ctx->h_bad) cudaFreeHost(ctx->h_bad);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->ev_dict) cudaEventDestroy(ctx->ev_dict);
  cudaFree(ctx->gp_code); cudaFree(ctx->gp_cls); cudaFree(ctx->gp_dict); cudaFree(ctx->gp_dict_key); cudaFree(ctx->gp_dict_over);
  cudaFree(ctx->stage_flags);
  pscl_blk_cache_flush(ctx);  // the blocks this context kept for reuse go back to the pool
  cudaStreamSynchronize(ctx->stream);
  {  // hand the cached blocks back to the driver
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* pscl_last_error(const pscl_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" void* pscl_stream(pscl_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int pscl_set_stream(pscl_ctx* ctx, void* stream) {
  if (!ctx) return PSCL_EINVAL;
  PsclScope scope__(ctx);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)stream;
  ctx->own_stream = false;
  return PSCL_OK;
}
extern "C" int pscl_sync(pscl_ctx* ctx) {
  if (!ctx) return PSCL_EINVAL;
  PsclScope scope__(ctx);
  PSCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PSCL_OK;
}
extern "C" int64_t pscl_launch_count(const pscl_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int pscl_debug_fail_alloc(pscl_ctx* ctx, int nth) {
  if (!ctx) return PSCL_EINVAL;
  ctx->fail_alloc_in = nth > 0 ? nth : 0;
  return PSCL_OK;
}
extern "C" int pscl_set_partial_budget(pscl_ctx* ctx, size_t bytes) {
  if (!ctx || bytes < (1u << 20)) return PSCL_EINVAL;
  ctx->partial_budget_bytes = bytes;
  return PSCL_OK;
}

// SNP id as stored: inside [0, V) whatever the input says (a malformed id raises the pileup's flag where it is found)
__device__ __forceinline__ int pscl_clamp_snp(int id, int V) { return (unsigned)id < (unsigned)V ? id : 0; }

// allele (0/1/2) and phred quality (<= 63) of one base-call packed into one byte
__global__ void k_pack_reads(const uint8_t* __restrict__ al, const uint8_t* __restrict__ q,
                             uint8_t* __restrict__ aq, int64_t n, int* bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t a = al[i], b = q[i];
  if (a > 2 || b > 63) atomicExch(bad, 1);
  aq[i] = (uint8_t)((a << 6) | (b & 63));
}

// ABI 6: base-calls as read_bits-wide indices into a palette of allele<<6|qual bytes -> one byte per base-call
__global__ void k_unpack_reads(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ palette, int bits, int64_t r0, int64_t n, uint8_t* __restrict__ aq, int* bad) {
  __shared__ uint8_t s_pal[64];
  if (threadIdx.x < (1u << bits)) s_pal[threadIdx.x] = palette[threadIdx.x];
  __syncthreads();
  const int64_t r = r0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // base-calls [r0, n) of the bit string
  if (r >= n) return;
  const int64_t o = r * bits;
  const unsigned w = (unsigned)packed[o >> 3] | ((unsigned)packed[(o >> 3) + 1] << 8);  // one byte of slack behind the string
  const uint8_t v = s_pal[(w >> (int)(o & 7)) & ((1u << bits) - 1u)];
  if ((v >> 6) == 3) atomicExch(bad, 1);
  aq[r] = v;
}

// already packed reads: only the validity check (allele code 3 does not exist)
__global__ void k_check_reads(const uint8_t* __restrict__ aq, int64_t n, int* bad) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i >= n) return;
  bool b = false;
  if (i + 16 <= n && (reinterpret_cast<uintptr_t>(aq + i) & 15) == 0) {
    const uint4 v = *reinterpret_cast<const uint4*>(aq + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) b |= ((w[k] & (w[k] >> 1)) & 0x40404040u) != 0u;  // bits 7 and 6 both set
  } else {
    for (int64_t k = i; k < n && k < i + 16; ++k) b |= (aq[k] >> 6) == 3;
  }
  if (b) atomicExch(bad, 1);
}

// ABI 3: SNP ids from 16-bit gaps, one warp per cell.  A lane takes 8 consecutive gaps (one 16-byte load, the next
// batch of 256 already in flight), sums them locally, the warp scans the lane totals on top of the running id.  The
// first gap of a cell is ignored (its id is cell_first_snp); ids only grow, so every id is range-checked cheaply.
__global__ void k_decode_snp(const int64_t* __restrict__ cell_ptr, const int32_t* __restrict__ first, const uint16_t* __restrict__ delta,
                             int32_t c_begin, int32_t c_end, int32_t V, int32_t* __restrict__ pair_snp, int* bad) {
  const int c = c_begin + (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (c >= c_end) return;
  const int64_t b = cell_ptr[c], e = cell_ptr[c + 1];
  if (b >= e) return;
  int run = first[c];
  bool oob = run < 0 || run >= V;
  auto fetch = [&](int64_t base) {
    const int64_t p = base + lane * 8;
    return (p < e) ? *reinterpret_cast<const uint4*>(delta + p) : make_uint4(0u, 0u, 0u, 0u);  // delta has 16 B of slack
  };
  int64_t base = b & ~(int64_t)7;
  uint4 cur = fetch(base);
  for (; base < e; base += 256) {
    const uint4 nxt = fetch(base + 256);
    const int64_t p = base + lane * 8;
    const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
    int x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t q = p + i;
      const int d = (int)((w[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
      x[i] = (q > b && q < e) ? d : 0;
    }
#pragma unroll
    for (int i = 1; i < 8; ++i) x[i] += x[i - 1];
    int tot = x[7], inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    const int off = run + inc - tot;
    // (an id outside [0, V) raises the flag below; what is stored is clamped, because a pipelined run scores before it reads
    // the flag and the kernels index their tables with these ids)
    int y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = pscl_clamp_snp(off + x[i], V);
    if (p >= b && p + 8 <= e) {
      int4* dst = reinterpret_cast<int4*>(pair_snp + p);
      dst[0] = make_int4(y[0], y[1], y[2], y[3]);
      dst[1] = make_int4(y[4], y[5], y[6], y[7]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const int64_t q = p + i; if (q >= b && q < e) pair_snp[q] = y[i]; }
    }
    run = __shfl_sync(0xffffffffu, off + tot, 31);
    oob |= run < 0 || run >= V;  // ids are non-decreasing: the running maximum is the last one
    cur = nxt;
  }
  if (oob && lane == 0) atomicExch(bad, 2);
}
// ABI 6: SNP ids from 8-bit gaps with the large ones on the side.  One CTA per cell, tiles of 2048 gaps (8 per thread): a
// block scan of the marker counts places every 255 in gap_big (indexed from the cell's own first large gap, cell_gap_ptr),
// a second one of the resolved gaps gives the ids — a cell is a couple of microseconds of dependent work instead of a warp
// walking its 60 steps of 32 gaps one after the other, which made every launch last ~70 us however few cells it had (the
// pipelined pscl_demux_run decodes slice by slice).
#define PSCL_DEC8_NT 256
#define PSCL_DEC8_PER 8
__device__ __forceinline__ int dec8_block_exscan(int v, int* s_warp, int& total) {  // exclusive scan over the CTA's 256 threads
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  int before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < PSCL_DEC8_NT / 32; ++w) { const int t = s_warp[w]; before += w < warp ? t : 0; all += t; }
  __syncthreads();
  total = all;
  return before + x - v;
}
__global__ void __launch_bounds__(PSCL_DEC8_NT) k_decode_snp8(const int64_t* __restrict__ cell_ptr, const int32_t* __restrict__ first,
                              const uint8_t* __restrict__ delta8, const uint32_t* __restrict__ gap_big, const int64_t* __restrict__ cell_gap_ptr,
                              int64_t n_gap_big, int32_t C, int32_t V, int32_t* __restrict__ pair_snp, int* bad) {
  __shared__ int s_warp[PSCL_DEC8_NT / 32];
  __shared__ int s_oob;
  const int c = blockIdx.x, tid = threadIdx.x;
  if (c >= C) return;
  const int64_t b = cell_ptr[c], e = cell_ptr[c + 1];
  if (b >= e) return;
  if (tid == 0) s_oob = 0;
  int run = first[c];
  int64_t big = cell_gap_ptr[c];
  bool oob = run < 0 || run >= V;
  for (int64_t base = b; base < e; base += PSCL_DEC8_NT * PSCL_DEC8_PER) {
    const int64_t p0 = base + (int64_t)tid * PSCL_DEC8_PER;
    int d[PSCL_DEC8_PER];
    int nmark = 0;
#pragma unroll
    for (int i = 0; i < PSCL_DEC8_PER; ++i) {
      const int64_t p = p0 + i;
      d[i] = (p < e && p > b) ? (int)delta8[p] : 0;
      nmark += d[i] == 255;
    }
    int tot_mark;
    int64_t k = big + dec8_block_exscan(nmark, s_warp, tot_mark);
    int sum = 0;
#pragma unroll
    for (int i = 0; i < PSCL_DEC8_PER; ++i) {
      if (d[i] == 255) {
        if (k < n_gap_big) d[i] = (int)gap_big[k]; else { d[i] = 0; oob = true; }
        ++k;
      }
      sum += d[i];
    }
    big += tot_mark;
    int tot_sum;
    int id = run + dec8_block_exscan(sum, s_warp, tot_sum);
#pragma unroll
    for (int i = 0; i < PSCL_DEC8_PER; ++i) {
      const int64_t p = p0 + i;
      id += d[i];
      if (p < e) { pair_snp[p] = pscl_clamp_snp(id, V); oob |= id < 0 || id >= V; }
    }
    run += tot_sum;
  }
  if (big != cell_gap_ptr[c + 1]) oob = true;  // the cell used more or fewer large gaps than the host says it owns
  if (oob) s_oob = 1;
  __syncthreads();
  if (tid == 0 && s_oob) atomicExch(bad, 2);
}
// ABI 6 + 7, everything of a cell in one CTA: SNP ids from the 8-bit gaps, read offsets from the 2-bit counts (seeded with the
// cell's first base-call, cell_read_ptr), and the cell's base-calls unpacked from the palette bit string.  With the offsets
// of the cells given, nothing in the decoding reaches across cells: a slice of whole cells is ONE launch that lasts about one
// cell (~10 us), instead of an expand kernel, a device-wide scan and an unpack kernel over the slice.
// Per tile of 2048 pairs (8 per thread): one block scan of the two marker counts (gaps = 255 / counts = 0, packed 16 + 16
// bits) places the large values in their side lists, one block scan of the resolved (gap, count) pairs (packed 32 + 32 bits)
// gives ids and offsets.  The rank of the cell's first large count inside its 1024-pair block of nreads_big_ptr is counted
// from the 2-bit fields in front of it.
__device__ __forceinline__ unsigned long long dec_block_exscan64(unsigned long long v, unsigned long long* s_warp, unsigned long long& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  unsigned long long before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < PSCL_DEC8_NT / 32; ++w) { const unsigned long long t = s_warp[w]; before += w < warp ? t : 0; all += t; }
  __syncthreads();
  total = all;
  return before + x - v;
}
__global__ void __launch_bounds__(PSCL_DEC8_NT, 4) k_decode_cells(const int64_t* __restrict__ cell_ptr, const int32_t* __restrict__ first,
                               const uint8_t* __restrict__ delta8, const uint32_t* __restrict__ gap_big, const int64_t* __restrict__ cell_gap_ptr,
                               int64_t n_gap_big, const uint8_t* __restrict__ n2, const uint8_t* __restrict__ nbig,
                               const int64_t* __restrict__ nblk, int64_t n_big, const int64_t* __restrict__ cell_rd,
                               const uint8_t* __restrict__ rpk, const uint8_t* __restrict__ rpal, int bits, int64_t n_reads,
                               int32_t C, int32_t V, int32_t* __restrict__ pair_snp, uint32_t* __restrict__ pair_rd,
                               uint8_t* __restrict__ aq, int* bad) {
  __shared__ unsigned long long s_warp[PSCL_DEC8_NT / 32];
  __shared__ uint8_t s_pal[64];
  __shared__ int s_flag;
  const int c = blockIdx.x, tid = threadIdx.x;
  if (c >= C) return;
  const int64_t b = cell_ptr[c], e = cell_ptr[c + 1], rb = cell_rd[c], re = cell_rd[c + 1];
  if (tid == 0) s_flag = 0;
  if (tid < (1 << bits)) s_pal[tid] = rpal[tid];
  int flag = 0;  // 2 = SNP ids / large gaps, 3 = counts, 1 = base-calls
  if (rb < 0 || re < rb || re > n_reads) flag = 3;
  if (b < e && !flag) {
    // rank of the cell's first large count: the block's start in nreads_big + the zero fields between the block's first pair and b
    const int64_t blk0 = (b >> 10) << 10;
    unsigned long long zeros = 0, tot;
    for (int64_t g = blk0 + tid; g < b; g += PSCL_DEC8_NT) zeros += ((n2[g >> 2] >> (2 * (int)(g & 3))) & 3u) == 0u;
    dec_block_exscan64(zeros, s_warp, tot);
    int64_t big_cnt = nblk[b >> 10] - nblk[0] + (int64_t)tot;
    int64_t big_gap = cell_gap_ptr[c];
    int run_snp = first[c];
    int64_t run_rd = rb;
    if (run_snp < 0 || run_snp >= V) flag = 2;
    for (int64_t base = b; base < e; base += PSCL_DEC8_NT * PSCL_DEC8_PER) {
      const int64_t p0 = base + (int64_t)tid * PSCL_DEC8_PER;
      int d[PSCL_DEC8_PER], f[PSCL_DEC8_PER];
      unsigned marks = 0;  // low 16 bits: gaps of 255, high 16 bits: counts of 0
#pragma unroll
      for (int i = 0; i < PSCL_DEC8_PER; ++i) {
        const int64_t p = p0 + i;
        d[i] = (p < e && p > b) ? (int)delta8[p] : 0;
        f[i] = p < e ? (int)((n2[p >> 2] >> (2 * (int)(p & 3))) & 3u) : 1;
        marks += (d[i] == 255 ? 1u : 0u) + (f[i] == 0 ? 0x10000u : 0u);
        if (p >= e) f[i] = 0;  // past the cell: contributes no base-call (it was counted as "not a marker" above)
      }
      unsigned long long tm;
      const unsigned long long em = dec_block_exscan64(marks, s_warp, tm);
      int64_t kg = big_gap + (int64_t)(em & 0xffffu), kc = big_cnt + (int64_t)((em >> 16) & 0xffffu);
      unsigned long long sums = 0;  // low 32 bits: gap sum, high 32 bits: base-call count sum
#pragma unroll
      for (int i = 0; i < PSCL_DEC8_PER; ++i) {
        const int64_t p = p0 + i;
        if (d[i] == 255) { if (kg < n_gap_big) d[i] = (int)gap_big[kg]; else { d[i] = 0; flag = 2; } ++kg; }
        if (p < e && f[i] == 0) { if (kc < n_big) { f[i] = (int)nbig[kc]; if (f[i] == 0) flag = 3; } else flag = 3; ++kc; }
        sums += (unsigned long long)(unsigned)d[i] + ((unsigned long long)(unsigned)f[i] << 32);
      }
      big_gap += (int64_t)(tm & 0xffffu);
      big_cnt += (int64_t)((tm >> 16) & 0xffffu);
      unsigned long long ts;
      const unsigned long long es = dec_block_exscan64(sums, s_warp, ts);
      int id = run_snp + (int)(unsigned)(es & 0xffffffffu);
      int64_t rd = run_rd + (int64_t)(es >> 32);
#pragma unroll
      for (int i = 0; i < PSCL_DEC8_PER; ++i) {
        const int64_t p = p0 + i;
        id += d[i];
        if (p < e) { pair_snp[p] = pscl_clamp_snp(id, V); pair_rd[p] = (uint32_t)rd; if (id < 0 || id >= V) flag = 2; }
        rd += f[i];
      }
      run_snp += (int)(unsigned)(ts & 0xffffffffu);
      run_rd += (int64_t)(ts >> 32);
    }
    if (big_gap != cell_gap_ptr[c + 1]) flag = 2;  // the cell used more or fewer large gaps than the host says it owns
    if (run_rd != re) flag = 3;                    // its counts do not add up to its base-calls
  } else if (b >= e && re != rb && !flag) flag = 3;
  if (tid == 0) pair_rd[e] = (uint32_t)re;  // the next cell writes the same value; the last cell closes the array
  __syncthreads();  // (s_pal)
  // base-calls: 8 per thread and step (their 4-6 bits sit in at most 7 consecutive bytes), so a cell of a few thousand
  // base-calls is one or two steps of independent loads
  if (flag != 3) {
    for (int64_t r0 = rb + (int64_t)tid * 8; r0 < re; r0 += (int64_t)PSCL_DEC8_NT * 8) {
      const int64_t o = r0 * bits;
      const int64_t by = o >> 3;
      unsigned long long w = 0;
#pragma unroll
      for (int k = 0; k < 7; ++k) w |= (unsigned long long)rpk[by + k] << (8 * k);  // (the bit string has slack bytes behind it)
      w >>= (int)(o & 7);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (r0 + k < re) {
          const uint8_t v = s_pal[(unsigned)(w >> (k * bits)) & ((1u << bits) - 1u)];
          if ((v >> 6) == 3) flag = flag ? flag : 1;
          aq[r0 + k] = v;
        }
      }
    }
  }
  if (flag) atomicMax(&s_flag, flag);
  __syncthreads();
  if (tid == 0 && s_flag) atomicExch(bad, s_flag);
}

// ABI 6: base-call counts from two bits per pair with the counts >= 4 on the side; one warp per block of 1024 GLOBAL pair
// indices (blk_ptr[k] = large counts before pair 1024*k); the image holds pairs [pair_base, pair_base + P) of the host's
// arrays (pair_base != 0 for the barcode shards of pscl_multi_demux_run).  n2 starts at the byte of pair g_first,
// big at entry blk_ptr[0] of the host's array.
__global__ void k_expand_counts2(const uint8_t* __restrict__ n2, const uint8_t* __restrict__ big, const int64_t* __restrict__ blk_ptr,
                                 int64_t g_first, int64_t pair_base, int64_t P, int64_t n_big, uint8_t* __restrict__ cnt8, int* bad) {
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t g0 = g_first + gw * 1024, end = pair_base + P;
  if (g0 >= end) return;
  int64_t run = blk_ptr[gw] - blk_ptr[0];
  bool wrong = false;
  for (int step = 0; step < 32; ++step) {
    const int64_t g = g0 + step * 32 + lane;
    unsigned f = 1;
    if (g < end) f = (n2[(g - g_first) >> 2] >> (2 * (int)(g & 3))) & 3u;
    const unsigned m = __ballot_sync(0xffffffffu, f == 0u);
    if (f == 0u) {
      const int64_t k = run + __popc(m & ((1u << lane) - 1u));
      if (k < n_big) f = big[k]; else wrong = true;
      if (f == 0u) wrong = true;  // a pair has at least one base-call
    }
    run += __popc(m);
    if (g >= pair_base && g < end) cnt8[g - pair_base] = (uint8_t)f;
  }
  if (__any_sync(0xffffffffu, wrong) && lane == 0) atomicExch(bad, 3);
}

// ABI 3: read offsets from 8-bit counts are an exclusive scan (CUB); this checks that they end at n_reads
__global__ void k_check_total(const uint32_t* __restrict__ pair_rd, int64_t P, int64_t N, int* bad) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && (int64_t)pair_rd[P] != N) atomicExch(bad, 3);
}
struct PsclU8ToU32 {
  __host__ __device__ uint32_t operator()(uint8_t x) const { return x; }
};

// int64 read offsets -> uint32 (device images hold < 2^32 reads)
__global__ void k_narrow_ptr(const int64_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n, int64_t base) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)(in[i] - base);
}
__global__ void k_rebase_u32(uint32_t* __restrict__ v, int64_t n, uint32_t base) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] -= base;
}

// wide forms: SNP ids inside [0, V), read offsets non-decreasing and <= N (the delta forms are checked while decoding)
__global__ void k_check_pairs(int32_t* __restrict__ pair_snp, const uint32_t* __restrict__ pair_rd, int64_t P, int32_t V, int64_t N,
                              int check_snp, int* bad) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  if (check_snp && (unsigned)pair_snp[p] >= (unsigned)V) { atomicExch(bad, 2); pair_snp[p] = 0; }  // flagged; never used as an index
  const uint32_t a = pair_rd[p], b = pair_rd[p + 1];
  if (b < a || (int64_t)b > N) atomicExch(bad, 5);
}

static const char* pscl_bad_pileup_msg(int bad) {
  return bad == 2 ? "pair_snp / pair_snp_delta16 / pair_snp_delta8 holds a SNP id outside [0, n_snps) (or the large-gap list does not match its markers)"
       : bad == 5 ? "pair_read_ptr is not non-decreasing inside [0, n_reads]"
       : bad == 3 ? "pair_nreads8 / pair_nreads2 does not sum to n_reads (or the large-count list does not match its markers)"
                  : "read_allele must be 0/1/2 and read_qual <= 63 (dsc-pileup writes phred <= 40, cmd_cram_dsc_pileup.cpp:19-20)";
}

extern "C" void pscl_plp_free(pscl_ctx* ctx, pscl_plp* p) {
  if (!p) return;
  if (ctx) {
    t_pscl_stream = ctx->stream; t_pscl_ctx = ctx; cudaSetDevice(ctx->device);
    if ((p->n_stages || p->n_slices) && ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);  // slices still in flight write into d_delta
    cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(p->d_delta); cudaFree(p->d_first); cudaFree(p->d_bad); cudaFree(p->d_delta8); cudaFree(p->d_gap_big); cudaFree(p->d_cell_gap_ptr);
  cudaFree(p->sl_cnt); cudaFree(p->sl_n2); cudaFree(p->sl_nbig); cudaFree(p->sl_nblk); cudaFree(p->sl_cell_rd); cudaFree(p->sl_rpk); cudaFree(p->sl_rpal); cudaFree(p->sl_scan_tmp);
  cudaFree(p->cell_ptr); cudaFree(p->pair_snp); cudaFree(p->pair_rd); cudaFree(p->rd_aq);
  cudaFree(p->snp_af); cudaFree(p->item_cell); cudaFree(p->item_pbeg);
  cudaFree(p->item_pend); cudaFree(p->item_order); cudaFree(p->cell_item_ptr); cudaFree(p->snp_ptr);
  cudaFree(p->snp_pair); cudaFree(p->pair_cell); cudaFree(p->scratch_h2d);
  cudaFree(p->ply_rec); cudaFree(p->ply_rng);
  cudaFree(p->dmx_pkt); cudaFree(p->dmx_deep); cudaFree(p->dmx_desc_nat); cudaFree(p->dmx_desc_sorted);
  delete p;
}

// Host image -> device image.  `stages` > 1 (pscl_demux_run only, ABI-3 delta arrays) leaves the SNP gaps out of the
// synchronous part: they cross PCIe on ctx->copy_stream in `stages` slices of whole cells, one event per slice, while
// the caller scores the slices that have arrived (pscl_demux_run); the decode kernel then runs per slice.
// Work items of a device image from its (host copy of the) cell_ptr: chunks of <= PSCL_ITEM_PAIRS pairs of one cell, listed
// by descending size (counting sort; a staged image is ordered slice by slice, so that the kernel's first items are the
// ones whose gaps land first).  Enqueues the five item arrays on ctx->stream; the host vectors live in *p.
// Pinned staging of a run's small host-built arrays.  pscl_stage_begin sizes the buffer (false: too large or no pinned
// memory, the caller copies from its pageable vectors and drains); pscl_stage_copy places `bytes` in it and queues the
// copy.  The buffer is reused by the next upload of the context, which starts after the previous run's final drain.
static bool pscl_stage_begin(pscl_ctx* ctx, size_t need) {
  ctx->h_stage_used = 0;
  if (need > ((size_t)256 << 20)) return false;
  if (need > ctx->h_stage_cap) {
    if (ctx->h_stage) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->h_stage); ctx->h_stage = nullptr; ctx->h_stage_cap = 0; }
    const size_t cap = std::max<size_t>(need + need / 2, (size_t)1 << 20);
    if (cudaHostAlloc((void**)&ctx->h_stage, cap, cudaHostAllocDefault) != cudaSuccess) { ctx->h_stage = nullptr; cudaGetLastError(); return false; }
    ctx->h_stage_cap = cap;
  }
  return true;
}
static cudaError_t pscl_stage_copy(pscl_ctx* ctx, void* dst, const void* src, size_t bytes, bool staged) {
  if (!bytes) return cudaSuccess;
  if (staged && ctx->h_stage_used + bytes + 16 <= ctx->h_stage_cap) {
    char* at = ctx->h_stage + ctx->h_stage_used;
    memcpy(at, src, bytes);
    ctx->h_stage_used += (bytes + 15) & ~(size_t)15;
    src = at;
  }
  if (ctx->pend_on) { ctx->pend.push_back({dst, src, bytes}); return cudaSuccess; }
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
}
// an H2D copy of one of the caller's arrays: at once on the context's stream, or (a run's deferred upload) listed for the
// copy stream
static cudaError_t pscl_h2d(pscl_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!bytes) return cudaSuccess;
  if (ctx->pend_on) { ctx->pend.push_back({dst, src, bytes}); return cudaSuccess; }
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
}

static cudaError_t plp_make_items(pscl_ctx* ctx, pscl_plp* p, const int64_t* cell_ptr, bool staged) {
  const int32_t C = p->C;
  const int64_t P = p->P;
  p->h_cell_ptr.assign(cell_ptr, cell_ptr + C + 1);
  std::vector<int32_t>& item_cell = p->h_item_cell;
  std::vector<int64_t>&pbeg = p->h_item_pbeg, &pend = p->h_item_pend;
  item_cell.clear(); pbeg.clear(); pend.clear();
  item_cell.reserve((size_t)C + (size_t)(P / PSCL_ITEM_PAIRS) + 1);
  pbeg.reserve(item_cell.capacity()); pend.reserve(item_cell.capacity());
  p->h_cell_item_ptr.resize(C + 1);
  for (int32_t c = 0; c < C; ++c) {
    p->h_cell_item_ptr[c] = (int32_t)item_cell.size();
    int64_t b = cell_ptr[c], e2 = cell_ptr[c + 1], n = e2 - b;
    int64_t nch = (n + PSCL_ITEM_PAIRS - 1) / PSCL_ITEM_PAIRS;
    for (int64_t i = 0; i < nch; ++i) {  // equal split, multiples of 32 pairs
      int64_t s0 = b + ((n * i / nch) & ~(int64_t)31), t = (i + 1 == nch) ? e2 : b + ((n * (i + 1) / nch) & ~(int64_t)31);
      item_cell.push_back(c); pbeg.push_back(s0); pend.push_back(t);
    }
  }
  p->h_cell_item_ptr[C] = (int32_t)item_cell.size();
  p->n_items = (int32_t)item_cell.size();
  std::vector<int32_t>& order = p->h_item_order;
  order.resize(p->n_items);
  {
    const int NB = PSCL_ITEM_PAIRS + 64;
    std::vector<int32_t> head(NB + 1);
    const int nseg = p->n_stages > 1 ? p->n_stages : p->n_slices > 1 ? p->n_slices : 1;
    for (int k = 0; k < nseg; ++k) {
      const int32_t ib = nseg > 1 ? p->h_cell_item_ptr[p->stage_cell[k]] : 0;
      const int32_t ie = nseg > 1 ? p->h_cell_item_ptr[p->stage_cell[k + 1]] : p->n_items;
      std::fill(head.begin(), head.end(), 0);
      for (int32_t i = ib; i < ie; ++i) head[NB - 1 - (int)std::min<int64_t>(pend[i] - pbeg[i], NB - 1) + 1]++;
      for (int b2 = 0; b2 < NB; ++b2) head[b2 + 1] += head[b2];
      for (int32_t i = ib; i < ie; ++i) order[ib + head[NB - 1 - (int)std::min<int64_t>(pend[i] - pbeg[i], NB - 1)]++] = i;
    }
  }
  cudaError_t e = cudaSuccess;
  auto up = [&](void** d, const void* src, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(d, bytes ? bytes : 16);
    if (e == cudaSuccess && bytes) e = pscl_stage_copy(ctx, *d, src, bytes, staged);
  };
  up((void**)&p->item_cell, item_cell.data(), sizeof(int32_t) * p->n_items);
  up((void**)&p->item_pbeg, pbeg.data(), sizeof(int64_t) * p->n_items);
  up((void**)&p->item_pend, pend.data(), sizeof(int64_t) * p->n_items);
  up((void**)&p->item_order, order.data(), sizeof(int32_t) * p->n_items);
  up((void**)&p->cell_item_ptr, p->h_cell_item_ptr.data(), sizeof(int32_t) * (C + 1));
  return e;
}

// read_base: the host's read offsets (pair_read_ptr / pair_read_ptr32) and read arrays belong to a longer pileup and
// this image starts at base-call `read_base` of it (barcode shards of pscl_multi_demux_run point into the caller's arrays).
// pair_base: the same for the ABI-6 count arrays (pair_nreads2 / nreads_big / nreads_big_ptr are indexed by the caller's
// global pair numbers; every other array of a shard view is already offset).
// deferred: the caller (pscl_demux_run) reads the validity flag itself at the end of its run and drains there, so nothing
// here waits for the device (the small host-built arrays go through the context's pinned staging buffer).
static int plp_upload_impl(pscl_ctx* ctx, const pscl_pileup* h, pscl_plp** out, int stages, int64_t read_base = 0, int64_t pair_base = 0, bool deferred = false,
                           const pscl_geno* geno_hook = nullptr, int slices = 1) {
  if (!h || !out) return pscl_fail(ctx, PSCL_EINVAL, "pscl_plp_upload: NULL argument");
  *out = nullptr;
  static const bool trace = getenv("PSCL_TRACE") != nullptr;  // wall-clock of the upload's phases on stderr
  auto tnow = [&](bool drain) { if (trace && drain) cudaStreamSynchronize(ctx->stream); return std::chrono::steady_clock::now(); };
  auto tms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto tr0 = tnow(false);
  const int32_t C = h->n_cells, V = h->n_snps;
  const int64_t P = h->n_pairs, N = h->n_reads;
  if (C < 0 || V < 0 || P < 0 || N < 0) return pscl_fail(ctx, PSCL_EINVAL, "negative size in pscl_pileup");
  const bool pal = h->read_packed != nullptr && h->read_palette != nullptr && h->read_bits >= 4 && h->read_bits <= 6;  // ABI 6
  const bool ptr32 = h->pair_read_ptr32 != nullptr, packed = pal || h->read_aq != nullptr;
  if (pal && read_base != 0) return pscl_fail(ctx, PSCL_EINVAL, "a shard view cannot point into read_packed (bit string): pass read_aq");
  // ABI 6 forms win over ABI 3 ones when both are given
  const bool dsnp8 = h->pair_snp_delta8 != nullptr && h->cell_first_snp != nullptr && h->cell_gap_big_ptr != nullptr && (h->n_gap_big == 0 || h->snp_gap_big);
  const bool cnt2 = h->pair_nreads2 != nullptr && h->nreads_big_ptr != nullptr && (h->n_nreads_big == 0 || h->nreads_big);
  const bool dsnp = !dsnp8 && h->pair_snp_delta16 != nullptr && h->cell_first_snp != nullptr;
  const bool cnt8 = cnt2 || h->pair_nreads8 != nullptr;  // cnt2 is expanded to 8-bit counts on the device first
  if (!h->cell_ptr || (P > 0 && ((!h->pair_snp && !dsnp && !dsnp8) || (!h->pair_read_ptr && !ptr32 && !cnt8))) ||
      (N > 0 && !packed && (!h->read_allele || !h->read_qual)))
    return pscl_fail(ctx, PSCL_EINVAL, "pscl_pileup has a NULL array");
  if (N >= ((int64_t)1 << 32) || P >= ((int64_t)1 << 32))
    return pscl_fail(ctx, PSCL_EINVAL, "a device pileup image holds < 2^32 pairs/reads; shard the barcodes or SNPs");
  if (h->cell_ptr[0] != 0 || h->cell_ptr[C] != P)
    return pscl_fail(ctx, PSCL_EINVAL, "cell_ptr must run from 0 to n_pairs");
  for (int32_t c = 0; c < C; ++c)
    if (h->cell_ptr[c + 1] < h->cell_ptr[c]) return pscl_fail(ctx, PSCL_EINVAL, "cell_ptr not monotone at cell %d", c);
  if (P > 0 && !cnt8 && (ptr32 ? ((int64_t)h->pair_read_ptr32[0] != read_base || (int64_t)h->pair_read_ptr32[P] != read_base + N)
                               : (h->pair_read_ptr[0] != read_base || h->pair_read_ptr[P] != read_base + N)))
    return pscl_fail(ctx, PSCL_EINVAL, "pair_read_ptr must run from 0 to n_reads");
  PSCL_CUDA(ctx, cudaSetDevice(ctx->device));
  if ((!dsnp && !dsnp8) || P == 0 || C < 2 || !ctx->copy_stream || !ctx->h_one) stages = 1;
  if (stages > PSCL_MAX_STAGES) stages = PSCL_MAX_STAGES;
  pscl_plp* p = new pscl_plp();
  p->C = C; p->V = V; p->P = P; p->N = N;
  // staging need: work items (<= C + P / PSCL_ITEM_PAIRS + 1 of them, 24 B each) + cell_item_ptr + rebased gap offsets
  if (deferred) deferred = ctx->copy_stream && ctx->ev_up && pscl_stage_begin(ctx, ((size_t)C + (size_t)(P / PSCL_ITEM_PAIRS) + 2) * 24 + ((size_t)C + 1) * 20 + 4096);
  // A deferred upload lists its H2D copies instead of queueing them: once every destination exists they all go to the copy
  // stream in one run (small arrays, counts, base-calls, then the gap slices), so the PCIe link is busy from the first
  // microsecond while the context's stream builds the genotype tables and then decodes what has landed.
  ctx->pend.clear();
  ctx->pend_on = deferred;
  struct PendOff { pscl_ctx* c; ~PendOff() { c->pend_on = false; } } pend_off__{ctx};

  // ---- 1. the big arrays first: their copies run while the host builds the work items below -------
  auto up = [&](void** d, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMalloc(d, bytes ? bytes : 16);
    if (e != cudaSuccess) return e;
    if (bytes) e = pscl_h2d(ctx, *d, src, bytes);
    return e;
  };
  cudaError_t e = cudaSuccess;
  const char* where = "";  // the first step that failed (named in the error message)
  int64_t n_gap_local = 0;
#define UP(field, src, bytes) do { if (e == cudaSuccess) { e = up((void**)&p->field, src, bytes); if (e != cudaSuccess) where = #field; } } while (0)
#define STEP(name) do { if (e != cudaSuccess && !*where) where = name; } while (0)
  UP(cell_ptr, h->cell_ptr, sizeof(int64_t) * (C + 1));
  uint8_t *d_al = nullptr, *d_q = nullptr, *d_cnt = nullptr, *d_n2 = nullptr, *d_nbig = nullptr, *d_rpk = nullptr, *d_rpal = nullptr;
  int64_t* d_nblk = nullptr;
  int64_t n_big_local = 0, n2_first = 0;
  void* d_scan_tmp = nullptr;
  // Fully sliced run (PSCL_SLICE_FULL=1; measured equal to slicing the gaps alone at configs[1], 1.72 against 1.71 ms per call:
  // four copies per slice instead of one cost the link 0.1 ms, and a slice's decoding cannot run under the previous group's
  // scoring, whose persistent CTAs hold every register file): the counts and base-calls are not copied whole ahead of the gaps but slice by slice with them, and
  // decoded by per-slice launches of the caller (pscl_demux_run).  Needs the ABI-6 forms and the host's read offsets at the
  // cuts (ABI 7's cell_read_ptr, or pair_read_ptr / pair_read_ptr32 when the host passes them beside the counts).
  const bool full = deferred && slices > 1 && cnt2 && pal && dsnp8 && (h->cell_read_ptr || h->pair_read_ptr || ptr32) && read_base == 0 && pair_base == 0 &&
                    P > 0 && C >= 2 * slices && ctx->copy_stream && ctx->ev_up && getenv("PSCL_SLICE_FULL");
  // Base-calls sliced with the gaps (PSCL_SLICE_READS=1; measured equal to slicing the gaps alone, 1.74 ms per call either way
  // under PSCL_TIMELINE: the first group starts 0.15 ms earlier, but every slice's unpacking then sits in the scoring stream's
  // chain instead of under the copies — profiles/r5e_slice_reads_timeline.txt): the counts still go whole and first
  // (5 MB at configs[1]; expanded and scanned under the other copies), then every slice brings its gaps AND its base-calls,
  // and the caller unpacks and decodes it with the ordinary kernels (k_unpack_reads on the slice's range, k_decode_snp8 on
  // its cells).  Scoring can start after the counts and two slices instead of after the counts and all the base-calls.
  const bool rsl = !full && deferred && slices > 1 && cnt8 && pal && (dsnp8 || dsnp) && (h->cell_read_ptr || h->pair_read_ptr || ptr32) &&
                   read_base == 0 && pair_base == 0 && P > 0 && C >= 2 * slices && ctx->copy_stream && ctx->ev_up && getenv("PSCL_SLICE_READS");
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_bad, sizeof(int));
  if (e == cudaSuccess) e = cudaMemsetAsync(p->d_bad, 0, sizeof(int), ctx->stream);
  if (cnt8) {  // ABI 3: 8-bit base-call counts, offsets by an exclusive scan
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_cnt, (size_t)P + 1);  // one zero byte of slack: the scan's last output is the total
    if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt + P, 0, 1, ctx->stream);
    if (cnt2 && P > 0) {  // ABI 6: two bits per pair + the large counts, expanded by k_expand_counts2
      const int64_t k0 = pair_base / 1024, k1 = (pair_base + P + 1023) / 1024, g_first = k0 * 1024;
      const int64_t byte0 = g_first / 4, byte1 = (pair_base + P + 3) / 4, big0 = h->nreads_big_ptr[k0], big1 = h->nreads_big_ptr[k1];
      if (big0 < 0 || big1 < big0 || big1 > h->n_nreads_big) { pscl_plp_free(ctx, p); return pscl_fail(ctx, PSCL_EINVAL, "nreads_big_ptr does not index nreads_big"); }
      n_big_local = big1 - big0;
      if (full) {  // buffers now, contents slice by slice (step 2b)
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_n2, (size_t)(byte1 - byte0) + 16);
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_nbig, (size_t)n_big_local + 16);
      } else {
        if (e == cudaSuccess) e = up((void**)&d_n2, h->pair_nreads2 + byte0, (size_t)(byte1 - byte0));
        if (e == cudaSuccess) e = up((void**)&d_nbig, h->nreads_big ? h->nreads_big + big0 : nullptr, (size_t)n_big_local);
      }
      if (e == cudaSuccess) e = up((void**)&d_nblk, h->nreads_big_ptr + k0, sizeof(int64_t) * (size_t)(k1 - k0 + 1));
      n2_first = g_first;
    } else if (e == cudaSuccess && P > 0) e = pscl_h2d(ctx, d_cnt, h->pair_nreads8, (size_t)P);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->pair_rd, sizeof(uint32_t) * (P + 1));
    STEP("pair_nreads8 / pair_rd");
    if (ctx->pend_on && !full) ctx->pend.push_back({nullptr, (const void*)1, 0});  // marker: the counts are complete up to here
  } else if (ptr32) { UP(pair_rd, h->pair_read_ptr32, sizeof(uint32_t) * (P + 1)); }
  else { UP(scratch_h2d, h->pair_read_ptr, sizeof(int64_t) * (P + 1)); }
  if (pal) {  // unpacked by k_unpack_reads
    if (full) {
      if (e == cudaSuccess) e = cudaMalloc((void**)&d_rpk, (size_t)((N * h->read_bits + 7) / 8 + 16));  // k_decode_cells reads 7 bytes at a time
      std::vector<int64_t> crd;  // the cells' first base-calls: given (ABI 7), or looked up in the host's offsets
      const int64_t* src = h->cell_read_ptr;
      if (!src) {
        crd.resize((size_t)C + 1);
        for (int32_t c = 0; c <= C; ++c) crd[(size_t)c] = ptr32 ? (int64_t)h->pair_read_ptr32[h->cell_ptr[c]] : h->pair_read_ptr[h->cell_ptr[c]];
        src = crd.data();
      }
      if (e == cudaSuccess) e = cudaMalloc((void**)&p->sl_cell_rd, sizeof(int64_t) * ((size_t)C + 1));
      if (e == cudaSuccess) e = pscl_stage_copy(ctx, p->sl_cell_rd, src, sizeof(int64_t) * ((size_t)C + 1), true);  // staged: `crd` is a local
    }
    else if (rsl) { if (e == cudaSuccess) e = cudaMalloc((void**)&d_rpk, (size_t)((N * h->read_bits + 7) / 8 + 16)); }  // filled slice by slice
    else if (e == cudaSuccess) e = up((void**)&d_rpk, h->read_packed, (size_t)((N * h->read_bits + 7) / 8 + 1));
    if (e == cudaSuccess) e = up((void**)&d_rpal, h->read_palette, (size_t)1 << h->read_bits);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->rd_aq, N ? (size_t)N : 16);
    STEP("read_packed");
  } else if (packed) {
    UP(rd_aq, h->read_aq, (size_t)N);
  } else {
    if (e == cudaSuccess) e = up((void**)&d_al, h->read_allele, (size_t)N);
    if (e == cudaSuccess) e = up((void**)&d_q, h->read_qual, (size_t)N);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->rd_aq, N ? (size_t)N : 16);
    STEP("read_allele / read_qual");
  }
  if (dsnp || dsnp8) {  // ABI 3 / 6: 16-bit or 8-bit SNP gaps, decoded per cell
    if (e == cudaSuccess) e = up((void**)&p->d_first, h->cell_first_snp, sizeof(int32_t) * C);
    if (dsnp) {
      if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_delta, sizeof(uint16_t) * (size_t)P + 16);  // k_decode_snp reads whole 16-byte groups
    } else {
      const int64_t g0 = h->cell_gap_big_ptr[0], g1 = h->cell_gap_big_ptr[C];
      if (g0 < 0 || g1 < g0 || g1 > h->n_gap_big) { pscl_plp_free(ctx, p); return pscl_fail(ctx, PSCL_EINVAL, "cell_gap_big_ptr does not index snp_gap_big"); }
      n_gap_local = g1 - g0;
      p->n_gap_big = n_gap_local;
      std::vector<int64_t>& cg = p->h_cell_gap_ptr;  // rebased to this image's first large gap
      cg.resize((size_t)C + 1);
      for (int32_t c = 0; c <= C; ++c) cg[c] = h->cell_gap_big_ptr[c] - g0;
      if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_delta8, (size_t)P + 16);
      if (e == cudaSuccess) e = up((void**)&p->d_gap_big, h->snp_gap_big ? h->snp_gap_big + g0 : nullptr, sizeof(uint32_t) * (size_t)n_gap_local);
      if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_cell_gap_ptr, sizeof(int64_t) * ((size_t)C + 1));
      if (e == cudaSuccess) e = pscl_stage_copy(ctx, p->d_cell_gap_ptr, cg.data(), sizeof(int64_t) * ((size_t)C + 1), deferred);
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->pair_snp, sizeof(int32_t) * (P ? P : 1));
    if (!deferred || !cnt8 || C < 2 * slices) slices = 1;  // (`full` implies none of these)
    if (slices > PSCL_MAX_STAGES) slices = PSCL_MAX_STAGES;
    if (stages > 1 || slices > 1) {  // slices of whole cells with about equal pair counts; their copies are queued in step 2b
      if (stages > 1) p->n_stages = stages; else { p->n_slices = slices; stages = slices; }
      p->stage_cell[0] = 0;
      for (int k = 1; k < stages; ++k) {
        const int64_t want = P * k / stages;
        int32_t c = (int32_t)(std::lower_bound(h->cell_ptr, h->cell_ptr + C + 1, want) - h->cell_ptr);
        p->stage_cell[k] = std::max(p->stage_cell[k - 1], std::min(c, C));
      }
      p->stage_cell[stages] = C;
    } else if (e == cudaSuccess && P > 0) {
      e = dsnp ? pscl_h2d(ctx, p->d_delta, h->pair_snp_delta16, sizeof(uint16_t) * P) : pscl_h2d(ctx, p->d_delta8, h->pair_snp_delta8, (size_t)P);
    }
  } else {
    UP(pair_snp, h->pair_snp, sizeof(int32_t) * P);
  }
  STEP("pair_snp_delta16 / pair_snp");
  if (h->snp_af) UP(snp_af, h->snp_af, sizeof(double) * V);

  // deferred upload: the copies listed so far go to the copy stream (once the context's stream has reached the point where
  // their destinations exist: stream-ordered allocation); called after the big arrays and again after the work items
  bool counts_marked = false;
  auto flush = [&]() {
    if (!deferred || e != cudaSuccess) return;
    e = cudaEventRecord(ctx->stage_go, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->stage_go, 0);
    for (size_t i = 0; i < ctx->pend.size() && e == cudaSuccess; ++i) {
      if (!ctx->pend[i].dst) { e = cudaEventRecord(ctx->ev_counts, ctx->copy_stream); counts_marked = true; continue; }
      e = cudaMemcpyAsync(ctx->pend[i].dst, ctx->pend[i].src, ctx->pend[i].bytes, cudaMemcpyHostToDevice, ctx->copy_stream);
    }
    ctx->pend.clear();
  };
  flush();  // the big arrays are on their way while the host builds the work items
  const auto tr1 = tnow(false);

  // ---- 2. work items (host; overlaps the copies) -------------------------------------------------------
  if (e == cudaSuccess) e = plp_make_items(ctx, p, h->cell_ptr, deferred);
  STEP("work items");
  const auto tr2 = tnow(false);
#undef UP
  // ---- 2b. deferred upload: the rest of the list; staged / sliced image: the gaps go last, slice by slice on the copy
  //          stream, a flag word (staged) or an event (sliced) behind each slice -----------------------------------------
  if (p->n_stages > 1 && e == cudaSuccess) e = cudaMemsetAsync(ctx->stage_flags, 0, sizeof(int) * PSCL_MAX_STAGES, ctx->stream);
  if (deferred) {
    flush();
    ctx->pend_on = false;
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_up, ctx->copy_stream);
    if (ctx->tl_on) cudaEventRecord(ctx->tl[1], ctx->copy_stream);
  } else if (p->n_stages > 1 && e == cudaSuccess) {
    e = cudaEventRecord(ctx->stage_go, ctx->stream);  // the buffers exist (stream-ordered allocation) and the flags are cleared
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->stage_go, 0);
  }
  if (p->n_stages > 1 || p->n_slices > 1) {
    const int ns = p->n_stages > 1 ? p->n_stages : p->n_slices;
    for (int k = 0; k < ns && e == cudaSuccess; ++k) {
      const int64_t pb = h->cell_ptr[p->stage_cell[k]], pe = h->cell_ptr[p->stage_cell[k + 1]];
      if (full) {  // the slice's counts (2-bit bytes from its first 1024-pair block on, its share of the large counts) and base-calls
        const int32_t ca = p->stage_cell[k], cb = p->stage_cell[k + 1];
        const int64_t rb = h->cell_read_ptr ? h->cell_read_ptr[ca] : ptr32 ? (int64_t)h->pair_read_ptr32[pb] : h->pair_read_ptr[pb];
        const int64_t re = h->cell_read_ptr ? h->cell_read_ptr[cb] : ptr32 ? (int64_t)h->pair_read_ptr32[pe] : h->pair_read_ptr[pe];
        p->sl_rb[k] = rb; p->sl_rb[k + 1] = re;
        if (rb < 0 || re < rb || re > N) { e = cudaErrorInvalidValue; where = "pair_read_ptr at a slice boundary"; break; }
        if (pe > pb) {
          const int64_t kb0 = pb / 1024, kb1 = (pe + 1023) / 1024, by0 = kb0 * 256, by1 = (pe + 3) / 4;
          const int64_t bg0 = h->nreads_big_ptr[kb0] - h->nreads_big_ptr[0], bg1 = h->nreads_big_ptr[kb1] - h->nreads_big_ptr[0];
          if (bg0 < 0 || bg1 < bg0 || bg1 > n_big_local) { e = cudaErrorInvalidValue; where = "nreads_big_ptr at a slice boundary"; break; }
          e = cudaMemcpyAsync(d_n2 + by0, h->pair_nreads2 + by0, (size_t)(by1 - by0), cudaMemcpyHostToDevice, ctx->copy_stream);
          if (e == cudaSuccess && bg1 > bg0) e = cudaMemcpyAsync(d_nbig + bg0, h->nreads_big + h->nreads_big_ptr[0] + bg0, (size_t)(bg1 - bg0), cudaMemcpyHostToDevice, ctx->copy_stream);
          const int64_t q0 = rb * h->read_bits / 8, q1 = (re * h->read_bits + 7) / 8 + 1;
          if (e == cudaSuccess && re > rb) e = cudaMemcpyAsync(d_rpk + q0, h->read_packed + q0, (size_t)(q1 - q0), cudaMemcpyHostToDevice, ctx->copy_stream);
        }
      }
      if (rsl) {  // the slice's base-calls (whole bytes of the bit string around its range)
        const int32_t ca = p->stage_cell[k], cb = p->stage_cell[k + 1];
        const int64_t rb = h->cell_read_ptr ? h->cell_read_ptr[ca] : ptr32 ? (int64_t)h->pair_read_ptr32[pb] : h->pair_read_ptr[pb];
        const int64_t re = h->cell_read_ptr ? h->cell_read_ptr[cb] : ptr32 ? (int64_t)h->pair_read_ptr32[pe] : h->pair_read_ptr[pe];
        p->sl_rb[k] = rb; p->sl_rb[k + 1] = re;
        if (rb < 0 || re < rb || re > N) { e = cudaErrorInvalidValue; where = "read offsets at a slice boundary"; break; }
        const int64_t q0 = rb * h->read_bits / 8, q1 = (re * h->read_bits + 7) / 8 + 1;
        if (re > rb) e = cudaMemcpyAsync(d_rpk + q0, h->read_packed + q0, (size_t)(q1 - q0), cudaMemcpyHostToDevice, ctx->copy_stream);
      }
      if (pe > pb && e == cudaSuccess)
        e = dsnp ? cudaMemcpyAsync(p->d_delta + pb, h->pair_snp_delta16 + pb, sizeof(uint16_t) * (size_t)(pe - pb), cudaMemcpyHostToDevice, ctx->copy_stream)
                 : cudaMemcpyAsync(p->d_delta8 + pb, h->pair_snp_delta8 + pb, (size_t)(pe - pb), cudaMemcpyHostToDevice, ctx->copy_stream);
      if (p->n_slices > 1) {  // pipelined run: the host queues the slice's decoding and scoring behind this event
        if (e == cudaSuccess) e = cudaEventRecord(ctx->slice_ev[k], ctx->copy_stream);
        continue;
      }
      // PSCL_FAULT=drop_stage_flag (fault-injection test): the last slice's flag never arrives, the kernel must time out
      const bool drop = k == p->n_stages - 1 && getenv("PSCL_FAULT") && !strcmp(getenv("PSCL_FAULT"), "drop_stage_flag");
      if (e == cudaSuccess && !drop) e = cudaMemcpyAsync(ctx->stage_flags + k, ctx->h_one, sizeof(int), cudaMemcpyHostToDevice, ctx->copy_stream);
    }
  }
  STEP("copy queue");
  if (ctx->tl_on) cudaEventRecord(ctx->tl[2], ctx->copy_stream);
  int hook_rc = PSCL_OK;
  if (geno_hook && e == cudaSuccess) hook_rc = pscl_demux_set_geno(ctx, geno_hook, V);  // on the context's stream, under the copies
  if (ctx->tl_on) cudaEventRecord(ctx->tl[3], ctx->stream);
  if (deferred && e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, counts_marked ? ctx->ev_counts : ctx->ev_up, 0);
  if (hook_rc != PSCL_OK) {
    cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->stream);
    const std::string msg = ctx->err;
    cudaFree(d_al); cudaFree(d_q); cudaFree(d_cnt); cudaFree(d_n2); cudaFree(d_nbig); cudaFree(d_nblk); cudaFree(d_rpk); cudaFree(d_rpal);
    pscl_plp_free(ctx, p);
    ctx->err = msg;
    return hook_rc;
  }

  // ---- 3. device-side decoding and checks ------------------------------------------------------------
  if ((dsnp || dsnp8) && p->n_stages == 0 && p->n_slices <= 1 && e == cudaSuccess && C > 0 && P > 0) {
    if (deferred && counts_marked) { e = cudaStreamWaitEvent(ctx->stream, ctx->ev_up, 0); counts_marked = false; }
    if (dsnp) k_decode_snp<<<(unsigned)(((int64_t)C * 32 + 255) / 256), 256, 0, ctx->stream>>>(p->cell_ptr, p->d_first, p->d_delta, 0, C, V, p->pair_snp, p->d_bad);
    else k_decode_snp8<<<(unsigned)C, PSCL_DEC8_NT, 0, ctx->stream>>>(p->cell_ptr, p->d_first, p->d_delta8, p->d_gap_big, p->d_cell_gap_ptr,
                                                                                       n_gap_local, C, V, p->pair_snp, p->d_bad);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (full && e == cudaSuccess) {
    // the caller decodes slice by slice (k_decode_cells): hand it the packed inputs
    p->sl_full = true;
    p->sl_cnt = d_cnt; p->sl_n2 = d_n2; p->sl_nbig = d_nbig; p->sl_nblk = d_nblk; p->sl_rpk = d_rpk; p->sl_rpal = d_rpal;
    p->sl_n_big = n_big_local; p->sl_read_bits = h->read_bits;
    p->sl_nbp.assign(h->nreads_big_ptr, h->nreads_big_ptr + (P + 1023) / 1024 + 1);
    d_cnt = d_n2 = d_nbig = d_rpk = d_rpal = nullptr; d_nblk = nullptr; d_scan_tmp = nullptr;  // owned by the image from here on
  }
  if (cnt2 && P > 0 && e == cudaSuccess && !p->sl_full) {
    const int64_t warps = (pair_base + P - n2_first + 1023) / 1024;
    k_expand_counts2<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, ctx->stream>>>(d_n2, d_nbig, d_nblk, n2_first, pair_base, P, n_big_local, d_cnt, p->d_bad);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (cnt8 && !p->sl_full) {
    size_t tb = 0;
    thrust::transform_iterator<PsclU8ToU32, const uint8_t*, uint32_t, uint32_t> it((const uint8_t*)d_cnt, PsclU8ToU32());
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tb, it, p->pair_rd, (int64_t)(P + 1), ctx->stream);
    if (e == cudaSuccess) e = cudaMalloc(&d_scan_tmp, tb ? tb : 16);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(d_scan_tmp, tb, it, p->pair_rd, (int64_t)(P + 1), ctx->stream);
    if (e == cudaSuccess) { k_check_total<<<1, 32, 0, ctx->stream>>>(p->pair_rd, P, N, p->d_bad); ctx->launches += 2; e = cudaGetLastError(); }
  } else if (cnt8) {
    // fully sliced: scanned per slice by the caller
  } else if (!ptr32) {
    if (deferred && counts_marked && e == cudaSuccess) { e = cudaStreamWaitEvent(ctx->stream, ctx->ev_up, 0); counts_marked = false; }
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->pair_rd, sizeof(uint32_t) * (P + 1));
    if (e == cudaSuccess) {
      k_narrow_ptr<<<(unsigned)((P + 1 + 255) / 256), 256, 0, ctx->stream>>>((const int64_t*)p->scratch_h2d, p->pair_rd, P + 1, read_base);
      ctx->launches++;
      e = cudaGetLastError();
    }
  } else if (read_base != 0 && e == cudaSuccess) {
    k_rebase_u32<<<(unsigned)((P + 1 + 255) / 256), 256, 0, ctx->stream>>>(p->pair_rd, P + 1, (uint32_t)read_base);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (deferred && counts_marked && e == cudaSuccess) { e = cudaStreamWaitEvent(ctx->stream, ctx->ev_up, 0); counts_marked = false; }  // the scan is queued
  if (e == cudaSuccess && P > 0 && ((!dsnp && !dsnp8) || !cnt8) && p->n_stages == 0 && p->n_slices <= 1) {
    k_check_pairs<<<(unsigned)((P + 255) / 256), 256, 0, ctx->stream>>>(p->pair_snp, p->pair_rd, P, V, N, (dsnp || dsnp8) ? 0 : 1, p->d_bad);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (rsl && e == cudaSuccess) {  // unpacked per slice by the caller
    p->sl_reads = true;
    p->sl_rpk = d_rpk; p->sl_rpal = d_rpal; p->sl_read_bits = h->read_bits;
    d_rpk = d_rpal = nullptr;  // owned by the image from here on
    // the host's offsets at the cuts must be the scanned counts' (a cut in the wrong place would unpack the wrong base-calls)
    for (int k = 1; k < p->n_slices && e == cudaSuccess; ++k) {
      k_check_total<<<1, 32, 0, ctx->stream>>>(p->pair_rd, h->cell_ptr[p->stage_cell[k]], p->sl_rb[k], p->d_bad);
      e = cudaGetLastError();
    }
  }
  if (p->sl_full || p->sl_reads) {
    // unpacked per slice by the caller
  } else if (e == cudaSuccess && N > 0 && pal) {
    k_unpack_reads<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_rpk, d_rpal, h->read_bits, 0, N, p->rd_aq, p->d_bad);
    ctx->launches++;
    e = cudaGetLastError();
  } else if (e == cudaSuccess && N > 0) {
    if (packed) k_check_reads<<<(unsigned)((N + 4095) / 4096), 256, 0, ctx->stream>>>(p->rd_aq, N, p->d_bad);
    else k_pack_reads<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_al, d_q, p->rd_aq, N, p->d_bad);
    ctx->launches++;
    e = cudaGetLastError();
  }
  STEP("decode / scan / checks (launch)");
  if (ctx->tl_on) cudaEventRecord(ctx->tl[4], ctx->stream);
  int bad = 0;
  if (!deferred) {
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, p->d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    // the host vectors above are pageable sources of async copies: drain before they go out of scope
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    STEP("drain (a copy or kernel of the upload failed on the device)");
  }
  cudaFree(d_al); cudaFree(d_q); cudaFree(d_cnt); cudaFree(d_scan_tmp); cudaFree(d_n2); cudaFree(d_nbig); cudaFree(d_nblk); cudaFree(d_rpk); cudaFree(d_rpal);
  if (p->n_stages == 0 && p->n_slices <= 1) {
    cudaFree(p->d_delta); cudaFree(p->d_first); cudaFree(p->d_delta8); cudaFree(p->d_gap_big); cudaFree(p->d_cell_gap_ptr);
    p->d_delta = nullptr; p->d_first = nullptr; p->d_delta8 = nullptr; p->d_gap_big = nullptr; p->d_cell_gap_ptr = nullptr;
  }
  if (trace) {
    const auto tr3 = tnow(false);
    fprintf(stderr, "[pscl_plp_upload] checks + enqueue %.3f ms | work items (host) %.3f | item arrays, decode, drain %.3f | stages %d\n", tms(tr0, tr1), tms(tr1, tr2), tms(tr2, tr3), p->n_stages);
  }
  if (e == cudaSuccess && bad) {
    pscl_plp_free(ctx, p);
    return pscl_fail(ctx, PSCL_EINVAL, "%s", pscl_bad_pileup_msg(bad));
  }
  if (e != cudaSuccess) {
    if (deferred) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->stream); }  // queued copies still read the caller's arrays
    pscl_plp_free(ctx, p);
    size_t mf = 0, mt = 0;
    cudaMemGetInfo(&mf, &mt);
    return pscl_fail(ctx, e == cudaErrorMemoryAllocation ? PSCL_ENOMEM : PSCL_ECUDA, "pileup upload failed at %s: %s (P=%lld N=%lld, device memory %zu of %zu MiB free)",
                     where, cudaGetErrorString(e), (long long)P, (long long)N, mf >> 20, mt >> 20);
  }
  cudaFree(p->scratch_h2d);
  p->scratch_h2d = nullptr;
  *out = p;
  return PSCL_OK;
}

extern "C" int pscl_plp_upload(pscl_ctx* ctx, const pscl_pileup* h, pscl_plp** out) {
  if (!ctx) return PSCL_EINVAL;
  PsclScope scope__(ctx);
  return plp_upload_impl(ctx, h, out, 1);
}
