// common.cuh — shared host/device plumbing of libpopscle_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <time.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <unordered_map>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

#include "../../include/popscle_b200.h"

#define PSCL_MAX_STAGES 16 /* slices of a staged pscl_demux_run (copy of slice k+1 under the scoring of slice k) */
#define PSCL_MAX_ALPHA 32 /* pG fold keeps n_alpha*9 values in one warp's registers (<= 9 per lane) */

struct pscl_plp {
  int32_t C = 0, V = 0;
  int64_t P = 0, N = 0;
  // cell-major CSR (device)
  int64_t* cell_ptr = nullptr;   // [C+1]
  int32_t* pair_snp = nullptr;   // [P]
  uint32_t* pair_rd = nullptr;   // [P+1] read offsets (N < 2^32 per device image)
  uint8_t* rd_aq = nullptr;      // [N] allele<<6 | qual (qual <= 63)
  double* snp_af = nullptr;      // [V] or null
  // work items: chunks of <= ITEM_PAIRS pairs of one cell, so that no single warp/CTA owns a
  // huge cell; items of a cell are contiguous, item_order lists them by descending size
  int32_t n_items = 0;
  int32_t* item_cell = nullptr;    // [n_items]
  int64_t* item_pbeg = nullptr;    // [n_items]
  int64_t* item_pend = nullptr;    // [n_items]
  int32_t* item_order = nullptr;   // [n_items]
  int32_t* cell_item_ptr = nullptr;  // [C+1]
  std::vector<int32_t> h_cell_item_ptr, h_item_cell, h_item_order;
  std::vector<int64_t> h_cell_ptr, h_item_pbeg, h_item_pend, h_cell_gap_ptr;
  // SNP-major view for the freemuxlet M-step (built lazily)
  int64_t* snp_ptr = nullptr;    // [V+1]
  uint32_t* snp_pair = nullptr;  // [P] pair ids, ascending cell id inside one SNP
  int32_t* pair_cell = nullptr;  // [P]
  void* scratch_h2d = nullptr;   // int64 staging for pair_read_ptr
  // class-ordered record stream of k_demux_poly (built lazily, demux_poly.inl)
  uint2* ply_rec = nullptr;      // [P]
  uint4* ply_rng = nullptr;      // [n_items] {begin, first M, first D, end}
  // demuxlet class streams for k_demux_cls (built lazily, demux_cls.inl): 8-byte records, class S
  // (<= 1 usable base-call) first, then M (2-3), then D (> 3), inside each cell's pair range
  unsigned char* dmx_pkt = nullptr;  // [n_pkt][272] batch packets: 16-byte header + 32 records (k_dmx_pack)
  double* dmx_deep = nullptr;        // [n_deep][6] folded factors of the pairs with > 3 usable base-calls
  uint4* dmx_desc_nat = nullptr;     // [n_items] {first packet, end packet, item, -}, natural item order
  uint4* dmx_desc_sorted = nullptr;  // [n_items] the same in item_order
  // ABI-3 delta form: first SNP per cell and 16-bit gaps (kept only while a staged pscl_demux_run decodes them slice
  // by slice), the upload's validity flag, and the slices themselves (cells [stage_cell[k], stage_cell[k+1]))
  int32_t* d_first = nullptr;
  uint16_t* d_delta = nullptr;
  // ABI-6 form of the same: 8-bit gaps (255 = the next entry of d_gap_big), first big gap of every cell
  uint8_t* d_delta8 = nullptr;
  uint32_t* d_gap_big = nullptr;
  int64_t* d_cell_gap_ptr = nullptr;
  int64_t n_gap_big = 0;
  int* d_bad = nullptr;
  int n_stages = 0;
  // fully sliced run: counts, base-calls and gaps of a slice land together and are decoded by per-slice launches, so these
  // upload temporaries live as long as the image (sl_full); sl_rb[k] = first base-call of slice k (from the host's offsets)
  bool sl_full = false;
  bool sl_reads = false;  // only the base-calls are sliced with the gaps (counts whole and early): unpacked per slice by k_unpack_reads
  uint8_t *sl_cnt = nullptr, *sl_n2 = nullptr, *sl_nbig = nullptr, *sl_rpk = nullptr, *sl_rpal = nullptr;
  int64_t* sl_nblk = nullptr;
  int64_t* sl_cell_rd = nullptr;  // [C+1] device copy of cell_read_ptr
  void* sl_scan_tmp = nullptr;
  size_t sl_scan_bytes = 0;
  int64_t sl_rb[PSCL_MAX_STAGES + 1] = {0};
  std::vector<int64_t> sl_nbp;  // host copy of nreads_big_ptr (per 1024 pairs)
  int64_t sl_n_big = 0;
  int sl_read_bits = 0;
  int n_slices = 0;  // pipelined pscl_demux_run: the gaps land slice by slice (one event each) and every slice is decoded and
                     // scored by its own launches; stage_cell[] holds the cuts of either form
  int32_t stage_cell[PSCL_MAX_STAGES + 1] = {0};
};

struct pscl_fmx_state;

struct pscl_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  std::string err;
  int64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  cudaStream_t copy_stream = nullptr;  // H2D slices of a staged pscl_demux_run
  cudaEvent_t stage_go = nullptr;
  cudaEvent_t ev_up = nullptr;         // a run's non-sliced arrays have landed (recorded on copy_stream)
  cudaEvent_t ev_counts = nullptr;     // ... its base-call counts have (they come first: their scan runs under the other copies)
  cudaEvent_t slice_ev[PSCL_MAX_STAGES] = {nullptr};  // ... slice k of the gaps has
  // deferred H2D copies of a run: queued on copy_stream once every destination buffer exists (see plp_upload_impl)
  struct PendingCopy { void* dst; const void* src; size_t bytes; };
  std::vector<PendingCopy> pend;
  bool pend_on = false;
  // Blocks this context has freed, kept by exact size for its next allocation of that size (pscl_pool_alloc / _free below):
  // a run allocates the same three dozen sizes every call, and although cudaMallocAsync / cudaFreeAsync are cheap on average
  // (0.6 us), a third of the calls of a 1.7 ms run stalled 1-4 ms inside them (profiles/r3r_allocator_jitter.txt)
  std::unordered_map<void*, size_t> blk_live;
  std::multimap<size_t, void*> blk_free;
  size_t blk_free_bytes = 0;
  // PSCL_TIMELINE=1: device-time stamps of a run (ms after its first enqueue), printed by pscl_demux_run
  cudaEvent_t tl[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool tl_on = false;
  int* stage_flags = nullptr;          // device [PSCL_MAX_STAGES]: slice k has landed (written by the copy queue)
  int* h_one = nullptr;                // pinned host word holding 1, the source of those flag writes
  long long stage_spin_ticks = 1ll << 32;  // how long a warp of the staged kernel waits for a slice (PSCL_STAGE_TIMEOUT_MS, default 2000)
  double* fold_tab = nullptr;   // [3][64][6] per-read factors of the default alpha grid (demux.inl)
  double* phred_err = nullptr;  // [256] device copy of PhredHelper's phred2Err (staged to smem by kernels)
  // demuxlet state
  int32_t nv = 0, geno_V = 0;
  double* gp = nullptr;       // [V][nv][3]
  uint8_t* has_gp = nullptr;  // [V] or null (= all); points at has_gp_buf
  uint8_t* has_gp_buf = nullptr;
  size_t gp_cap = 0, has_gp_cap = 0, gp_code_cap = 0, gp_cls_cap = 0;  // set_geno keeps its buffers from call to call
  uint8_t* geno_raw = nullptr; double* geno_err = nullptr; int* geno_bad = nullptr;  // ABI 4: the raw form before mixing
  size_t geno_raw_cap = 0, geno_err_cap = 0, geno_bad_cap = 0;
  double* gpM = nullptr;      // [V][(3nv+1)&~1] 16-B padded genotype rows (k_demux_cls, built lazily)
  double* gpS = nullptr;      // [V][nv][2] (S_j, M_j) moments of the rows
  // dictionary-coded genotypes (demux.inl, built by pscl_demux_set_geno when nv <= 8): usable when *h_dict_over == 0
  unsigned long long* gp_code = nullptr;   // [V] 8-bit code per sample
  unsigned long long* gp_cls = nullptr;    // [V] the SNP's <= 3 distinct codes + 2-bit class per sample (usable when *h_dict_over == 0)
  double* gp_dict = nullptr;               // [256][3] the distinct triples
  unsigned long long* gp_dict_key = nullptr;  // [256] hash keys claiming the slots
  int* gp_dict_over = nullptr;             // device flag: bit 0 more than 256 distinct triples (or a hash clash), bit 1 a SNP with a fourth triple
  int* h_dict_over = nullptr;              // pinned host copy of the flag
  int* h_geno_bad = nullptr;               // pinned: the raw genotype input (ABI 4) held an invalid hard-call code
  int* h_bad = nullptr;                    // pinned: the pileup image's validity flag, read back at the end of a run
  // pinned staging for the small host-built arrays of a run (work items, rebased offsets): copies out of pageable vectors
  // would stall the host behind the big copies already queued on the stream
  char* h_stage = nullptr;
  size_t h_stage_cap = 0, h_stage_used = 0;
  cudaEvent_t ev_dict = nullptr;           // the host copy is valid once this has completed
  bool dict_built = false;
  int dm_last_kernel = 0;                  // what the last pscl_demux_score launched (pscl_demux_select_kernel's numbering)
  int demux_kernel = 0;       // 0 auto, 1 k_demux_default (rows), 2 k_demux_general, 3 k_demux_cls, 4 k_demux_poly, 5 ab, 6 default (dictionary)
  bool keep_grid = false, force_general = false, dm_single_batch = true;
  int32_t dm_cell_begin = 0, dm_cell_end = 0, dm_nalpha = 0;
  void* dm_cells = nullptr;   // pscl_demux_cell[cells]
  size_t dm_cells_cap = 0;
  double* dm_grid = nullptr;  // [cells][nv][nv][nalpha] when keep_grid
  size_t dm_grid_cap = 0;
  double* dm_partial = nullptr;  // [items in batch][nv*nv*nalpha]
  size_t dm_partial_cap = 0;
  int* dm_counter = nullptr;     // persistent-kernel work counter
  size_t partial_budget_bytes = (size_t)1 << 30;
  float dm_ms_main = 0.f, dm_ms_total = 0.f;
  bool dm_timed = false;
  pscl_fmx_state* fmx = nullptr;
  int fail_alloc_in = 0;  // pscl_debug_fail_alloc
};

// ---- device memory: stream-ordered allocation out of the device's default memory pool -------------
// pscl_demux_run / pscl_fmx_run build and drop a device image per call; with plain cudaMalloc/cudaFree
// that cost 15-750 ms per call on the GPU box (traced with PSCL_TRACE=1), an order of magnitude more
// than the copies and kernels together.  Every allocation of this library therefore goes through
// cudaMallocAsync / cudaFreeAsync on the calling context's stream; pscl_create raises the pool's
// release threshold so freed blocks stay mapped for the next call.  The two macros below route the
// (many) existing call sites; PsclScope publishes the current context's stream to them.
static thread_local cudaStream_t t_pscl_stream = nullptr;
static thread_local pscl_ctx* t_pscl_ctx = nullptr;
static thread_local double t_pscl_alloc_ms = 0.0;  // PSCL_TIMELINE: host time inside the allocator calls of the current run
static thread_local int t_pscl_alloc_n = 0;
static const size_t PSCL_BLK_CACHE_BYTES = (size_t)16 << 30;  // what a context keeps at most (the pool would keep it mapped anyway)
static const size_t PSCL_BLK_CACHE_N = 512;
static inline cudaError_t pscl_raw_alloc(void** p, size_t n) {
  if (t_pscl_ctx && t_pscl_ctx->tl_on) {
    const auto t0 = std::chrono::steady_clock::now();
    const cudaError_t e = cudaMallocAsync(p, n, t_pscl_stream);
    t_pscl_alloc_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); ++t_pscl_alloc_n;
    return e;
  }
  return cudaMallocAsync(p, n, t_pscl_stream);
}
static inline cudaError_t pscl_raw_free(void* p) {
  if (t_pscl_ctx && t_pscl_ctx->tl_on) {
    const auto t0 = std::chrono::steady_clock::now();
    const cudaError_t e = cudaFreeAsync(p, t_pscl_stream);
    t_pscl_alloc_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); ++t_pscl_alloc_n;
    return e;
  }
  return cudaFreeAsync(p, t_pscl_stream);
}
static inline void pscl_blk_cache_flush(pscl_ctx* c) {  // hand every kept block back to the pool (out of memory, destroy)
  for (auto& kv : c->blk_free) pscl_raw_free(kv.second);
  c->blk_free.clear();
  c->blk_free_bytes = 0;
}
// Stream-ordered allocation with a per-context cache of freed blocks.  A kept block was "freed" in the order of the context's
// stream and is handed out again in that order, so work queued before the free is done before work queued after the reuse
// (copies a run puts on its copy stream are queued, and waited for, before the run frees anything).
static inline cudaError_t pscl_pool_alloc(void** p, size_t n) {
  pscl_ctx* const c = t_pscl_ctx;
  // fault injection (pscl_debug_fail_alloc): the n-th allocation from now fails the way an exhausted device does
  if (c && c->fail_alloc_in > 0 && --c->fail_alloc_in == 0) { *p = nullptr; return cudaErrorMemoryAllocation; }
  if (!n) n = 16;
  if (c) {
    auto it = c->blk_free.find(n);
    if (it != c->blk_free.end()) {
      *p = it->second;
      c->blk_free.erase(it);
      c->blk_free_bytes -= n;
      c->blk_live[*p] = n;
      return cudaSuccess;
    }
  }
  // a large request the kept blocks could cover between them: the workload has changed, so they go back to the pool first
  // (it then reuses their memory instead of mapping more)
  if (c && n >= ((size_t)64 << 20) && c->blk_free_bytes >= n) pscl_blk_cache_flush(c);
  cudaError_t e = pscl_raw_alloc(p, n);
  if (e == cudaErrorMemoryAllocation && c && !c->blk_free.empty()) {  // give the kept blocks back and try once more
    cudaGetLastError();
    pscl_blk_cache_flush(c);
    e = pscl_raw_alloc(p, n);
  }
  if (e == cudaSuccess && c) c->blk_live[*p] = n;
  return e;
}
static inline cudaError_t pscl_pool_free(void* p) {
  if (!p) return cudaSuccess;
  pscl_ctx* const c = t_pscl_ctx;
  if (c) {
    auto it = c->blk_live.find(p);
    if (it != c->blk_live.end()) {
      const size_t n = it->second;
      c->blk_live.erase(it);
      if (c->blk_free_bytes + n <= PSCL_BLK_CACHE_BYTES && c->blk_free.size() < PSCL_BLK_CACHE_N) {
        c->blk_free.emplace(n, p);
        c->blk_free_bytes += n;
        return cudaSuccess;
      }
    }
  }
  return pscl_raw_free(p);
}
#define cudaMalloc(p, n) pscl_pool_alloc((void**)(p), (n))
#define cudaFree(p) pscl_pool_free((void*)(p))
struct PsclScope {
  explicit PsclScope(const pscl_ctx* c) { t_pscl_stream = c->stream; t_pscl_ctx = const_cast<pscl_ctx*>(c); }
};

static inline int pscl_fail(pscl_ctx* ctx, int code, const char* fmt, ...) __attribute__((format(printf, 3, 4)));
#include <stdarg.h>
static inline int pscl_fail(pscl_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

#define PSCL_CUDA(ctx, call)                                                                        \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return pscl_fail(ctx, e__ == cudaErrorMemoryAllocation ? PSCL_ENOMEM : PSCL_ECUDA,            \
                       "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

template <typename T>
static inline int pscl_reserve(pscl_ctx* ctx, T** p, size_t* cap, size_t bytes) {
  if (*cap >= bytes && *p) return PSCL_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  PSCL_CUDA(ctx, cudaMalloc((void**)p, bytes ? bytes : 16));
  *cap = bytes;
  return PSCL_OK;
}

// ---- device helpers --------------------------------------------------------------------------
// phred -> P(error): PhredHelper.cpp:30 `phred2Err[i] = (i > 1) ? pow(0.1, i*0.1) : 0.75`, filled on
// the host with the same libm call so the table is bit-identical to the reference's.
// (single translation unit: popscle_b200.cu includes every part, so these are plain statics)
static __constant__ double c_alpha[PSCL_MAX_ALPHA];

// Running product kept as (mantissa in [1,2), integer exponent): one log per accumulator at the
// end instead of one per term (the reference calls log() per (j,k,n) term,
// cmd_cram_demuxlet.cpp:746).  renorm() moves the double's exponent field into `e`.
__device__ __forceinline__ void pscl_renorm(double& m, int& e) {
  int hi = __double2hiint(m);
  int ex = (hi >> 20) & 0x7ff;
  if (ex != 0 && ex != 0x7ff) {  // leave 0 / denormal / inf / nan untouched (-> log gives -inf/nan)
    e += ex - 1023;
    m = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(m));
  }
}
__device__ __forceinline__ double pscl_prod_log(double m, int e) {
  return log(m) + (double)e * 0.693147180559945309417232121458;
}

// Warp transpose-reduce of N (power of two) running products per lane: after the call lane L holds in
// m[0], x[0] the product over all 32 lanes of element (L * N) / 32.  Recursive halving: at every step a
// lane keeps one half of its elements, hands the other half to its partner and multiplies what it gets.
template <int N>
__device__ __forceinline__ void pscl_transpose_prod(double (&m)[N], int (&x)[N], const int lane) {
  int o = 16;
#pragma unroll
  for (int n = N / 2; n >= 1; n >>= 1, o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const double sm = up ? m[i] : m[i + n], km = up ? m[i + n] : m[i];
      const int sx = up ? x[i] : x[i + n], kx = up ? x[i + n] : x[i];
      m[i] = km * __shfl_xor_sync(0xffffffffu, sm, o);
      x[i] = kx + __shfl_xor_sync(0xffffffffu, sx, o);
    }
  }
  for (; o >= 1; o >>= 1) {  // fewer elements than lanes: finish with a plain butterfly
    m[0] *= __shfl_xor_sync(0xffffffffu, m[0], o);
    x[0] += __shfl_xor_sync(0xffffffffu, x[0], o);
  }
}

