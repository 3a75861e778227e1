// multi.inl — one process, N GPUs (part of the single TU popscle_b200.cu).  SURVEY.md §8(e):
//
//   demuxlet   every cell's likelihood grid depends only on its own reads and the read-only genotype table
//              (cmd_cram_demuxlet.cpp:636-1013 carries no cross-cell state): BARCODES are sharded into contiguous ranges
//              balanced by pair count, the genotype table is replicated (raw forms: one small H2D copy per GPU; the wide
//              FP64 table: one H2D copy and a binary tree of NVLink peer copies), no data-path collective.
//   freemuxlet llk[c][pair] is a sum over SNPs (cmd_cram_freemux2.cpp:454-455): SNPs are sharded into ranges balanced by
//              pair count; stage 1 and the greedy seeding need every SNP of a cell and run on GPU 0 over the whole pileup,
//              the initial clusters go to the others through host memory (400 KB at 100k cells); then per EM iteration
//              every GPU scores its SNP range, ONE all-reduce of the C x npairs partial sums, redundant classification,
//              SNP-local M-step.  The all-reduce is this library's own kernel over NVLink peer memory: GPU g sums slice
//              g of all N partial buffers in rank order (deterministic, the same bits on every GPU) and stores the sum
//              into every GPU's result buffer — a reduce-scatter and an all-gather fused in one launch per GPU.
//
// One host thread per GPU (bound to the CPUs of the GPU's NUMA node), the per-device C ABI underneath.  A device may be
// listed more than once (tests on a single-GPU box exercise the sharding and the collective that way).

#include <condition_variable>
#include <mutex>
#include <thread>
#include <thrust/iterator/counting_iterator.h>
#include <sched.h>

// ---- NUMA: run the calling thread on the CPUs next to a GPU (its pageable staging copies and pinned buffers then sit in
// that node's memory; the 8-GPU end-to-end curve of round 1 was bound by eight ranks sharing node 0) -------------------
extern "C" int pscl_bind_thread_to_device(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return PSCL_ENODEV;
  for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
  char path[128];
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
  FILE* f = fopen(path, "r");
  if (!f) return PSCL_EINVAL;
  char line[4096] = {0};
  const bool got = fgets(line, sizeof line, f) != nullptr;
  fclose(f);
  if (!got) return PSCL_EINVAL;
  cpu_set_t set;
  CPU_ZERO(&set);
  int n = 0;
  for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {  // "0-31,64-95"
    int a = 0, b = 0;
    const int k = sscanf(tok, "%d-%d", &a, &b);
    if (k == 1) b = a;
    if (k < 1) continue;
    for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, &set); ++n; }
  }
  if (n == 0) return PSCL_EINVAL;
  return sched_setaffinity(0, sizeof set, &set) == 0 ? PSCL_OK : PSCL_EINVAL;
}

// ---- a reusable host barrier that can be broken (a failed worker must not leave the others waiting) ------------------
struct MultiBarrier {
  std::mutex m;
  std::condition_variable cv;
  int n = 1, waiting = 0, gen = 0;
  bool broken = false;
  bool wait() {  // false = some worker failed
    std::unique_lock<std::mutex> lk(m);
    if (broken) return false;
    const int g = gen;
    if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); return true; }
    cv.wait(lk, [&] { return gen != g || broken; });
    return !broken;
  }
  void abort() { std::lock_guard<std::mutex> lk(m); broken = true; cv.notify_all(); }
  void reset(int n_) { n = n_; waiting = 0; broken = false; }
};

#define PSCL_MULTI_MAX 16

struct pscl_multi {
  int n = 0;
  int dev[PSCL_MULTI_MAX] = {0};
  pscl_ctx* ctx[PSCL_MULTI_MAX] = {nullptr};
  bool peer_ok = true;   // every pair of distinct devices can map the other's memory
  bool bind_numa = true;
  std::string err;
  MultiBarrier bar;
  // events of the collectives, one per GPU per phase
  cudaEvent_t ev_in[PSCL_MULTI_MAX] = {nullptr}, ev_out[PSCL_MULTI_MAX] = {nullptr}, ev_geno[PSCL_MULTI_MAX] = {nullptr};
  pscl_multi_timing tm{};
};

static int multi_fail(pscl_multi* m, int code, const std::string& msg) {
  if (m) m->err = msg;
  return code;
}

extern "C" int pscl_multi_create(const int* gpu_ids, int n_gpu, pscl_multi** out, char* err, size_t errlen) {
  auto fail = [&](int code, const std::string& msg) {
    if (err && errlen) snprintf(err, errlen, "%s", msg.c_str());
    return code;
  };
  if (!out) return fail(PSCL_EINVAL, "pscl_multi_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(PSCL_ENODEV, "no CUDA device (popscle_b200 has no CPU fallback)");
  if (n_gpu <= 0) n_gpu = ndev;  // all of them
  if (n_gpu > PSCL_MULTI_MAX) return fail(PSCL_EINVAL, "at most 16 GPUs");
  pscl_multi* m = new pscl_multi();
  m->n = n_gpu;
  m->bind_numa = getenv("PSCL_NO_NUMA_BIND") == nullptr;
  for (int i = 0; i < n_gpu; ++i) {
    m->dev[i] = gpu_ids ? gpu_ids[i] : i;
    char e2[512] = {0};
    const int rc = pscl_create(m->dev[i], &m->ctx[i], e2, sizeof e2);
    if (rc != PSCL_OK) {
      for (int k = 0; k < i; ++k) pscl_destroy(m->ctx[k]);
      delete m;
      return fail(rc, std::string("GPU ") + std::to_string(gpu_ids ? gpu_ids[i] : i) + ": " + e2);
    }
  }
  // peer access for plain allocations: the buffers of the all-reduce are cudaMalloc'ed outside the stream-ordered pool
  // (pool memory would need its own access list, cudaMemPoolSetAccess)
  for (int i = 0; i < n_gpu; ++i) {
    cudaSetDevice(m->dev[i]);
    for (int k = 0; k < n_gpu; ++k) {
      if (m->dev[k] == m->dev[i]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, m->dev[i], m->dev[k]);
      if (!can) { m->peer_ok = false; continue; }
      const cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[k], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) m->peer_ok = false;
      cudaGetLastError();
    }
    cudaEventCreateWithFlags(&m->ev_in[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&m->ev_out[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&m->ev_geno[i], cudaEventDisableTiming);
  }
  *out = m;
  return PSCL_OK;
}

extern "C" void pscl_multi_destroy(pscl_multi* m) {
  if (!m) return;
  for (int i = 0; i < m->n; ++i) {
    cudaSetDevice(m->dev[i]);
    if (m->ev_in[i]) cudaEventDestroy(m->ev_in[i]);
    if (m->ev_out[i]) cudaEventDestroy(m->ev_out[i]);
    if (m->ev_geno[i]) cudaEventDestroy(m->ev_geno[i]);
    pscl_destroy(m->ctx[i]);
  }
  delete m;
}
extern "C" const char* pscl_multi_last_error(const pscl_multi* m) { return m ? m->err.c_str() : "null pscl_multi"; }
extern "C" int pscl_multi_size(const pscl_multi* m) { return m ? m->n : 0; }
extern "C" pscl_ctx* pscl_multi_ctx(pscl_multi* m, int i) { return (m && i >= 0 && i < m->n) ? m->ctx[i] : nullptr; }
extern "C" int pscl_multi_last_timing(const pscl_multi* m, pscl_multi_timing* out) {
  if (!m || !out) return PSCL_EINVAL;
  *out = m->tm;
  return PSCL_OK;
}

// Runs fn(rank) on one host thread per GPU; the first failure (lowest rank) is reported.
template <typename F>
static int multi_run_workers(pscl_multi* m, F fn) {
  int rc[PSCL_MULTI_MAX];
  m->bar.reset(m->n);
  std::vector<std::thread> th;
  for (int r = 0; r < m->n; ++r)
    th.emplace_back([&, r] {
      if (m->bind_numa) pscl_bind_thread_to_device(m->dev[r]);
      cudaSetDevice(m->dev[r]);
      rc[r] = fn(r);
      if (rc[r] != PSCL_OK) m->bar.abort();
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < m->n; ++r)
    if (rc[r] != PSCL_OK) {
      const char* e = pscl_last_error(m->ctx[r]);
      return multi_fail(m, rc[r], "GPU " + std::to_string(m->dev[r]) + ": " + ((e && *e) ? e : "another GPU's worker failed"));
    }
  return PSCL_OK;
}

static double multi_now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// contiguous ranges [cut[r], cut[r+1]) of a non-decreasing prefix array ptr[0..n] with near-equal ptr differences
static void multi_balanced_cuts(const int64_t* ptr, int64_t n, int parts, int64_t* cut) {
  const int64_t total = ptr[n] - ptr[0];
  cut[0] = 0;
  for (int r = 1; r < parts; ++r) {
    const int64_t want = ptr[0] + (int64_t)((__int128)total * r / parts);
    int64_t c = std::lower_bound(ptr, ptr + n + 1, want) - ptr;
    cut[r] = std::max(cut[r - 1], std::min(c, n));
  }
  cut[parts] = n;
}

// ------------------------------------------------------------------------------------------------------------------------
// demuxlet
// ------------------------------------------------------------------------------------------------------------------------
// the wide FP64 table reaches GPU r from GPU `src` over NVLink instead of over PCIe a second time
static int demux_set_geno_peer(pscl_ctx* ctx, const pscl_ctx* src, cudaEvent_t src_ready) {
  PsclScope scope__(ctx);
  PSCL_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t bytes = sizeof(double) * (size_t)src->geno_V * src->nv * 3;
  int rc;
  if ((rc = demux_geno_reserve(ctx, bytes, src->has_gp ? (size_t)src->geno_V : 0)) != PSCL_OK) return rc;
  PSCL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, src_ready, 0));
  PSCL_CUDA(ctx, cudaMemcpyPeerAsync(ctx->gp, ctx->device, src->gp, src->device, bytes, ctx->stream));
  if (src->has_gp) {
    PSCL_CUDA(ctx, cudaMemcpyPeerAsync(ctx->has_gp, ctx->device, src->has_gp, src->device, (size_t)src->geno_V, ctx->stream));
  }
  if (ctx->h_geno_bad) *ctx->h_geno_bad = 0;
  return demux_geno_finish(ctx, src->nv, src->geno_V);
}

extern "C" int pscl_multi_demux_run(pscl_multi* m, const pscl_pileup* host, const pscl_geno* geno, const pscl_demux_opts* opts,
                                    pscl_demux_cell* out, double* llk_grid) {
  if (!m) return PSCL_EINVAL;
  if (!host || !geno || !opts || !out || !host->cell_ptr) return multi_fail(m, PSCL_EINVAL, "pscl_multi_demux_run: NULL argument");
  const int N = m->n;
  const int32_t C = host->n_cells;
  const int64_t P = host->n_pairs;
  if (C < 0 || P < 0 || host->cell_ptr[0] != 0 || host->cell_ptr[C] != P) return multi_fail(m, PSCL_EINVAL, "cell_ptr must run from 0 to n_pairs");
  // barcode ranges balanced by pair count; base-call offset of every cut
  int64_t cut[PSCL_MULTI_MAX + 1], rcut[PSCL_MULTI_MAX + 1];
  multi_balanced_cuts(host->cell_ptr, C, N, cut);
  for (int r = 0; r <= N; ++r) {
    const int64_t p = host->cell_ptr[cut[r]];
    if (host->pair_read_ptr) rcut[r] = host->pair_read_ptr[p];
    else if (host->pair_read_ptr32) rcut[r] = (int64_t)host->pair_read_ptr32[p];
    else if (host->pair_nreads2 && host->nreads_big_ptr) {  // ABI 6 counts: two-bit fields + the large counts on the side
      int64_t s = r ? rcut[r - 1] : 0;
      const int64_t q0 = r ? host->cell_ptr[cut[r - 1]] : 0;
      int64_t big = host->nreads_big_ptr[q0 / 1024];  // rank of the first marker at or after q0
      for (int64_t q = (q0 / 1024) * 1024; q < q0; ++q) big += ((host->pair_nreads2[q >> 2] >> (2 * (q & 3))) & 3) == 0;
      for (int64_t q = q0; q < p; ++q) {
        const int f = (host->pair_nreads2[q >> 2] >> (2 * (q & 3))) & 3;
        s += f ? f : (big < host->n_nreads_big ? host->nreads_big[big++] : 0);
      }
      rcut[r] = s;
    } else if (host->pair_nreads8) {  // counts only: sum them up to the cut (continuing from the previous cut)
      int64_t s = r ? rcut[r - 1] : 0;
      for (int64_t q = r ? host->cell_ptr[cut[r - 1]] : 0; q < p; ++q) s += host->pair_nreads8[q];
      rcut[r] = s;
    } else return multi_fail(m, PSCL_EINVAL, "pscl_pileup has no read offsets");
  }
  // the wide table travels once over PCIe and then GPU to GPU (a binary tree over NVLink) when it is large
  const size_t gp_bytes = geno->gp ? sizeof(double) * (size_t)host->n_snps * geno->n_samples * 3 : 0;
  const bool tree = N > 1 && m->peer_ok && gp_bytes >= ((size_t)32 << 20) && getenv("PSCL_NO_GENO_TREE") == nullptr;
  std::vector<std::vector<int64_t>> cp(N);
  std::vector<std::vector<uint8_t>> aqs(N);
  m->tm = pscl_multi_timing{};
  m->tm.n_gpus = N;
  const size_t G = (size_t)geno->n_samples * geno->n_samples * opts->n_alpha;
  const double t_begin = multi_now_ms();
  const int rc = multi_run_workers(m, [&](int r) -> int {
    pscl_ctx* ctx = m->ctx[r];
    const double t0 = multi_now_ms();
    int rc2 = PSCL_OK;
    if (!tree || r == 0) {
      rc2 = pscl_demux_set_geno(ctx, geno, host->n_snps);
      if (rc2 == PSCL_OK && tree && cudaEventRecord(m->ev_geno[0], ctx->stream) != cudaSuccess) rc2 = PSCL_ECUDA;
    }
    if (tree) {
      // level l: ranks [2^l, 2^(l+1)) copy from rank - 2^l; a host barrier per level orders the event records
      for (int lvl = 1; lvl < N; lvl <<= 1) {
        if (!m->bar.wait()) return PSCL_ECUDA;
        if (rc2 == PSCL_OK && r >= lvl && r < 2 * lvl) {
          rc2 = demux_set_geno_peer(ctx, m->ctx[r - lvl], m->ev_geno[r - lvl]);
          if (rc2 == PSCL_OK && cudaEventRecord(m->ev_geno[r], ctx->stream) != cudaSuccess) rc2 = PSCL_ECUDA;
        }
      }
    }
    if (rc2 != PSCL_OK) return rc2;
    // this rank's barcodes: pointers into the caller's arrays, a rebased copy of cell_ptr
    const int64_t c0 = cut[r], c1 = cut[r + 1], p0 = host->cell_ptr[c0], p1 = host->cell_ptr[c1];
    std::vector<int64_t>& mine = cp[r];
    mine.resize((size_t)(c1 - c0) + 1);
    for (int64_t c = c0; c <= c1; ++c) mine[(size_t)(c - c0)] = host->cell_ptr[c] - p0;
    pscl_pileup sh = *host;
    sh.cell_read_ptr = nullptr;  // (global offsets: a shard view is uploaded whole)
    sh.n_cells = (int32_t)(c1 - c0); sh.n_pairs = p1 - p0; sh.n_reads = rcut[r + 1] - rcut[r];
    sh.cell_ptr = mine.data();
    if (host->pair_snp) sh.pair_snp = host->pair_snp + p0;
    if (host->pair_read_ptr) sh.pair_read_ptr = host->pair_read_ptr + p0;
    if (host->pair_read_ptr32) sh.pair_read_ptr32 = host->pair_read_ptr32 + p0;
    if (host->read_allele) sh.read_allele = host->read_allele + rcut[r];
    if (host->read_qual) sh.read_qual = host->read_qual + rcut[r];
    if (host->read_aq) sh.read_aq = host->read_aq + rcut[r];
    if (host->read_packed) {  // a bit string cannot be entered at a base-call offset: this shard's bytes are unpacked on the host
      if (!host->read_aq && host->read_palette && host->read_bits >= 4 && host->read_bits <= 6) {
        std::vector<uint8_t>& aq = aqs[r];
        aq.resize((size_t)sh.n_reads);
        const int bits = host->read_bits;
        for (int64_t i = 0; i < sh.n_reads; ++i) {
          const int64_t o = (rcut[r] + i) * bits;
          const unsigned w = (unsigned)host->read_packed[o >> 3] | ((unsigned)host->read_packed[(o >> 3) + 1] << 8);
          aq[(size_t)i] = host->read_palette[(w >> (int)(o & 7)) & ((1u << bits) - 1u)];
        }
        sh.read_aq = aq.data();
      }
      sh.read_packed = nullptr; sh.read_palette = nullptr; sh.read_bits = 0;
    }
    if (host->cell_first_snp) sh.cell_first_snp = host->cell_first_snp + c0;
    if (host->pair_snp_delta16) sh.pair_snp_delta16 = host->pair_snp_delta16 + p0;
    if (host->pair_nreads8) sh.pair_nreads8 = host->pair_nreads8 + p0;
    // ABI 6: the gaps and each cell's first large gap are offset; the count arrays stay the caller's (they are indexed by
    // global pair numbers, plp_upload_impl gets pair_base = p0)
    if (host->pair_snp_delta8) sh.pair_snp_delta8 = host->pair_snp_delta8 + p0;
    if (host->cell_gap_big_ptr) sh.cell_gap_big_ptr = host->cell_gap_big_ptr + c0;
    PsclScope scope__(ctx);
    pscl_plp* plp = nullptr;
    rc2 = plp_upload_impl(ctx, &sh, &plp, 1, rcut[r], p0);
    if (rc2 != PSCL_OK) return rc2;
    const double t1 = multi_now_ms();
    const bool keep = ctx->keep_grid;
    if (llk_grid) ctx->keep_grid = true;
    rc2 = pscl_demux_score(ctx, plp, opts, 0, sh.n_cells);
    if (rc2 == PSCL_OK) rc2 = pscl_demux_fetch(ctx, out + c0, llk_grid ? llk_grid + (size_t)c0 * G : nullptr);
    ctx->keep_grid = keep;
    float k_ms = 0.f, k_tot = 0.f;
    if (rc2 == PSCL_OK) pscl_demux_last_kernel_ms(ctx, &k_ms, &k_tot);
    const std::string e = ctx->err;
    pscl_plp_free(ctx, plp);
    ctx->err = e;
    const double t2 = multi_now_ms();
    m->tm.upload_ms[r] = t1 - t0; m->tm.compute_ms[r] = t2 - t1; m->tm.kernel_ms[r] = k_tot;
    m->tm.units[r] = sh.n_pairs;
    return rc2;
  });
  m->tm.total_ms = multi_now_ms() - t_begin;
  return rc;
}

// ------------------------------------------------------------------------------------------------------------------------
// all-reduce over peer memory
// ------------------------------------------------------------------------------------------------------------------------
struct PeerPtrs { double* p[PSCL_MULTI_MAX]; };

// This GPU owns elements [begin, end): sum of the N partial buffers in rank order -> every GPU's result buffer.
__global__ void __launch_bounds__(256) k_p2p_allreduce(PeerPtrs in, PeerPtrs out, int n, size_t begin, size_t end) {
  const size_t stride = 2 * (size_t)gridDim.x * blockDim.x;
  const size_t end2 = begin + ((end - begin) & ~(size_t)1);  // pairs of doubles (begin is even by construction)
  for (size_t k = begin + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); k < end2; k += stride) {
    double2 s = *reinterpret_cast<const double2*>(in.p[0] + k);
    for (int q = 1; q < n; ++q) { const double2 t = *reinterpret_cast<const double2*>(in.p[q] + k); s.x += t.x; s.y += t.y; }
    for (int q = 0; q < n; ++q) *reinterpret_cast<double2*>(out.p[q] + k) = s;
  }
  if (end2 < end && blockIdx.x == 0 && threadIdx.x == 0) {  // odd tail element
    double s = in.p[0][end2];
    for (int q = 1; q < n; ++q) s += in.p[q][end2];
    for (int q = 0; q < n; ++q) out.p[q][end2] = s;
  }
}

struct MultiReduce {  // buffers of one all-reduced quantity, one pair per GPU
  PeerPtrs in{}, out{};
  size_t n = 0;
};

// Called by every worker: in[r] holds this GPU's partial values (work enqueued on its stream).  Returns with the sum
// enqueued into out[r] on every GPU's stream (two host barriers; no device-wide synchronisation).
static int multi_allreduce(pscl_multi* m, int r, const MultiReduce& b) {
  pscl_ctx* ctx = m->ctx[r];
  const int N = m->n;
  PSCL_CUDA(ctx, cudaEventRecord(m->ev_in[r], ctx->stream));
  if (!m->bar.wait()) return PSCL_ECUDA;  // every partial's event has been recorded
  for (int q = 0; q < N; ++q)
    if (q != r) PSCL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, m->ev_in[q], 0));
  const size_t per = (((b.n + N - 1) / N) + 1) & ~(size_t)1;  // even slice starts: 16-byte aligned double2 accesses
  const size_t begin = std::min(b.n, per * r), end = std::min(b.n, per * (r + 1));
  if (end > begin) {
    const int grid = (int)std::min<size_t>((size_t)ctx->sm_count * 4, ((end - begin) / 2 + 255) / 256 + 1);
    k_p2p_allreduce<<<grid, 256, 0, ctx->stream>>>(b.in, b.out, N, begin, end);
    ctx->launches++;
    PSCL_CUDA(ctx, cudaGetLastError());
  }
  PSCL_CUDA(ctx, cudaEventRecord(m->ev_out[r], ctx->stream));
  if (!m->bar.wait()) return PSCL_ECUDA;  // every slice's event has been recorded
  for (int q = 0; q < N; ++q)
    if (q != r) PSCL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, m->ev_out[q], 0));
  return PSCL_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// SNP shard of a device-resident pileup image (freemuxlet): the pairs whose SNP id lies in [v0, v1), cells keep their ids
// ------------------------------------------------------------------------------------------------------------------------
struct KeepFlag {
  const int32_t* snp; int32_t v0, v1;
  __host__ __device__ uint32_t operator()(int64_t p) const { const int32_t s = snp[p]; return (s >= v0 && s < v1) ? 1u : 0u; }
};
struct KeepReads {
  const int32_t* snp; const uint32_t* rd; int32_t v0, v1;
  __host__ __device__ uint32_t operator()(int64_t p) const { const int32_t s = snp[p]; return (s >= v0 && s < v1) ? rd[p + 1] - rd[p] : 0u; }
};
__global__ void k_shard_scatter(const int32_t* __restrict__ snp, const uint32_t* __restrict__ rd, const uint8_t* __restrict__ aq,
                                const uint32_t* __restrict__ pos, const uint32_t* __restrict__ rpos, int64_t P, int32_t v0, int32_t v1,
                                int32_t* __restrict__ snp_o, uint32_t* __restrict__ rd_o, uint8_t* __restrict__ aq_o) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p > P) return;
  if (p == P) { rd_o[pos[P]] = rpos[P]; return; }
  const int32_t s = snp[p];
  if (s < v0 || s >= v1) return;
  const uint32_t q = pos[p], r0 = rd[p], n = rd[p + 1] - r0, w = rpos[p];
  snp_o[q] = s;
  rd_o[q] = w;
  for (uint32_t i = 0; i < n; ++i) aq_o[w + i] = aq[r0 + i];
}
__global__ void k_shard_cell_ptr(const int64_t* __restrict__ cell_ptr, const uint32_t* __restrict__ pos, int32_t C, int64_t* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c <= C) out[c] = (int64_t)pos[cell_ptr[c]];
}
__global__ void k_snp_hist(const int32_t* __restrict__ snp, int64_t P, unsigned int* __restrict__ hist) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P) atomicAdd(&hist[snp[p]], 1u);
}

static int plp_filter_snps(pscl_ctx* ctx, const pscl_plp* full, int32_t v0, int32_t v1, pscl_plp** out) {
  *out = nullptr;
  const int64_t P = full->P;
  const int32_t C = full->C;
  pscl_plp* p = new pscl_plp();
  p->C = C; p->V = full->V;
  uint32_t *pos = nullptr, *rpos = nullptr;
  void* tmp = nullptr;
  size_t tb1 = 0, tb2 = 0;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** d, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(d, bytes ? bytes : 16); };
  alloc((void**)&pos, sizeof(uint32_t) * (size_t)(P + 1));
  alloc((void**)&rpos, sizeof(uint32_t) * (size_t)(P + 1));
  thrust::counting_iterator<int64_t> idx(0);
  auto it1 = thrust::make_transform_iterator(idx, KeepFlag{full->pair_snp, v0, v1});
  auto it2 = thrust::make_transform_iterator(idx, KeepReads{full->pair_snp, full->pair_rd, v0, v1});
  // the scans read P flags and write P + 1 offsets: the last input is a dummy (index P is never dereferenced by a
  // transform of an exclusive scan's LAST element only if we scan P + 1 items), so scan P items and take the totals apart
  uint32_t hP = 0, hN = 0;
  if (P > 0) {
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tb1, it1, pos, (int64_t)P, ctx->stream);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tb2, it2, rpos, (int64_t)P, ctx->stream);
    alloc(&tmp, std::max(tb1, tb2));
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tb1, it1, pos, (int64_t)P, ctx->stream);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tb2, it2, rpos, (int64_t)P, ctx->stream);
    ctx->launches += 2;
    // totals = last exclusive value + last element
    uint32_t lastpos = 0, lastr = 0, last_rd[2] = {0, 0};
    int32_t last_snp = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&lastpos, pos + P - 1, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&lastr, rpos + P - 1, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&last_snp, full->pair_snp + P - 1, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(last_rd, full->pair_rd + P - 1, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    const bool kept = last_snp >= v0 && last_snp < v1;
    hP = lastpos + (kept ? 1u : 0u);
    hN = lastr + (kept ? last_rd[1] - last_rd[0] : 0u);
    if (e == cudaSuccess) e = cudaMemcpyAsync(pos + P, &hP, 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(rpos + P, &hN, 4, cudaMemcpyHostToDevice, ctx->stream);
  } else {
    if (e == cudaSuccess) e = cudaMemsetAsync(pos, 0, 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(rpos, 0, 4, ctx->stream);
  }
  p->P = hP; p->N = hN;
  alloc((void**)&p->cell_ptr, sizeof(int64_t) * ((size_t)C + 1));
  alloc((void**)&p->pair_snp, sizeof(int32_t) * (size_t)(hP ? hP : 1));
  alloc((void**)&p->pair_rd, sizeof(uint32_t) * ((size_t)hP + 1));
  alloc((void**)&p->rd_aq, hN ? hN : 16);
  alloc((void**)&p->d_bad, sizeof(int));
  if (e == cudaSuccess) e = cudaMemsetAsync(p->d_bad, 0, sizeof(int), ctx->stream);
  if (full->snp_af) {
    alloc((void**)&p->snp_af, sizeof(double) * (size_t)(full->V ? full->V : 1));
    if (e == cudaSuccess && full->V) e = cudaMemcpyAsync(p->snp_af, full->snp_af, sizeof(double) * (size_t)full->V, cudaMemcpyDeviceToDevice, ctx->stream);
  }
  if (e == cudaSuccess) {
    k_shard_scatter<<<(unsigned)((P + 1 + 255) / 256), 256, 0, ctx->stream>>>(full->pair_snp, full->pair_rd, full->rd_aq, pos, rpos, P, v0, v1,
                                                                             p->pair_snp, p->pair_rd, p->rd_aq);
    k_shard_cell_ptr<<<(unsigned)((C + 1 + 255) / 256), 256, 0, ctx->stream>>>(full->cell_ptr, pos, C, p->cell_ptr);
    ctx->launches += 2;
    e = cudaGetLastError();
  }
  std::vector<int64_t> h_cp((size_t)C + 1);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_cp.data(), p->cell_ptr, sizeof(int64_t) * ((size_t)C + 1), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = plp_make_items(ctx, p, h_cp.data(), false);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(pos); cudaFree(rpos); cudaFree(tmp);
  if (e != cudaSuccess) {
    pscl_plp_free(ctx, p);
    return pscl_fail(ctx, e == cudaErrorMemoryAllocation ? PSCL_ENOMEM : PSCL_ECUDA, "SNP shard of the pileup image failed: %s", cudaGetErrorString(e));
  }
  *out = p;
  return PSCL_OK;
}

// SNP ranges with near-equal pair counts from the image's own SNP histogram
static int plp_snp_cuts(pscl_ctx* ctx, const pscl_plp* full, int parts, int64_t* cut) {
  const int32_t V = full->V;
  unsigned int* d_hist = nullptr;
  PSCL_CUDA(ctx, cudaMalloc((void**)&d_hist, sizeof(unsigned int) * (size_t)(V ? V : 1)));
  PSCL_CUDA(ctx, cudaMemsetAsync(d_hist, 0, sizeof(unsigned int) * (size_t)(V ? V : 1), ctx->stream));
  if (full->P > 0) {
    k_snp_hist<<<(unsigned)((full->P + 255) / 256), 256, 0, ctx->stream>>>(full->pair_snp, full->P, d_hist);
    ctx->launches++;
    PSCL_CUDA(ctx, cudaGetLastError());
  }
  std::vector<unsigned int> h((size_t)(V ? V : 1));
  PSCL_CUDA(ctx, cudaMemcpyAsync(h.data(), d_hist, sizeof(unsigned int) * (size_t)V, cudaMemcpyDeviceToHost, ctx->stream));
  PSCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(d_hist);
  std::vector<int64_t> cum((size_t)V + 1, 0);
  for (int32_t v = 0; v < V; ++v) cum[(size_t)v + 1] = cum[v] + h[v];
  multi_balanced_cuts(cum.data(), V, parts, cut);
  return PSCL_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// freemuxlet
// ------------------------------------------------------------------------------------------------------------------------
extern "C" int pscl_multi_fmx_run(pscl_multi* m, const pscl_pileup* host, const pscl_fmx_opts* opts, const int32_t* init_clust,
                                  pscl_fmx_cell* out, double* clust_gl, int32_t* clust_cnt, pscl_fmx_result* res) {
  if (!m) return PSCL_EINVAL;
  if (!host || !opts || !out) return multi_fail(m, PSCL_EINVAL, "pscl_multi_fmx_run: NULL argument");
  const int N = m->n;
  if (N > 1 && !m->peer_ok) return multi_fail(m, PSCL_ENODEV, "the GPUs cannot map each other's memory (no NVLink / PCIe peer access): SNP-sharded freemuxlet needs it");
  const int32_t C = host->n_cells, V = host->n_snps, nS = opts->n_clusters;
  const size_t C1 = (size_t)(C ? C : 1), npairs = (size_t)nS * (nS + 1) / 2;
  if (init_clust)
    for (int32_t c = 0; c < C; ++c)
      if (init_clust[c] >= nS) return multi_fail(m, PSCL_EINVAL, "init_clust[" + std::to_string(c) + "] is not below n_clusters");
  // shared between the workers
  std::vector<double> h_stage1(4 * C1);
  std::vector<int32_t> h_clust(C1, -1);
  if (init_clust) std::copy(init_clust, init_clust + C, h_clust.begin());
  int64_t vcut[PSCL_MULTI_MAX + 1] = {0};
  MultiReduce red_llk, red_s1;
  red_llk.n = C1 * npairs; red_s1.n = 4 * C1;
  pscl_fmx_result results[PSCL_MULTI_MAX];
  memset(results, 0, sizeof results);
  // seeding needs every SNP of a cell: one GPU does it over the whole pileup, the shards then start from its clusters
  const bool seed_whole = N > 1 && (!init_clust || (opts->mode_old && opts->iter_init > 0));
  pscl_fmx_opts shard_opts = *opts;
  shard_opts.iter_init = 0;
  m->tm = pscl_multi_timing{};
  m->tm.n_gpus = N;
  const double t_begin = multi_now_ms();
  const int rc = multi_run_workers(m, [&](int r) -> int {
    pscl_ctx* ctx = m->ctx[r];
    PsclScope scope__(ctx);
    pscl_plp *full = nullptr, *shard = nullptr;
    double *d_s1 = nullptr, *d_s1_sum = nullptr, *d_llk = nullptr, *d_llk_sum = nullptr;
    int32_t *d_clust = nullptr, *d_init = nullptr;
    int rc2 = PSCL_OK;
    auto body = [&]() -> int {
      const double t0 = multi_now_ms();
      int rr = pscl_plp_upload(ctx, host, &full);  // every GPU takes the whole (compact) pileup over its own PCIe link
      if (rr != PSCL_OK) return rr;
      // the four buffers other GPUs read or write: plain (peer-mapped) device memory, not the pool
      PSCL_CUDA(ctx, (cudaMalloc)((void**)&d_s1, sizeof(double) * 4 * C1));
      PSCL_CUDA(ctx, (cudaMalloc)((void**)&d_s1_sum, sizeof(double) * 4 * C1));
      PSCL_CUDA(ctx, (cudaMalloc)((void**)&d_llk, sizeof(double) * C1 * npairs));
      PSCL_CUDA(ctx, (cudaMalloc)((void**)&d_llk_sum, sizeof(double) * C1 * npairs));
      PSCL_CUDA(ctx, cudaMalloc((void**)&d_clust, sizeof(int32_t) * C1));
      PSCL_CUDA(ctx, cudaMalloc((void**)&d_init, sizeof(int32_t) * C1));
      red_s1.in.p[r] = d_s1; red_s1.out.p[r] = d_s1_sum; red_llk.in.p[r] = d_llk; red_llk.out.p[r] = d_llk_sum;
      if (r == 0) {
        if ((rr = plp_snp_cuts(ctx, full, N, vcut)) != PSCL_OK) return rr;
        if (seed_whole) {
          // stage 1 + seeding over ALL SNPs on this GPU: the greedy chain over the cells (cmd_cram_freemux2.cpp:223-260),
          // or freemuxlet-old's pairwise matrix and votes (cmd_cram_freemuxlet.cpp:165-346, also refining a given init_clust)
          const double ts = multi_now_ms();
          if ((rr = pscl_fmx_init(ctx, full, opts)) != PSCL_OK) return rr;
          if ((rr = pscl_fmx_stage1(ctx, d_s1)) != PSCL_OK) return rr;
          if (init_clust) PSCL_CUDA(ctx, cudaMemcpyAsync(d_init, h_clust.data(), sizeof(int32_t) * C1, cudaMemcpyHostToDevice, ctx->stream));
          if ((rr = pscl_fmx_seed(ctx, d_s1, init_clust ? d_init : nullptr, d_clust)) != PSCL_OK) return rr;
          PSCL_CUDA(ctx, cudaMemcpyAsync(h_stage1.data(), d_s1, sizeof(double) * 4 * C1, cudaMemcpyDeviceToHost, ctx->stream));
          PSCL_CUDA(ctx, cudaMemcpyAsync(h_clust.data(), d_clust, sizeof(int32_t) * C1, cudaMemcpyDeviceToHost, ctx->stream));
          PSCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
          fmx_state_free(ctx);
          m->tm.seed_ms = multi_now_ms() - ts;
        }
      }
      if (!m->bar.wait()) return PSCL_ECUDA;  // cuts, seeds and the buffer addresses of every GPU are published
      const bool seeded = seed_whole;
      if (N > 1) {
        if ((rr = plp_filter_snps(ctx, full, (int32_t)vcut[r], (int32_t)vcut[r + 1], &shard)) != PSCL_OK) return rr;
        pscl_plp_free(ctx, full);
        full = nullptr;
      } else {
        shard = full;
        full = nullptr;
      }
      const double t1 = multi_now_ms();
      if ((rr = pscl_fmx_init(ctx, shard, seeded ? &shard_opts : opts)) != PSCL_OK) return rr;
      pscl_fmx_state* s = ctx->fmx;
      if ((rr = pscl_fmx_stage1(ctx, d_s1)) != PSCL_OK) return rr;
      const double* s1_use = d_s1;
      if (seeded) {  // the whole-pileup sums of GPU 0: the same bits as a single-GPU run prints in .lmix
        PSCL_CUDA(ctx, cudaMemcpyAsync(d_s1_sum, h_stage1.data(), sizeof(double) * 4 * C1, cudaMemcpyHostToDevice, ctx->stream));
        s1_use = d_s1_sum;
      } else if (N > 1) {
        if ((rr = multi_allreduce(m, r, red_s1)) != PSCL_OK) return rr;
        s1_use = d_s1_sum;
      }
      const bool have_init = init_clust || seeded;
      if (have_init) PSCL_CUDA(ctx, cudaMemcpyAsync(d_init, h_clust.data(), sizeof(int32_t) * C1, cudaMemcpyHostToDevice, ctx->stream));
      if ((rr = pscl_fmx_seed(ctx, s1_use, have_init ? d_init : nullptr, d_clust)) != PSCL_OK) return rr;
      if ((rr = pscl_fmx_mstep(ctx, d_clust)) != PSCL_OK) return rr;  // :277-288
      const double t2 = multi_now_ms();
      pscl_fmx_result rres;
      memset(&rres, 0, sizeof rres);
      double ar_ms = 0.0;
      int iters = 0;
      for (int iter = 0; iter < s->o.max_iter; ++iter) {
        if ((rr = pscl_fmx_estep(ctx, iter, d_llk)) != PSCL_OK) return rr;
        const double* llk_use = d_llk;
        if (N > 1) {
          const double ta = multi_now_ms();
          if (getenv("PSCL_MULTI_TIME_ALLREDUCE")) PSCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // isolate the collective (bench)
          const double tb = multi_now_ms();
          if ((rr = multi_allreduce(m, r, red_llk)) != PSCL_OK) return rr;
          if (getenv("PSCL_MULTI_TIME_ALLREDUCE")) { PSCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); ar_ms += multi_now_ms() - tb; }
          (void)ta;
          llk_use = d_llk_sum;
        }
        if ((rr = pscl_fmx_classify(ctx, llk_use, d_clust, &rres)) != PSCL_OK) return rr;
        if ((rr = pscl_fmx_mstep(ctx, nullptr)) != PSCL_OK) return rr;
        ++iters;
        if (!s->o.mode_old && s->o.early_stop && rres.n_changed == 0) break;  // :601-604 — the same decision on every GPU
      }
      results[r] = rres;
      PSCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      const double t3 = multi_now_ms();  // the EM loop; the read-back below (cluster pileups: V x nS x 84 B) only shows in total_ms
      // records from GPU 0; every GPU holds the cluster pileups of its own SNP range
      const int32_t w0 = N > 1 ? (int32_t)vcut[r] : 0, w1 = N > 1 ? (int32_t)vcut[r + 1] : V;
      if ((rr = fmx_fetch_range(ctx, r == 0 ? out : nullptr, clust_gl, clust_cnt, w0, w1)) != PSCL_OK) return rr;
      m->tm.upload_ms[r] = t1 - t0; m->tm.setup_ms[r] = t2 - t1; m->tm.compute_ms[r] = t3 - t2; m->tm.units[r] = shard->P;
      if (r == 0) { m->tm.iters = iters; m->tm.allreduce_ms = iters ? ar_ms / iters : 0.0; m->tm.allreduce_bytes = (int64_t)(red_llk.n * sizeof(double)); }
      return PSCL_OK;
    };
    rc2 = body();
    const std::string e = ctx->err;
    // nobody frees a buffer another GPU may still be reading
    if (rc2 == PSCL_OK) { cudaStreamSynchronize(ctx->stream); m->bar.wait(); }
    (cudaFree)(d_s1); (cudaFree)(d_s1_sum); (cudaFree)(d_llk); (cudaFree)(d_llk_sum); cudaFree(d_clust); cudaFree(d_init);
    fmx_state_free(ctx);
    if (shard) pscl_plp_free(ctx, shard);
    if (full) pscl_plp_free(ctx, full);
    ctx->err = e;
    return rc2;
  });
  m->tm.total_ms = multi_now_ms() - t_begin;
  if (rc == PSCL_OK && res) *res = results[0];
  return rc;
}
