// gather_bench.cu — micro-benchmark of the genotype-row gather of the demuxlet kernels on one B200:
// R random 128-byte (or 192-byte) rows out of an L2-resident table per second, by mechanism.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench tools/gather_bench.cu && ./gather_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp16ca(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void waitg() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// mode 0: cooperative LDGSTS.cg (ROWB/16 pieces per row), DEPTH batches in flight per warp
// mode 1: the same with .ca
// mode 2: every lane loads its own row with LDG.128 into registers (DEPTH ignored: ROWB/16 loads in flight per lane)
// mode 3: every lane issues one bulk copy for its row, completion on an mbarrier per buffer
// mode 4: as mode 0, but all source addresses of a batch are computed BEFORE the first LDGSTS, so every copy has
//         its own address registers (no write-after-read wait on the registers of the copy in front)
template <int ROWB, int DEPTH, int MODE, int STRIDE_O = 0, int WORK = 0>
__global__ void __launch_bounds__(256) k_gather(const unsigned char* __restrict__ tab, const int32_t* __restrict__ idx, int64_t n_batches,
                                                double* __restrict__ sink) {
  constexpr int STRIDE = STRIDE_O ? STRIDE_O : ((ROWB / 16) | 1) * 16, NCH = ROWB / 16;
  constexpr int LPR = NCH <= 8 ? 8 : 16, RPI = 32 / LPR;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* rows = smem + (size_t)warp * ((DEPTH + 1) * 32 * STRIDE + 128);
  const uint32_t rows_u32 = (uint32_t)__cvta_generic_to_shared(rows);
  const uint32_t bar_u32 = rows_u32 + (DEPTH + 1) * 32 * STRIDE;
  const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
  const int piece = lane % LPR, rsub = lane / LPR;
  double acc = 0.0;
  if (MODE == 3) {
    if (lane == 0) for (int i = 0; i <= DEPTH; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_u32 + i * 8), "r"(1) : "memory");
    __syncwarp();
  }
  auto issue = [&](int64_t b, int buf) {
    const int snp = idx[b * 32 + lane];
    if (MODE == 4) {
      const unsigned char* src[LPR];
#pragma unroll
      for (int i = 0; i < LPR; ++i) {
        const int s = __shfl_sync(0xffffffffu, snp, i * RPI + rsub);
        src[i] = tab + (size_t)s * ROWB + piece * 16;
      }
#pragma unroll
      for (int i = 0; i < LPR; ++i)
        if (piece < NCH) cp16(rows_u32 + buf * 32 * STRIDE + (i * RPI + rsub) * STRIDE + piece * 16, src[i]);
      commit();
    } else if (MODE == 0 || MODE == 1) {
#pragma unroll
      for (int i = 0; i < LPR; ++i) {
        const int s = __shfl_sync(0xffffffffu, snp, i * RPI + rsub);
        if (piece < NCH) {
          const uint32_t dst = rows_u32 + buf * 32 * STRIDE + (i * RPI + rsub) * STRIDE + piece * 16;
          const unsigned char* src = tab + (size_t)s * ROWB + piece * 16;
          if (MODE == 0) cp16(dst, src); else cp16ca(dst, src);
        }
      }
      commit();
    } else if (MODE == 3) {
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_u32 + buf * 8), "r"(32 * ROWB) : "memory");
      __syncwarp();
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(rows_u32 + buf * 32 * STRIDE + lane * STRIDE), "l"(tab + (size_t)snp * ROWB), "r"(ROWB), "r"(bar_u32 + buf * 8) : "memory");
    }
  };
  if (MODE == 2) {
    for (int64_t b = gw; b < n_batches; b += nw) {
      const int snp = idx[b * 32 + lane];
      const double2* r = reinterpret_cast<const double2*>(tab + (size_t)snp * ROWB);
      double2 v[NCH];
#pragma unroll
      for (int i = 0; i < NCH; ++i) v[i] = r[i];
#pragma unroll
      for (int i = 0; i < NCH; ++i) acc += v[i].x * v[i].y;
    }
  } else {
    int64_t b = gw;
    int head = 0;
    for (int d = 0; d < DEPTH; ++d) { if (b + (int64_t)d * nw < n_batches) issue(b + (int64_t)d * nw, d); else if (MODE != 3) commit(); }
    uint32_t it = 0;
    for (; b < n_batches; b += nw, ++it) {
      const int buf = it % (DEPTH + 1), nbuf = (it + DEPTH) % (DEPTH + 1);
      __syncwarp();
      if (b + (int64_t)DEPTH * nw < n_batches) issue(b + (int64_t)DEPTH * nw, nbuf); else if (MODE != 3) commit();
      if (MODE == 3) {
        uint32_t ok;
        const uint32_t par = (it / (DEPTH + 1)) & 1u;
        do {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(bar_u32 + buf * 8), "r"(par) : "memory");
        } while (!ok);
      } else {
        waitg<DEPTH>();
        __syncwarp();
      }
      const double2* r = reinterpret_cast<const double2*>(rows + buf * 32 * STRIDE + lane * STRIDE);
#pragma unroll
      for (int i = 0; i < NCH; ++i) { double2 v = r[i]; acc += v.x * v.y; }
      if (WORK) {  // WORK dependent-free DFMA per lane: stands in for the likelihood math
        double w0 = acc, w1 = acc + 1, w2 = acc + 2, w3 = acc + 3;
#pragma unroll
        for (int i = 0; i < WORK / 4; ++i) { w0 = fma(w0, 1.0000001, 1e-9); w1 = fma(w1, 1.0000001, 1e-9); w2 = fma(w2, 1.0000001, 1e-9); w3 = fma(w3, 1.0000001, 1e-9); }
        acc = w0 + w1 + w2 + w3;
      }
      (void)head;
    }
    if (MODE != 3) waitg<0>();
  }
  if (acc == 123.456) sink[0] = acc;
}

// mode 5: cooperative LDG.128 into registers (same lane mapping as mode 0), stored to shared memory one batch later:
//         plain loads pipeline inside a warp, LDGSTS does not (one in flight per warp at issue, ~145 cycles each)
template <int ROWB, int WORK>
__global__ void __launch_bounds__(256) k_gather_ldg(const unsigned char* __restrict__ tab, const int32_t* __restrict__ idx, int64_t n_batches,
                                                    double* __restrict__ sink) {
  constexpr int STRIDE = 208, NCH = ROWB / 16;
  constexpr int LPR = NCH <= 8 ? 8 : 16, RPI = 32 / LPR;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* rows = smem + (size_t)warp * (32 * STRIDE);
  const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
  const int piece = lane % LPR, rsub = lane / LPR;
  double acc = 0.0;
  uint4 v[LPR];
  auto issue = [&](int64_t b) {
    const int snp = idx[b * 32 + lane];
#pragma unroll
    for (int i = 0; i < LPR; ++i) {
      const int s = __shfl_sync(0xffffffffu, snp, i * RPI + rsub);
      if (piece < NCH) v[i] = *reinterpret_cast<const uint4*>(tab + (size_t)s * ROWB + piece * 16);
    }
  };
  int64_t b = gw;
  if (b < n_batches) issue(b);
  for (; b < n_batches; b += nw) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < LPR; ++i)
      if (piece < NCH) *reinterpret_cast<uint4*>(rows + (i * RPI + rsub) * STRIDE + piece * 16) = v[i];
    __syncwarp();
    if (b + nw < n_batches) issue(b + nw);
    const double2* r = reinterpret_cast<const double2*>(rows + lane * STRIDE);
#pragma unroll
    for (int i = 0; i < NCH; ++i) { double2 x = r[i]; acc += x.x * x.y; }
    if (WORK) {
      double w0 = acc, w1 = acc + 1, w2 = acc + 2, w3 = acc + 3;
#pragma unroll
      for (int i = 0; i < WORK / 4; ++i) { w0 = fma(w0, 1.0000001, 1e-9); w1 = fma(w1, 1.0000001, 1e-9); w2 = fma(w2, 1.0000001, 1e-9); w3 = fma(w3, 1.0000001, 1e-9); }
      acc = w0 + w1 + w2 + w3;
    }
  }
  if (acc == 123.456) sink[0] = acc;
}
template <int ROWB, int WORK>
static void run_ldg(const unsigned char* tab, const int32_t* idx, int64_t nb, double* sink, int ctas_per_sm) {
  size_t smem = 8 * 32 * 208;
  CK(cudaFuncSetAttribute(k_gather_ldg<ROWB, WORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int grid = 148 * ctas_per_sm;
  for (int w = 0; w < 2; ++w) k_gather_ldg<ROWB, WORK><<<grid, 256, smem>>>(tab, idx, nb, sink);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int w = 0; w < 5; ++w) k_gather_ldg<ROWB, WORK><<<grid, 256, smem>>>(tab, idx, nb, sink);
  cudaEventRecord(e1);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  printf("%-34s stride=208 work=%3d rowB=%3d depth=1 ctas/SM=%d smem=%6zu : %.3f ms  %.2f Grows/s\n", "LDG.128 cooperative -> regs -> STS", WORK, ROWB, ctas_per_sm, smem, ms, nb * 32 / ms / 1e6);
}

template <int ROWB, int DEPTH, int MODE, int STRIDE_O = 0, int WORK = 0>
static void run(const char* name, const unsigned char* tab, const int32_t* idx, int64_t nb, double* sink, int ctas_per_sm) {
  constexpr int STRIDE = STRIDE_O ? STRIDE_O : ((ROWB / 16) | 1) * 16;
  size_t smem = 8 * ((DEPTH + 1) * 32 * STRIDE + 128);
  CK(cudaFuncSetAttribute(k_gather<ROWB, DEPTH, MODE, STRIDE_O, WORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int grid = 148 * ctas_per_sm;
  for (int w = 0; w < 2; ++w) k_gather<ROWB, DEPTH, MODE, STRIDE_O, WORK><<<grid, 256, smem>>>(tab, idx, nb, sink);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int w = 0; w < 5; ++w) k_gather<ROWB, DEPTH, MODE, STRIDE_O, WORK><<<grid, 256, smem>>>(tab, idx, nb, sink);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  printf("%-34s stride=%3d work=%3d rowB=%3d depth=%d ctas/SM=%d smem=%6zu : %.3f ms  %.2f Grows/s  %.2f TB/s\n", name, STRIDE, WORK, ROWB, DEPTH, ctas_per_sm, smem, ms,
         nb * 32 / ms / 1e6, (double)nb * 32 * ROWB / ms / 1e9);
}

int main() {
  const int V = 100000;
  const int64_t nb = 620000;  // batches of 32 rows (config 2: 19.9 M pairs)
  unsigned char* tab; int32_t* idx; double* sink;
  CK(cudaMalloc(&tab, (size_t)V * 192)); CK(cudaMemset(tab, 0, (size_t)V * 192));
  std::vector<int32_t> h(nb * 32);
  uint64_t s = 88172645463325252ull;
  for (auto& x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (int32_t)(s % V); }
  CK(cudaMalloc(&idx, h.size() * 4)); CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&sink, 8));
  run_ldg<128, 0>(tab, idx, nb, sink, 1);
  run_ldg<128, 128>(tab, idx, nb, sink, 1);
  run_ldg<128, 256>(tab, idx, nb, sink, 1);
  run_ldg<192, 128>(tab, idx, nb, sink, 1);
  run_ldg<128, 128>(tab, idx, nb, sink, 2);
  run_ldg<128, 256>(tab, idx, nb, sink, 2);
  run<128, 1, 4>("addresses first", tab, idx, nb, sink, 1);
  run<128, 2, 4>("addresses first", tab, idx, nb, sink, 1);
  run<128, 2, 4, 208, 128>("addresses first", tab, idx, nb, sink, 1);
  run<128, 2, 4, 208, 256>("addresses first", tab, idx, nb, sink, 1);
  run<192, 2, 4, 208, 128>("addresses first", tab, idx, nb, sink, 1);
  run<128, 2, 4, 208, 128>("addresses first", tab, idx, nb, sink, 2);
  run<128, 1, 0>("uniform random rows", tab, idx, nb, sink, 1);
  run<128, 2, 0, 208, 128>("uniform random rows", tab, idx, nb, sink, 1);
  {  // the kernels' real pattern: a warp's 32 rows are consecutive entries of ONE cell's ascending SNP list
    const int K = 2000;  // SNPs per cell
    std::vector<int32_t> h2(h.size());
    uint64_t s2 = 0x9E3779B97F4A7C15ull;
    for (size_t c0 = 0; c0 < h2.size(); c0 += K) {
      const size_t n = std::min<size_t>(K, h2.size() - c0);
      for (size_t i = 0; i < n; ++i) { s2 ^= s2 << 13; s2 ^= s2 >> 7; s2 ^= s2 << 17; h2[c0 + i] = (int32_t)(s2 % V); }
      std::sort(h2.begin() + c0, h2.begin() + c0 + n);
    }
    CK(cudaMemcpy(idx, h2.data(), h2.size() * 4, cudaMemcpyHostToDevice));
    run<128, 1, 0>("sorted per cell (as the kernels)", tab, idx, nb, sink, 1);
    run<128, 2, 0, 208, 128>("sorted per cell (as the kernels)", tab, idx, nb, sink, 1);
    run<192, 2, 0, 208, 128>("sorted per cell (as the kernels)", tab, idx, nb, sink, 1);
    // the same lists with the row index bit-mixed (a table stored in hashed row order)
    for (auto& x : h2) { uint32_t y = (uint32_t)x * 2654435761u; x = (int32_t)(y % (uint32_t)V); }
    CK(cudaMemcpy(idx, h2.data(), h2.size() * 4, cudaMemcpyHostToDevice));
    run<128, 1, 0>("sorted per cell, hashed row order", tab, idx, nb, sink, 1);
    run<128, 2, 0, 208, 128>("sorted per cell, hashed row order", tab, idx, nb, sink, 1);
  }
  return 0;
}
