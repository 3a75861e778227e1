// pair_body_bench.cu — micro-benchmark of the per-pair arithmetic of k_demux_default's dictionary variant on one B200,
// with everything it reads already in shared memory: how fast can 8 (or 12) warps per SM push pairs through
// fold -> max -> h -> genotype products -> 38 running products, by the ORDER in which the body issues them?
// The final ncu profile of the real kernel (profiles/r06_k_demux_default_dict8_ncu.txt) shows no saturated unit
// (FP64 40 %, issue 45 %, L1/shared 47 %) and stall_wait at 34 %: this isolates that.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pair_body_bench tools/pair_body_bench.cu && ./pair_body_bench
// Variants (template MODE):
//   0  the kernel's order: per doublet (j,k) one dependent chain  term = G_k . v_j ; acc *= term
//   1  the same arithmetic, four (j,k) chains interleaved by hand (explicit __dmul_rn/__fma_rn, so the contraction
//      is pinned and the result is bit-identical to variant 2's)
//   2  variant 0 written with the same explicit intrinsics (reference for the bit-identity check of 1)
//   3  variant 1 + the serial front (fold, max, h) of pair i+1 issued before the wide part of pair i
// Every variant prints ns per pair per SM-resident warp set and a checksum; 1, 2 and 3 must print the same checksum.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int NV = 8, ND = NV * (NV - 1) / 2, NE = NV + ND + 2;
constexpr int FOLD_ROWS = 3 * 64, FOLD_ROW = 6, DICT_N = 256;

__device__ __forceinline__ double pmax(double x, double y) { return x > y ? x : y; }
__device__ __forceinline__ void renorm(double& m, int& e) {
  const int hi = __double2hiint(m);
  const int ex = (hi >> 20) & 0x7ff;
  if (ex != 0 && ex != 0x7ff) { e += ex - 1023; m = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(m)); }
}

struct Front { double h0, h1, h2, h3, h4, mx; };

__device__ __forceinline__ Front front(const double* s_tab, uint32_t b0, uint32_t b1, uint32_t b2) {
  const double* t0 = s_tab + b0 * FOLD_ROW;
  const double* t1 = s_tab + b1 * FOLD_ROW;
  const double* t2 = s_tab + b2 * FOLD_ROW;
  const double f0 = t0[0] * t1[0] * t2[0], f1 = t0[1] * t1[1] * t2[1], f2 = t0[2] * t1[2] * t2[2], f3 = t0[3] * t1[3] * t2[3],
               f4 = t0[4] * t1[4] * t2[4];
  Front r;
  r.mx = pmax(pmax(pmax(f0, f1), pmax(f2, f3)), f4);
  r.h0 = fma(1e-10, r.mx, f0); r.h1 = fma(1e-10, r.mx, f1); r.h2 = fma(1e-10, r.mx, f2); r.h3 = fma(1e-10, r.mx, f3); r.h4 = fma(1e-10, r.mx, f4);
  return r;
}

// G . (a, b, c) with the contraction written out: mul, fma, fma
__device__ __forceinline__ double dot3(const double* g, double a, double b, double c) {
  return __fma_rn(g[2], c, __fma_rn(g[1], b, __dmul_rn(g[0], a)));
}

template <int MODE>
__device__ __forceinline__ void wide(const Front& f, const double (&G)[NV][3], double (&acc)[NE]) {
  acc[NV + ND + 1] *= f.mx;
  if (MODE == 0) {
    acc[NV + ND] *= (G[0][0] + G[0][1] + G[0][2]);
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] *= (G[j][0] * f.h0 + G[j][1] * f.h2 + G[j][2] * f.h4);
#pragma unroll
    for (int j = 1; j < NV; ++j) {
      const double v0 = f.h0 * G[j][0] + f.h1 * G[j][1] + f.h2 * G[j][2];
      const double v1 = f.h1 * G[j][0] + f.h2 * G[j][1] + f.h3 * G[j][2];
      const double v2 = f.h2 * G[j][0] + f.h3 * G[j][1] + f.h4 * G[j][2];
#pragma unroll
      for (int k = 0; k < j; ++k) acc[NV + j * (j - 1) / 2 + k] *= (G[k][0] * v0 + G[k][1] * v1 + G[k][2] * v2);
    }
  } else {
    acc[NV + ND] = __dmul_rn(acc[NV + ND], __dadd_rn(__dadd_rn(G[0][0], G[0][1]), G[0][2]));
    if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < NV; ++j) acc[j] = __dmul_rn(acc[j], dot3(G[j], f.h0, f.h2, f.h4));
    } else {  // four singlet chains at a time
#pragma unroll
      for (int j = 0; j < NV; j += 4) {
        double t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) t[u] = __dmul_rn(G[j + u][0], f.h0);
#pragma unroll
        for (int u = 0; u < 4; ++u) t[u] = __fma_rn(G[j + u][1], f.h2, t[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) t[u] = __fma_rn(G[j + u][2], f.h4, t[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[j + u] = __dmul_rn(acc[j + u], t[u]);
      }
    }
#pragma unroll
    for (int j = 1; j < NV; ++j) {
      const double hv[5] = {f.h0, f.h1, f.h2, f.h3, f.h4};
      double v[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) v[m] = __dmul_rn(hv[m], G[j][0]);
#pragma unroll
      for (int m = 0; m < 3; ++m) v[m] = __fma_rn(hv[m + 1], G[j][1], v[m]);
#pragma unroll
      for (int m = 0; m < 3; ++m) v[m] = __fma_rn(hv[m + 2], G[j][2], v[m]);
      if (MODE == 2) {
#pragma unroll
        for (int k = 0; k < j; ++k) acc[NV + j * (j - 1) / 2 + k] = __dmul_rn(acc[NV + j * (j - 1) / 2 + k], dot3(G[k], v[0], v[1], v[2]));
      } else {  // up to four (j,k) chains in flight
#pragma unroll
        for (int k0 = 0; k0 < j; k0 += 4) {
          double t[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) if (k0 + u < j) t[u] = __dmul_rn(G[k0 + u][0], v[0]);
#pragma unroll
          for (int u = 0; u < 4; ++u) if (k0 + u < j) t[u] = __fma_rn(G[k0 + u][1], v[1], t[u]);
#pragma unroll
          for (int u = 0; u < 4; ++u) if (k0 + u < j) t[u] = __fma_rn(G[k0 + u][2], v[2], t[u]);
#pragma unroll
          for (int u = 0; u < 4; ++u) if (k0 + u < j) acc[NV + j * (j - 1) / 2 + k0 + u] = __dmul_rn(acc[NV + j * (j - 1) / 2 + k0 + u], t[u]);
        }
      }
    }
  }
}

template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_body(const double* __restrict__ fold, const double* __restrict__ dict,
                                                     const uint32_t* __restrict__ bytes, const unsigned long long* __restrict__ codes,
                                                     int iters, double* __restrict__ sink) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* s_tab = reinterpret_cast<double*>(smem);            // [192][6]
  double* s_dict = s_tab + FOLD_ROWS * FOLD_ROW;               // [256*3][16]
  int* s_exp = reinterpret_cast<int*>(s_dict + DICT_N * 48);   // [NE][THREADS]
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < FOLD_ROWS * FOLD_ROW; i += THREADS) s_tab[i] = fold[i];
  for (int i = tid; i < DICT_N * 48; i += THREADS) s_dict[i] = dict[i >> 4];
  for (int e = 0; e < NE; ++e) s_exp[e * THREADS + tid] = 0;
  __syncthreads();
  double acc[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) acc[e] = 1.0;
  // per-lane inputs of one pair: three packed base-call bytes and the genotype code word (L1/L2-resident streams)
  const size_t base = ((size_t)blockIdx.x * THREADS + tid);
  const size_t stride = (size_t)gridDim.x * THREADS;
  auto load_G = [&](unsigned long long code, double (&G)[NV][3]) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const double* d = s_dict + (uint32_t)((code >> (8 * j)) & 255ull) * 48 + (lane & 15);
      G[j][0] = d[0]; G[j][1] = d[16]; G[j][2] = d[32];
    }
  };
  uint32_t b = bytes[base];
  unsigned long long code = codes[base];
  Front fn = front(s_tab, b & 255u, (b >> 8) & 255u, (b >> 16) & 255u);
  for (int it = 0; it < iters; ++it) {
    const size_t nx = base + (size_t)((it + 1) % iters) * stride;
    const uint32_t b_next = bytes[nx];
    const unsigned long long code_next = codes[nx];
    double G[NV][3];
    if (MODE == 3) {
      const Front fc = fn;
      load_G(code, G);
      fn = front(s_tab, b_next & 255u, (b_next >> 8) & 255u, (b_next >> 16) & 255u);  // next pair's serial front first
      wide<1>(fc, G, acc);
    } else {
      const Front fc = front(s_tab, b & 255u, (b >> 8) & 255u, (b >> 16) & 255u);
      load_G(code, G);
      wide<MODE>(fc, G, acc);
    }
    b = b_next; code = code_next;
    if ((it & 7) == 7) {
#pragma unroll
      for (int e = 0; e < NE; ++e) { int ex = 0; renorm(acc[e], ex); s_exp[e * THREADS + tid] += ex; }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int e = 0; e < NE; ++e) s += acc[e] + (double)s_exp[e * THREADS + tid];
  sink[base] = s;
}

template <int MODE, int THREADS>
static void run(const char* name, const double* d_fold, const double* d_dict, const uint32_t* d_bytes, const unsigned long long* d_codes,
                int iters, double* d_sink, size_t n_threads_max) {
  const size_t smem = sizeof(double) * (FOLD_ROWS * FOLD_ROW + DICT_N * 48) + sizeof(int) * NE * THREADS;
  CK(cudaFuncSetAttribute(k_body<MODE, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int grid = 148;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    k_body<MODE, THREADS><<<grid, THREADS, smem>>>(d_fold, d_dict, d_bytes, d_codes, iters, d_sink);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
  }
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<double> h((size_t)grid * THREADS);
  CK(cudaMemcpy(h.data(), d_sink, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
  unsigned long long ck = 1469598103934665603ull;
  for (double x : h) { unsigned long long u; memcpy(&u, &x, 8); ck = (ck ^ u) * 1099511628211ull; }
  const double pairs = (double)grid * THREADS * iters;
  printf("%-44s %4d threads  %8.3f ms  %7.2f Gpairs/s  (config 2's 19.9 M pairs: %.3f ms)  checksum %016llx\n", name, THREADS, ms,
         pairs / ms * 1e-6, 19.86e6 / (pairs / ms), ck);
  (void)n_threads_max;
}

int main() {
  const int iters = 512;
  const size_t n_max = (size_t)148 * 384;
  std::vector<double> fold(FOLD_ROWS * FOLD_ROW), dict(DICT_N * 3);
  srand(7);
  for (auto& x : fold) x = 0.05 + 0.9 * (rand() / (double)RAND_MAX);
  for (int i = 0; i < FOLD_ROW; ++i) fold[(2 * 64) * FOLD_ROW + i] = 1.0;  // the "no read" row
  for (int s = 0; s < DICT_N; ++s) {
    double a = 0.01 + rand() / (double)RAND_MAX, b = 0.01 + rand() / (double)RAND_MAX, c = 0.01 + rand() / (double)RAND_MAX;
    dict[3 * s] = a / (a + b + c); dict[3 * s + 1] = b / (a + b + c); dict[3 * s + 2] = c / (a + b + c);
  }
  std::vector<uint32_t> bytes(n_max * iters);
  std::vector<unsigned long long> codes(n_max * iters);
  for (size_t i = 0; i < bytes.size(); ++i) {
    const uint32_t q0 = (rand() % 2) * 64 + 13 + rand() % 28;  // first base-call: allele 0/1, phred 13..40
    const bool one = rand() % 4 != 0;                        // 3/4 of the pairs have a single base-call
    const uint32_t q1 = one ? 2 * 64 : (rand() % 2) * 64 + 13 + rand() % 28, q2 = 2 * 64;
    bytes[i] = q0 | (q1 << 8) | (q2 << 16);
    unsigned long long c = 0;
    const int pal[3] = {rand() % DICT_N, rand() % DICT_N, rand() % DICT_N};  // hard calls: three triples per SNP
    for (int j = 0; j < NV; ++j) c |= (unsigned long long)pal[rand() % 3] << (8 * j);
    codes[i] = c;
  }
  double *d_fold, *d_dict, *d_sink;
  uint32_t* d_bytes;
  unsigned long long* d_codes;
  CK(cudaMalloc(&d_fold, fold.size() * 8)); CK(cudaMalloc(&d_dict, dict.size() * 8)); CK(cudaMalloc(&d_sink, n_max * 8));
  CK(cudaMalloc(&d_bytes, bytes.size() * 4)); CK(cudaMalloc(&d_codes, codes.size() * 8));
  CK(cudaMemcpy(d_fold, fold.data(), fold.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_dict, dict.data(), dict.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_bytes, bytes.data(), bytes.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_codes, codes.data(), codes.size() * 8, cudaMemcpyHostToDevice));
  run<0, 256>("0 kernel order (compiler's contraction)", d_fold, d_dict, d_bytes, d_codes, iters, d_sink, n_max);
  run<2, 256>("2 kernel order, explicit mul/fma", d_fold, d_dict, d_bytes, d_codes, iters, d_sink, n_max);
  run<1, 256>("1 four chains interleaved", d_fold, d_dict, d_bytes, d_codes, iters, d_sink, n_max);
  run<3, 256>("3 interleaved + next pair's front first", d_fold, d_dict, d_bytes, d_codes, iters, d_sink, n_max);
  run<1, 384>("1 four chains interleaved", d_fold, d_dict, d_bytes, d_codes, iters, d_sink, n_max);
  run<3, 384>("3 interleaved + next pair's front first", d_fold, d_dict, d_bytes, d_codes, iters, d_sink, n_max);
  return 0;
}
