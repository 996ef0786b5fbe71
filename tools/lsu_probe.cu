// Micro-probe (sm_100a): what the LSU charges for (a) 256-bit global stores whose lanes cover whole 128-byte lines in a
// permuted order, at 128 / 64 / 32-byte granularity, and (b) 128-bit shared loads whose bank conflicts are inside a
// quarter-warp vs. spread over the warp.  The planner's scratch colouring and cell order rely on these two answers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lsu_probe tools/lsu_probe.cu && tools/lsu_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

// mode 0: linear; 1: lines (4 cells) permuted inside a 512-cell window; 2: 64-byte pairs permuted; 3: single cells permuted;
// 4: linear but shifted by one cell (every 4-lane group straddles two lines)
__global__ void k_store(uint32_t* out, size_t n_cells, int mode) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + 512 < n_cells; i += stride) {
    const size_t win = i & ~(size_t)511;
    const uint32_t k = (uint32_t)(i & 511);
    uint32_t c = k;
    if (mode == 1) c = (((k >> 2) * 37u) & 127u) * 4u + (k & 3u);
    if (mode == 2) c = (((k >> 1) * 101u) & 255u) * 2u + (k & 1u);
    if (mode == 3) c = (k * 201u) & 511u;
    if (mode == 4) c = k + 1;
    const uint32_t v = (uint32_t)i;
    asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(out + (win + c) * 8), "r"(v) : "memory");
  }
}

// mode 0: every quarter-warp reads 8 different bank groups; 1: inside a quarter lanes pair up on 4 bank groups (different
// addresses) while over the whole warp every bank group is still used 4 times; 2: all lanes of a quarter on one bank group
__global__ void k_lds(uint32_t* out, int iters, int mode) {
  __shared__ uint4 tab[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = make_uint4(i, i * 3, i * 5, i * 7);
  __syncthreads();
  const int lane = threadIdx.x & 31, q = lane >> 3, l8 = lane & 7;
  int idx;
  if (mode == 0) idx = l8 + 8 * q;                                    // bank group l8, distinct in the quarter
  else if (mode == 1) idx = ((l8 >> 1) + 4 * (q & 1)) + 8 * (l8 & 1) + 16 * q;   // two lanes per bank group, different rows
  else idx = 8 * l8 + 64 * q;                                          // bank group 0 for every lane
  uint32_t acc = 0;
  int off = (threadIdx.x >> 5) * 8;
  for (int it = 0; it < iters; it++) {
    const uint4 v = tab[(idx + off) & 1023];
    acc ^= v.x + v.y + v.z + v.w;
    off = (off + 8 * (int)(acc & 1u) + 8) & 1023;   // keeps the bank group, defeats hoisting
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// copy loop like the engine's: NL conflict-free 128-bit shared loads + one 256-bit store per lane and iteration; the cell a
// lane stores to follows pattern `mode` inside 512-cell windows
__device__ __forceinline__ uint32_t pattern(uint32_t k, int mode) {
  switch (mode) {
    case 1: return k + 1;                                          // linear, lines misaligned by one cell
    case 2: return k + 4;                                          // linear, line-aligned but not 256-byte aligned
    case 3: return (((k >> 2) * 37u) & 127u) * 4u + (k & 3u);      // whole lines permuted
    case 4: return (((k >> 3) * 37u) & 63u) * 8u + (k & 7u);       // 256-byte groups permuted
    case 5: return (((k >> 1) * 101u) & 255u) * 2u + (k & 1u);     // 64-byte pairs permuted
    case 6: return (k * 201u) & 511u;                              // cells permuted
    case 7: return (k & ~31u) | ((k * 13u) & 31u);                 // cells permuted inside each aligned 1 KB
    case 8: return (k & ~7u) | ((k * 5u) & 7u);                    // cells permuted inside each aligned 256 B
    case 9: return (k & ~31u) | ((((k >> 2) * 5u) & 7u) * 4u) | (k & 3u);   // lines permuted inside each aligned 1 KB
    default: return k;
  }
}
template <int NL>
__global__ void k_copy(uint32_t* out, size_t n_cells, int mode) {
  __shared__ uint4 tab[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) tab[i] = make_uint4(i, i * 3, i * 5, i * 7);
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  uint32_t row = (threadIdx.x >> 5) * 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + 1024 < n_cells; i += stride) {
    const size_t win = i & ~(size_t)511;
    const uint32_t c = pattern((uint32_t)(i & 511), mode);
    uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
#pragma unroll
    for (int l = 0; l < NL; l += 2) {
      const uint4 a = tab[(row + (lane & 7) + 8 * l) & 2047], b = tab[(row + (lane & 7) + 8 * l + 8) & 2047];
      lo.x ^= a.x; lo.y ^= a.y; lo.z ^= a.z; lo.w ^= a.w; hi.x ^= b.x; hi.y ^= b.y; hi.z ^= b.z; hi.w ^= b.w;
    }
    row = (row + 64 + (lo.x & 8u)) & 2047;
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out + (win + c) * 8), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w),
                 "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
  }
}

int main() {
  size_t n_cells = (size_t)1 << 26;  // 2 GiB of 32-byte cells
  uint32_t* d;
  cudaMalloc(&d, n_cells * 32);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const char* sn[] = {"linear", "line-permuted (128 B)", "pair-permuted (64 B)", "cell-permuted (32 B)", "linear, shifted one cell"};
  for (int m = 0; m < 5; m++)
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(a);
      k_store<<<148 * 8, 256>>>(d, n_cells, m);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep) printf("stg256 %-28s %.3f ms  %.1f GB/s\n", sn[m], ms, n_cells * 32 / ms / 1e6);
    }
  // sustained regime (power cap): incompressible linear stores back to back for ~2 s, then 20 timed launches
  {
    for (int r = 0; r < 5000; r++) k_store<<<148 * 8, 256>>>(d, n_cells, 0);
    cudaEventRecord(a);
    for (int r = 0; r < 20; r++) k_store<<<148 * 8, 256>>>(d, n_cells, 0);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("stg256 linear, sustained (after 5000 launches)   %.3f ms  %.1f GB/s\n", ms / 20, 20.0 * n_cells * 32 / ms / 1e6);
  }
  // the same stores into a 32 MiB window that stays in L2: exposes the LSU / L2 cost of each pattern without the HBM bound
  for (int m = 0; m < 5; m++)
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(a);
      for (int r = 0; r < 64; r++) k_store<<<148 * 8, 256>>>(d, (size_t)1 << 20, m);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep) printf("stg256 L2-resident %-28s %.3f ms  %.1f GB/s\n", sn[m], ms, 64.0 * (1 << 20) * 32 / ms / 1e6);
    }
  const char* cn[] = {"linear", "linear +1 cell", "linear +4 cells", "lines permuted", "256 B groups permuted", "64 B pairs permuted", "cells permuted",
                      "cells permuted inside 1 KB", "cells permuted inside 256 B", "lines permuted inside 1 KB"};
  for (int nl = 2; nl <= 6; nl += 2)
    for (int m = 0; m < 10; m++)
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(a);
        if (nl == 2) k_copy<2><<<148 * 2, 768>>>(d, n_cells, m);
        if (nl == 4) k_copy<4><<<148 * 2, 768>>>(d, n_cells, m);
        if (nl == 6) k_copy<6><<<148 * 2, 768>>>(d, n_cells, m);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep) printf("copy lds128 x%d + stg256 %-30s %.3f ms  %.1f GB/s\n", nl, cn[m], ms, n_cells * 32 / ms / 1e6);
      }
  const char* ln[] = {"conflict-free quarters", "2-way inside quarters, balanced over the warp", "8-way inside quarters"};
  for (int m = 0; m < 3; m++)
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(a);
      k_lds<<<148 * 2, 512>>>(d, 20000, m);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep) printf("lds128 %-48s %.3f ms\n", ln[m], ms);
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
