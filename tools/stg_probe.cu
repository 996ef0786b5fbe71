// Micro-probe: LSU data-pipe wavefronts per byte for 128-bit vs 256-bit coalesced global stores (sm_100a).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_stg256(uint32_t* out, size_t n_cells) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n_cells; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t v = (uint32_t)i;
    asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(out + i * 8), "r"(v) : "memory");
  }
}
__global__ void k_stg128(uint32_t* out, size_t n_cells) {   // lane pair per cell: each store instruction writes 512 contiguous bytes
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < 2 * n_cells; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t v = (uint32_t)i;
    asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(out + i * 4), "r"(v) : "memory");
  }
}
__global__ void k_stg128x2(uint32_t* out, size_t n_cells) {  // one lane writes both halves of its cell (two instructions, stride 32 B)
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n_cells; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t v = (uint32_t)i;
    asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(out + i * 8), "r"(v) : "memory");
    asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(out + i * 8 + 4), "r"(v) : "memory");
  }
}
int main() {
  size_t n_cells = (size_t)1 << 26;  // 2 GiB
  uint32_t* d; cudaMalloc(&d, n_cells * 32);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int k = 0; k < 3; k++) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(a);
      if (k == 0) k_stg256<<<148 * 8, 256>>>(d, n_cells);
      if (k == 1) k_stg128<<<148 * 8, 256>>>(d, n_cells);
      if (k == 2) k_stg128x2<<<148 * 8, 256>>>(d, n_cells);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep) printf("%s: %.3f ms  %.1f GB/s\n", k == 0 ? "stg256" : k == 1 ? "stg128 coalesced" : "stg128 x2 strided", ms, n_cells * 32 / ms / 1e6);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
