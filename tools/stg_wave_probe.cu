// Micro-probe (sm_100a), run under ncu: LSU data-pipe wavefronts of ONE 256-bit store instruction per warp as a function of
// how the 32 lanes map onto an aligned 1 KB run of 32-byte cells.
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum tools/stg_wave_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t pattern(uint32_t k, int mode) {
  const uint32_t q = k >> 3, l8 = k & 7, ln = k >> 2, l4 = k & 3;
  switch (mode) {
    case 1: return (((ln * 5u) & 7u) << 2) | l4;                 // lines permuted, lanes in order inside a line
    case 2: return (ln << 2) | ((l4 * 3u) & 3u);                  // lanes permuted inside each line
    case 3: return (q << 3) | ((l8 * 5u) & 7u);                   // lanes permuted inside each 256 B
    case 4: return (k * 13u) & 31u;                               // lanes permuted inside the 1 KB
    case 5: return (((l8 >> 2) ? q + 4 : q) << 2) | l4;            // quarter q = lines q and q+4 (4 consecutive lanes each)
    case 6: return ((((l8 & 1) ? q + 4 : q)) << 2) | (l8 >> 1);   // quarter q = lines q and q+4, lanes interleaved
    case 7: return k + 1;                                         // linear, shifted by one cell
    case 8: return k + 4;                                         // linear, shifted by one line
    case 9: return (((q * 3u) & 3u) << 3) | l8;                   // 256 B groups permuted
    case 10: return (ln << 2) | (l4 ^ 1u);                        // 64 B halves of each line swapped... (pairs of lanes swapped)
    case 11: return (k & ~1u) | ((k & 1u) ^ 1u);                  // adjacent lanes swapped
    default: break;
  }
  if (mode >= 20) {   // quarter q stores lines PAIRS[mode-20][q][0] (lanes 0-3) and [1] (lanes 4-7)
    static const uint8_t P[][4][2] = {
        {{0, 2}, {1, 3}, {4, 6}, {5, 7}}, {{0, 7}, {1, 6}, {2, 5}, {3, 4}}, {{1, 0}, {3, 2}, {5, 4}, {7, 6}}, {{0, 3}, {1, 2}, {4, 7}, {5, 6}},
        {{0, 6}, {1, 7}, {2, 4}, {3, 5}}, {{0, 5}, {1, 4}, {2, 7}, {3, 6}}, {{4, 0}, {5, 1}, {6, 2}, {7, 3}}, {{0, 5}, {2, 7}, {4, 1}, {6, 3}},
        {{0, 4}, {2, 6}, {1, 5}, {3, 7}}, {{0, 1}, {4, 5}, {2, 3}, {6, 7}}, {{0, 2}, {4, 6}, {1, 3}, {5, 7}}, {{3, 6}, {0, 5}, {2, 7}, {1, 4}}};
    return ((uint32_t)P[mode - 20][q][l8 >> 2] << 2) | l4;
  }
  return k;
}
__global__ void k_store(uint32_t* out, int mode) {
  const uint32_t lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t c = pattern(lane, mode);
  const uint32_t v = (uint32_t)warp;
  asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(out + (warp * 64 + c) * 8), "r"(v) : "memory");
}
int main() {
  uint32_t* d;
  cudaMalloc(&d, (size_t)1 << 28);
  for (int m = 0; m < 12; m++) k_store<<<1024, 256>>>(d, m);
  for (int m = 20; m < 32; m++) k_store<<<1024, 256>>>(d, m);   // 8192 warps = 8192 store instructions per launch
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
