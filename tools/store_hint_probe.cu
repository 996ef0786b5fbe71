// Sustained (power-capped) write bandwidth of the 256-bit cell store with different cache hints.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/store_hint_probe tools/store_hint_probe.cu && tools/store_hint_probe
#include <chrono>
#include <thread>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k_store(uint32_t* out, uint64_t n_cells, uint32_t salt) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += stride) {
    const uint32_t v = (uint32_t)i * 2654435761u + salt;
    if (MODE == 0) asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out + i * 8), "r"(v), "r"(v ^ 0x9E3779B1u), "r"(v + 0x85EBCA77u), "r"(v ^ 0xC2B2AE3Du), "r"(v + 0x27D4EB2Fu), "r"(v ^ 0x165667B1u), "r"(v + 0xD3A2646Du), "r"(v ^ 0xFD7046C5u) : "memory");
    if (MODE == 1) asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out + i * 8), "r"(v), "r"(v ^ 0x9E3779B1u), "r"(v + 0x85EBCA77u), "r"(v ^ 0xC2B2AE3Du), "r"(v + 0x27D4EB2Fu), "r"(v ^ 0x165667B1u), "r"(v + 0xD3A2646Du), "r"(v ^ 0xFD7046C5u) : "memory");
    if (MODE == 2) asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out + i * 8), "r"(v), "r"(v ^ 0x9E3779B1u), "r"(v + 0x85EBCA77u), "r"(v ^ 0xC2B2AE3Du), "r"(v + 0x27D4EB2Fu), "r"(v ^ 0x165667B1u), "r"(v + 0xD3A2646Du), "r"(v ^ 0xFD7046C5u) : "memory");
    if (MODE == 3) asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out + i * 8), "r"(v), "r"(v ^ 0x9E3779B1u), "r"(v + 0x85EBCA77u), "r"(v ^ 0xC2B2AE3Du), "r"(v + 0x27D4EB2Fu), "r"(v ^ 0x165667B1u), "r"(v + 0xD3A2646Du), "r"(v ^ 0xFD7046C5u) : "memory");
    if (MODE == 4) { asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(out + i * 8), "r"(v), "r"(v ^ 0x9E3779B1u), "r"(v + 0x85EBCA77u), "r"(v ^ 0xC2B2AE3Du) : "memory");
                     asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(out + i * 8 + 4), "r"(v + 0x27D4EB2Fu), "r"(v ^ 0x165667B1u), "r"(v + 0xD3A2646Du), "r"(v ^ 0xFD7046C5u) : "memory"); }
  }
}
template __global__ void k_store<0>(uint32_t*, uint64_t, uint32_t);
template __global__ void k_store<1>(uint32_t*, uint64_t, uint32_t);
template __global__ void k_store<2>(uint32_t*, uint64_t, uint32_t);
template __global__ void k_store<3>(uint32_t*, uint64_t, uint32_t);
template __global__ void k_store<4>(uint32_t*, uint64_t, uint32_t);

// CTA-contiguous streams: CTA b of G owns cells [b*n/G, (b+1)*n/G) and walks them front to back (what a persistent writer does)
__global__ void k_store_cta_contig(uint32_t* out, uint64_t n_cells, uint32_t salt) {
  const uint64_t per = (n_cells + gridDim.x - 1) / gridDim.x;
  const uint64_t lo = (uint64_t)blockIdx.x * per, hi = lo + per < n_cells ? lo + per : n_cells;
  for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const uint32_t v = (uint32_t)i * 2654435761u + salt;
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out + i * 8), "r"(v), "r"(v ^ 0x9E3779B1u), "r"(v + 0x85EBCA77u), "r"(v ^ 0xC2B2AE3Du), "r"(v + 0x27D4EB2Fu), "r"(v ^ 0x165667B1u), "r"(v + 0xD3A2646Du), "r"(v ^ 0xFD7046C5u) : "memory");
  }
}
double run_contig(uint32_t* buf, uint64_t cells, int ctas, int threads, bool sustained) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int k = 0;
  if (sustained) {
    auto t0 = std::chrono::steady_clock::now();
    while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 2.5) { for (int i = 0; i < 20; i++) k_store_cta_contig<<<ctas, threads>>>(buf, cells, k++); cudaDeviceSynchronize(); }
  } else {
    k_store_cta_contig<<<ctas, threads>>>(buf, cells, k++); cudaDeviceSynchronize();
    std::this_thread::sleep_for(std::chrono::milliseconds(500));
  }
  const int n = sustained ? 50 : 1;
  cudaEventRecord(a);
  for (int i = 0; i < n; i++) k_store_cta_contig<<<ctas, threads>>>(buf, cells, k++);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return cells * 32.0 * n / (ms * 1e-3) / 1e9;
}

template <int MODE> double run(uint32_t* buf, uint64_t cells) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto t0 = std::chrono::steady_clock::now();
  int k = 0;
  while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 2.5) { for (int i = 0; i < 20; i++) k_store<MODE><<<148 * 8, 256>>>(buf, cells, k++); cudaDeviceSynchronize(); }
  cudaEventRecord(a);
  for (int i = 0; i < 50; i++) k_store<MODE><<<148 * 8, 256>>>(buf, cells, k++);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return cells * 32.0 * 50 / (ms * 1e-3) / 1e9;
}
int main() {
  const uint64_t bytes = 2680ull << 20; uint32_t* buf; cudaMalloc(&buf, bytes);
  const uint64_t cells = bytes / 32;
  const char* names[] = {"st.global.v8 (default)", "st.global.cs.v8", "st.global.L1::no_allocate.v8", "st.global.L2::evict_first.v8", "2 x st.global.v4"};
  for (int rep = 0; rep < 1; rep++) {
    printf("%-32s %7.1f GB/s sustained\n", names[0], run<0>(buf, cells));
    printf("%-32s %7.1f GB/s sustained\n", names[1], run<1>(buf, cells));
    printf("%-32s %7.1f GB/s sustained\n", names[2], run<2>(buf, cells));
    printf("%-32s %7.1f GB/s sustained\n", names[3], run<3>(buf, cells));
    printf("%-32s %7.1f GB/s sustained\n", names[4], run<4>(buf, cells));
  }
  for (int threads : {256, 640, 1024})
    for (int per_sm : {1, 2, 4}) {
      if (threads * per_sm > 2048) continue;
      printf("CTA-contiguous, %d CTAs x %4d threads: %7.1f GB/s single launch after idle, %7.1f GB/s sustained\n", 148 * per_sm, threads,
             run_contig(buf, cells, 148 * per_sm, threads, false), run_contig(buf, cells, 148 * per_sm, threads, true));
    }
  return 0;
}
