// Micro-probe (sm_100a) behind the round-2 store path of k_expand: is it cheaper to (A) read every cell's value back from a
// shared-memory scratch (2 x LDS.128) and store it with one STG.256 per lane -- the round-1 copy loop -- or (B) to scatter the
// values into a shared-memory tile that has the output's layout (2 x STS.128 per cell) and hand the tile to the bulk-copy engine
// (cp.async.bulk.global.shared::cta, SASS UBLKCP), which keeps the global store off the LSU pipe?  Reports burst and sustained
// (power-capped) GB/s for 2.68 GB per launch (one cfg2 launch of the engine) with 768 threads per CTA, one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tile_probe tools/tile_probe.cu -ldl && tools/tile_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <dlfcn.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// (A) round-1 copy loop: 1 x LDS.32 (cell entry) + 2 x LDS.128 (scratch) + STG.256 per cell, cells permuted inside aligned 1 KB
__global__ void __launch_bounds__(768, 1) k_copy_stg(uint32_t* out, size_t n_cells, int work) {
  __shared__ uint4 lo[1024], hi[1024];
  __shared__ uint32_t ent[2048];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) { lo[i] = make_uint4(i, i * 3, i * 5, i * 7); hi[i] = make_uint4(i * 11, i * 13, i * 17, i * 19); }
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) ent[i] = (i & ~31u) | ((i * 13u) & 31u);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t n_tiles = n_cells / 128;
  for (size_t t = (size_t)blockIdx.x * nw + warp; t < n_tiles; t += (size_t)gridDim.x * nw) {
    uint32_t* dst = out + t * 128 * 8;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const uint32_t e = ent[(warp * 128 + r * 32 + lane) & 2047];
      const uint32_t src = (e * 7u + (uint32_t)t) & 1023u;
      uint4 a = lo[(src & ~7u) | (lane & 7)], b = hi[(src & ~7u) | (lane & 7)];
      for (int w = 0; w < work; w++) { a.x = a.x * 2654435761u + b.y; b.x ^= a.x >> 3; }
      const uint32_t c = (e & 127u);
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + c * 8), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
                   "r"(b.z), "r"(b.w) : "memory");
    }
  }
}

// (B) value-major scatter into a per-warp tile + bulk copy.  mode 0: lanes alternate low / high halves so that a quarter-warp
// covers 8 bank groups; mode 1: all lanes write the low half, then the high half (only 4 of the 8 bank groups per instruction).
template <int TILE_CELLS>
__global__ void __launch_bounds__(768, 1) k_tile_bulk(uint32_t* out, size_t n_cells, int mode, int work) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t ent[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) ent[i] = (i & ~31u) | ((i * 13u) & 31u);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint8_t* tiles = smem + (size_t)warp * 2 * TILE_CELLS * 32;
  const size_t n_tiles = n_cells / TILE_CELLS;
  uint32_t buf = 0;
  for (size_t t = (size_t)blockIdx.x * nw + warp; t < n_tiles; t += (size_t)gridDim.x * nw, buf ^= 1u) {
    uint8_t* tile = tiles + buf * TILE_CELLS * 32;
    // the bulk copy that read this buffer two iterations ago must have finished READING it
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int r = 0; r < TILE_CELLS / 32; r++) {
      const uint32_t e = ent[(warp * 128 + r * 32 + lane) & 2047];
      uint4 a = make_uint4(e + (uint32_t)t, e * 3u, e * 5u, e * 7u), b = make_uint4(e * 11u, e * 13u + (uint32_t)t, e * 17u, e * 19u);
      for (int w = 0; w < work; w++) { a.x = a.x * 2654435761u + b.y; b.x ^= a.x >> 3; }
      const uint32_t c = (e & 31u) + 32u * r;
      uint4* p = reinterpret_cast<uint4*>(tile + c * 32);
      if (mode == 0) {
        // lane parity decides which half goes first: within a quarter-warp 4 lanes hit even and 4 odd bank groups
        const bool odd = lane & 1;
        p[odd ? 1 : 0] = odd ? b : a;
        p[odd ? 0 : 1] = odd ? a : b;
      } else {
        p[0] = a;
        p[1] = b;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + t * TILE_CELLS * 8), "r"(smem_u32(tile)), "r"(TILE_CELLS * 32) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// (C) value-major direct stores: the lane that holds a value stores it (no shared memory at all); cells permuted inside 1 KB
__global__ void __launch_bounds__(768, 1) k_direct(uint32_t* out, size_t n_cells, int work) {
  __shared__ uint32_t ent[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) ent[i] = (i & ~31u) | ((i * 13u) & 31u);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t n_tiles = n_cells / 128;
  for (size_t t = (size_t)blockIdx.x * nw + warp; t < n_tiles; t += (size_t)gridDim.x * nw) {
    uint32_t* dst = out + t * 128 * 8;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const uint32_t e = ent[(warp * 128 + r * 32 + lane) & 2047];
      uint4 a = make_uint4(e + (uint32_t)t, e * 3u, e * 5u, e * 7u), b = make_uint4(e * 11u, e * 13u + (uint32_t)t, e * 17u, e * 19u);
      for (int w = 0; w < work; w++) { a.x = a.x * 2654435761u + b.y; b.x ^= a.x >> 3; }
      const uint32_t c = (e & 31u) + 32u * r;
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + c * 8), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
                   "r"(b.z), "r"(b.w) : "memory");
    }
  }
}

// plain linear STG.256 writer (the store ceiling of round 1)
__global__ void __launch_bounds__(768, 1) k_plain(uint32_t* out, size_t n_cells) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += stride) {
    const uint32_t v = (uint32_t)i * 2654435761u;
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out + i * 8), "r"(v), "r"(v ^ 0x9E3779B1u), "r"(v + 0x85EBCA77u), "r"(v ^ 0xC2B2AE3Du),
                 "r"(v + 0x27D4EB2Fu), "r"(v ^ 0x165667B1u), "r"(v + 0xD3A2646Du), "r"(v ^ 0xFD7046C5u) : "memory");
  }
}

// power / SM clock through NVML (loaded at run time; absent -> zeros)
typedef int (*nvml_init_t)();
typedef int (*nvml_handle_t)(unsigned, void**);
typedef int (*nvml_power_t)(void*, unsigned*);
typedef int (*nvml_clock_t)(void*, int, unsigned*);
static void* g_dev = nullptr;
static nvml_power_t g_power = nullptr;
static nvml_clock_t g_clock = nullptr;
static void nvml_setup() {
  void* h = dlopen("libnvidia-ml.so.1", RTLD_NOW);
  if (!h) return;
  nvml_init_t init = (nvml_init_t)dlsym(h, "nvmlInit_v2");
  nvml_handle_t get = (nvml_handle_t)dlsym(h, "nvmlDeviceGetHandleByIndex_v2");
  g_power = (nvml_power_t)dlsym(h, "nvmlDeviceGetPowerUsage");
  g_clock = (nvml_clock_t)dlsym(h, "nvmlDeviceGetClockInfo");
  if (!init || !get || init() != 0 || get(0, &g_dev) != 0) g_dev = nullptr;
}

template <class F>
void run(const char* name, F launch, size_t bytes) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int r = 0; r < 3; r++) launch();
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
  }
  // sustained: ~1.5 s back to back, then 50 timed launches
  const int warm = (int)(1500.0f / best);
  for (int r = 0; r < warm; r++) launch();
  cudaEventRecord(a);
  for (int r = 0; r < 200; r++) launch();
  cudaEventRecord(b);
  unsigned pw = 0, mhz = 0, n = 0;
  while (cudaEventQuery(b) == cudaErrorNotReady) {
    unsigned x = 0, y = 0;
    if (g_dev && g_power && g_power(g_dev, &x) == 0 && g_clock && g_clock(g_dev, 1 /*NVML_CLOCK_SM*/, &y) == 0) { pw += x / 1000; mhz += y; n++; }
  }
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  printf("%-58s burst %.4f ms %7.1f GB/s | sustained %.4f ms %7.1f GB/s  %4u W %4u MHz %s\n", name, best, bytes / best / 1e6, ms / 200, 200.0 * bytes / ms / 1e6,
         n ? pw / n : 0, n ? mhz / n : 0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  fflush(stdout);
}

int main() {
  const size_t n_cells = (size_t)1024 * 81774 / 128 * 128;   // one cfg2 launch
  const size_t bytes = n_cells * 32;
  uint32_t* d;
  cudaMalloc(&d, bytes + 4096);
  nvml_setup();
  cudaFuncSetAttribute(k_tile_bulk<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_tile_bulk<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  run("plain STG.256 linear", [&] { k_plain<<<148, 768>>>(d, n_cells); }, bytes);
  for (int work = 0; work <= 64; work += 32) {
    char nm[128];
    snprintf(nm, sizeof nm, "A: LDS.32 + 2xLDS.128 + STG.256 (work %d)", work);
    run(nm, [&] { k_copy_stg<<<148, 768>>>(d, n_cells, work); }, bytes);
    snprintf(nm, sizeof nm, "C: direct STG.256 from registers (work %d)", work);
    run(nm, [&] { k_direct<<<148, 768>>>(d, n_cells, work); }, bytes);
    snprintf(nm, sizeof nm, "B: 2xSTS.128 alternating halves + bulk 4 KB (work %d)", work);
    run(nm, [&] { k_tile_bulk<128><<<148, 768, 24 * 2 * 128 * 32>>>(d, n_cells, 0, work); }, bytes);
    snprintf(nm, sizeof nm, "B: 2xSTS.128 low then high + bulk 4 KB (work %d)", work);
    run(nm, [&] { k_tile_bulk<128><<<148, 768, 24 * 2 * 128 * 32>>>(d, n_cells, 1, work); }, bytes);
    snprintf(nm, sizeof nm, "B: 2xSTS.128 alternating halves + bulk 2 KB (work %d)", work);
    run(nm, [&] { k_tile_bulk<64><<<148, 768, 24 * 2 * 64 * 32>>>(d, n_cells, 0, work); }, bytes);
  }
  return 0;
}
