// native_runner.cc -- the whole path with no Python and no PyTorch: C++ host code -> C-ABI (include/h2sha_b200.h) -> CUDA,
// one host thread per GPU, NCCL only to gather digests + cell checksums after the hot path.
//
// This is what a Rust host (cudarc + bindgen over the same header, rust/src/lib.rs) does; C++ because the image has no
// cargo.  It mirrors bench.py's timed legs:
//   resident : inputs already in HBM, K back-to-back h2sha_digest_batch calls per GPU, CUDA events, max over GPUs
//   e2e      : pinned host message buffers in, digests + checksums out to pinned host memory, every step
// and checks every digest of every GPU against a host SHA-256 (FIPS 180-4, below) after the all-gather.
//
//   nvcc -std=c++17 -O2 -o tools/native_runner tools/native_runner.cc -Lhalo2-dynamic-sha256_b200 -lh2sha_b200 -lnccl \
//        -Xlinker -rpath,$PWD/halo2-dynamic-sha256_b200 -cudart shared
//   tools/native_runner --workload cfg2 --gpus 2 --steps 20 --warmup 5
#include <cuda_runtime.h>
#include <nccl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../halo2-dynamic-sha256_b200/csrc/host_api.hpp"

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } \
  } while (0)
#define NK(x)                                                                                   \
  do {                                                                                          \
    ncclResult_t r_ = (x);                                                                      \
    if (r_ != ncclSuccess) { fprintf(stderr, "%s: %s\n", #x, ncclGetErrorString(r_)); exit(2); } \
  } while (0)

// ---- synthetic workloads: same counter-based generator as halo2-dynamic-sha256_b200/synthetic.py ----
static const uint64_t SEED = 0x5348413235360001ULL, GOLDEN = 0x9E3779B97F4A7C15ULL, LEN_TAG = 0xA5A5A5A55A5A5A5AULL;
static uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + GOLDEN;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
struct Workload {
  const char* name;
  uint32_t max_bytes;
  uint64_t n_instances;
  uint32_t len_lo, len_hi;
};
static const Workload WORKLOADS[] = {{"cfg1", 128, 1, 64, 64},
                                     {"cfg2", 64, 1024, 55, 55},
                                     {"cfg3", 1088, 4096, 0, 1024},
                                     {"cfg4", 320, 1u << 16, 256, 256},
                                     {"cfg5", 2112, 1u << 18, 1, 2048}};
static void generate(const Workload& w, uint64_t first, uint64_t count, std::vector<uint8_t>* blob, std::vector<uint64_t>* offs, std::vector<uint32_t>* lens) {
  blob->clear(); offs->resize(count); lens->resize(count);
  for (uint64_t i = 0; i < count; i++) {
    const uint64_t m = first + i;
    uint32_t len = w.len_lo;
    if (w.len_lo != w.len_hi) len = w.len_lo + (uint32_t)(splitmix64((SEED ^ LEN_TAG) + m) % (w.len_hi - w.len_lo + 1));
    (*lens)[i] = len; (*offs)[i] = blob->size();
    const uint64_t key = SEED + m * GOLDEN;
    for (uint32_t j = 0; j < len; j += 8) {
      const uint64_t v = splitmix64(key + j / 8);
      for (uint32_t b = 0; b < 8 && j + b < len; b++) blob->push_back((uint8_t)(v >> (8 * b)));
    }
  }
}

// ---- host SHA-256 (FIPS 180-4) for the self-check ----
static void sha256(const uint8_t* msg, size_t len, uint8_t out[32]) {
  static const uint32_t K[64] = {
      0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
      0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
      0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
      0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
      0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
      0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  std::vector<uint8_t> p(msg, msg + len);
  p.push_back(0x80);
  while (p.size() % 64 != 56) p.push_back(0);
  for (int i = 7; i >= 0; i--) p.push_back((uint8_t)(((uint64_t)len * 8) >> (8 * i)));
  auto rotr = [](uint32_t x, int n) { return (x >> n) | (x << (32 - n)); };
  for (size_t b = 0; b < p.size(); b += 64) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[b + 4 * i] << 24 | (uint32_t)p[b + 4 * i + 1] << 16 | (uint32_t)p[b + 4 * i + 2] << 8 | p[b + 4 * i + 3];
    for (int i = 16; i < 64; i++) {
      const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      const uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
      const uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & bb) ^ (a & c) ^ (bb & c));
      hh = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
    }
    h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16); out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i]; }
}

struct Shared {
  const Workload* w;
  int n_gpus, steps, warmup;
  uint64_t per_gpu;
  pthread_barrier_t bar;
  std::vector<ncclComm_t> comms;
  std::vector<float> resident_ms, e2e_ms;
  std::vector<int> bad_digests;
  std::vector<uint64_t> job_checksum, violations;
  h2sha_layout_t layout{};
};

static void worker(Shared* S, int g) {
  CK(cudaSetDevice(g));
  const Workload& w = *S->w;
  auto sha = h2sha::Sha256DynamicConfig::configure({w.max_bytes}, 0, 0, 0, 0, true, g);   // 0 -> the reference's defaults (k = 17, 16-bit range table, 8-bit limbs, 2 column pairs)
  const h2sha_layout_t lay = sha.layout();
  if (g == 0) S->layout = lay;
  const uint64_t n = S->per_gpu, first = (uint64_t)g * n;
  std::vector<uint8_t> blob; std::vector<uint64_t> offs; std::vector<uint32_t> lens;
  generate(w, first, n, &blob, &offs, &lens);
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  void *gate, *lookup, *spread; uint8_t *d_blob, *d_dig; uint64_t* d_ck;
  CK(cudaMalloc(&gate, n * lay.gate_bytes)); CK(cudaMalloc(&lookup, n * lay.lookup_bytes)); CK(cudaMalloc(&spread, n * lay.spread_bytes));
  CK(cudaMemset(gate, 0, n * lay.gate_bytes)); CK(cudaMemset(lookup, 0, n * lay.lookup_bytes)); CK(cudaMemset(spread, 0, n * lay.spread_bytes));
  // gather buffers: [n_gpus][n] digests (32 B) and checksums (4 x u64); this GPU's slice is written in place
  CK(cudaMalloc(&d_dig, (size_t)S->n_gpus * n * 32)); CK(cudaMalloc(&d_ck, (size_t)S->n_gpus * n * 32));
  CK(cudaMalloc(&d_blob, blob.size() + 16)); CK(cudaMemcpy(d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  uint8_t *h_blob, *h_dig; uint64_t* h_ck;
  CK(cudaMallocHost(&h_blob, blob.size() + 16)); memcpy(h_blob, blob.data(), blob.size());
  CK(cudaMallocHost(&h_dig, n * 32)); CK(cudaMallocHost(&h_ck, n * 32));

  h2sha_batch_t b{};
  b.n_instances = n; b.msgs = d_blob; b.msgs_on_device = 1; b.msgs_bytes = blob.size(); b.offsets = offs.data(); b.lens = lens.data();
  b.gate = gate; b.lookup = lookup; b.spread = spread; b.digests_dev = d_dig + (size_t)g * n * 32; b.checksums_dev = d_ck + (size_t)g * n * 4; b.stream = st;
  sha.digest_batch(b);
  h2sha_batch_t r = b; r.reuse_inputs = 1;
  for (int i = 0; i < std::max(S->warmup, 3); i++) sha.digest_batch(r);
  CK(cudaStreamSynchronize(st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  pthread_barrier_wait(&S->bar);
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < S->steps; i++) sha.digest_batch(r);
  CK(cudaEventRecord(e1, st));
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&S->resident_ms[g], e0, e1));
  pthread_barrier_wait(&S->bar);

  // end to end: host buffers in, digests + checksums out, every step
  h2sha_batch_t e = b;
  e.msgs = h_blob; e.msgs_on_device = 0; e.digests_dev = nullptr; e.checksums_dev = nullptr; e.digests_host = h_dig; e.checksums_host = h_ck;
  for (int i = 0; i < 3; i++) { sha.digest_batch(e); CK(cudaStreamSynchronize(st)); }
  pthread_barrier_wait(&S->bar);
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < S->steps; i++) { sha.digest_batch(e); CK(cudaStreamSynchronize(st)); }
  CK(cudaEventRecord(e1, st));
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&S->e2e_ms[g], e0, e1));
  pthread_barrier_wait(&S->bar);

  // the only collective: all-gather of digests and checksums (64 B per instance) over NCCL
  sha.digest_batch(r);
  if (S->n_gpus > 1) {
    // h2sha_gather: ncclAllGather of this rank's slice into the rank-major buffers, over the communicator created in main()
    if (h2sha_gather(S->comms[g], n, 1, d_dig + (size_t)g * n * 32, d_ck + (size_t)g * n * 4, d_dig, d_ck, st) != H2SHA_OK) {
      fprintf(stderr, "h2sha_gather: %s\n", h2sha_last_error());
      exit(2);
    }
  }
  CK(cudaStreamSynchronize(st));
  // MockProver-style pass over every instance of this GPU's shard (gates, copies, lookups, digest bytes), on the device
  {
    uint64_t v[5];
    sha.check_batch(n, gate, lookup, spread, d_dig + (size_t)g * n * 32, v, st);
    S->violations[g] = v[0] + v[1] + v[2] + v[3] + v[4];
  }
  // every rank now holds every digest: rank g checks the shard of rank (g + 1) % n_gpus against the host SHA-256
  {
    const int src = (g + 1) % S->n_gpus;
    std::vector<uint8_t> all((size_t)S->n_gpus * n * 32);
    std::vector<uint64_t> cks((size_t)S->n_gpus * n * 4);
    CK(cudaMemcpy(all.data(), d_dig, all.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cks.data(), d_ck, cks.size() * 8, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> sb; std::vector<uint64_t> so; std::vector<uint32_t> sl;
    generate(w, (uint64_t)src * n, n, &sb, &so, &sl);
    int bad = 0;
    const uint64_t stride = std::max<uint64_t>(1, n / 256);
    for (uint64_t i = 0; i < n; i += stride) {
      uint8_t want[32];
      sha256(sb.data() + so[i], sl[i], want);
      if (memcmp(want, &all[((size_t)src * n + i) * 32], 32) != 0) bad++;
    }
    S->bad_digests[g] = bad;
    uint64_t ck = 0;
    for (size_t i = 0; i < cks.size(); i += 4) ck += cks[i + 3] * (2 * (i / 4) + 1);   // position-weighted sum of the per-instance totals
    S->job_checksum[g] = ck;
  }
  cudaFree(gate); cudaFree(lookup); cudaFree(spread); cudaFree(d_dig); cudaFree(d_ck); cudaFree(d_blob);
  cudaFreeHost(h_blob); cudaFreeHost(h_dig); cudaFreeHost(h_ck);
}

int main(int argc, char** argv) {
  std::string wl = "cfg2";
  int gpus = 1, steps = 20, warmup = 5;
  uint64_t instances = 0;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string k = argv[i];
    if (k == "--workload") wl = argv[i + 1];
    else if (k == "--gpus") gpus = atoi(argv[i + 1]);
    else if (k == "--steps") steps = atoi(argv[i + 1]);
    else if (k == "--warmup") warmup = atoi(argv[i + 1]);
    else if (k == "--instances") instances = strtoull(argv[i + 1], nullptr, 10);
    else { fprintf(stderr, "unknown option %s\n", k.c_str()); return 2; }
  }
  const Workload* w = nullptr;
  for (const Workload& x : WORKLOADS) if (wl == x.name) w = &x;
  if (!w) { fprintf(stderr, "unknown workload %s\n", wl.c_str()); return 2; }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < gpus) { fprintf(stderr, "needs %d CUDA device(s); the engine has no CPU path\n", gpus); return 3; }

  Shared S;
  S.w = w; S.n_gpus = gpus; S.steps = steps; S.warmup = warmup;
  // instances per GPU: the whole workload when it fits in ~60 % of HBM, else a shard of that size (like bench.py)
  {
    CK(cudaSetDevice(0));
    auto probe = h2sha::Sha256DynamicConfig::configure({w->max_bytes}, 0, 0, 0, 0, true, 0);
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const h2sha_layout_t& l = probe.layout();
    const uint64_t cap = std::max<uint64_t>(1, (uint64_t)(0.6 * free_b) / (l.gate_bytes + l.lookup_bytes + l.spread_bytes));
    S.per_gpu = std::min<uint64_t>(instances ? instances : w->n_instances, cap);
  }
  S.resident_ms.assign(gpus, 0); S.e2e_ms.assign(gpus, 0); S.bad_digests.assign(gpus, 0); S.job_checksum.assign(gpus, 0); S.violations.assign(gpus, 0);
  pthread_barrier_init(&S.bar, nullptr, gpus);
  if (gpus > 1) {
    S.comms.resize(gpus);
    std::vector<int> devs(gpus);
    for (int g = 0; g < gpus; g++) devs[g] = g;
    NK(ncclCommInitAll(S.comms.data(), gpus, devs.data()));
  }
  std::vector<std::thread> th;
  for (int g = 0; g < gpus; g++) th.emplace_back(worker, &S, g);
  for (auto& t : th) t.join();
  for (auto c : S.comms) ncclCommDestroy(c);

  const float res_ms = *std::max_element(S.resident_ms.begin(), S.resident_ms.end()) / steps;
  const float e2e_ms = *std::max_element(S.e2e_ms.begin(), S.e2e_ms.end()) / steps;
  const double blocks = (double)gpus * S.per_gpu * S.layout.n_blocks;
  int bad = 0;
  for (int b : S.bad_digests) bad += b;
  uint64_t viol = 0;
  for (uint64_t v : S.violations) viol += v;
  bool same_ck = true;
  for (int g = 1; g < gpus; g++) same_ck = same_ck && S.job_checksum[g] == S.job_checksum[0];
  printf("{\"runner\": \"native C++ (no Python, no PyTorch)\", \"metric\": \"SHA-256 blocks/sec witness-gen (bit-exact cells)\", \"workload\": \"%s\", "
         "\"n_gpus\": %d, \"instances_per_gpu\": %llu, \"blocks_per_instance\": %u, \"steps\": %d, \"value\": %.1f, \"unit\": \"blocks/s\", \"ms_per_step\": %.4f, "
         "\"e2e\": {\"value\": %.1f, \"ms_per_step\": %.4f}, \"cells_per_s\": %.4g, \"digest_mismatches\": %d, \"constraint_violations_all_instances\": %llu, \"gathered_checksum\": %llu, "
         "\"all_ranks_hold_the_same_gather\": %s}\n",
         w->name, gpus, (unsigned long long)S.per_gpu, S.layout.n_blocks, steps, blocks / (res_ms * 1e-3), res_ms, blocks / (e2e_ms * 1e-3), e2e_ms,
         blocks / (res_ms * 1e-3) * (double)S.layout.cells_per_instance / S.layout.n_blocks, bad, (unsigned long long)viol, (unsigned long long)S.job_checksum[0],
         same_ck ? "true" : "false");
  return (bad == 0 && same_ck && viol == 0) ? 0 : 1;
}
