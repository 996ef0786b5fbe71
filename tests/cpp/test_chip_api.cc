// The reference's TestCircuit (halo2-dynamic-sha256 src/lib.rs:400-494) written against csrc/chip_api.hpp: the same
// call sequence as `synthesize` (lib.rs:435-484) -- load_lookup_table, load, new_context, digest, digest, range.finalize --
// with a recording Region / Layouter standing in for halo2_proofs.  Everything the recorder sees (advice / fixed
// assignments, selectors, copy constraints, tables) is written to a file; tests/test_gpu_chip_api.py rebuilds the region from
// that file alone and hands it to the MockProver-style checker and to a cell-by-cell comparison with the oracle.
//
//   test_chip_api <out.bin> <case>      case: 1 = test_sha256_correct1, 3 = test_sha256_correct3, 4 = test_sha256_correct4 (precomputed 128),
//                                             14 = case 4 taken from a 3-instance batch generated earlier (Context::attach),
//                                             0 = keygen (shape only, plan-only engine: runs without a GPU)
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../halo2-dynamic-sha256_b200/csrc/chip_api.hpp"

using namespace h2sha::chip;

struct Recorder : Region, Layouter {
  std::vector<uint32_t> w;   // event stream, u32 words
  void fr(const Fr& v) { for (int i = 0; i < 4; i++) { w.push_back((uint32_t)v[i]); w.push_back((uint32_t)(v[i] >> 32)); } }
  void assign_advice(Column c, uint32_t row, const Fr* v) override {
    w.push_back(1); w.push_back((uint32_t)c.kind); w.push_back(c.index); w.push_back(row); w.push_back(v ? 1 : 0);
    if (v) fr(*v);
  }
  void assign_fixed(Column c, uint32_t row, const Fr& v) override { w.push_back(2); w.push_back(c.index); w.push_back(row); fr(v); }
  void enable_selector(Column c, uint32_t row) override { w.push_back(3); w.push_back(c.index); w.push_back(row); }
  void constrain_equal(Cell a, Cell b) override {
    w.push_back(4);
    w.push_back((uint32_t)a.column.kind); w.push_back(a.column.index); w.push_back(a.row);
    w.push_back((uint32_t)b.column.kind); w.push_back(b.column.index); w.push_back(b.row);
  }
  void assign_table(const std::string& name, const std::vector<std::vector<uint64_t>>& cols) override {
    w.push_back(5); w.push_back((uint32_t)name.size()); w.push_back((uint32_t)cols.size()); w.push_back((uint32_t)cols[0].size());
    for (char ch : name) w.push_back((uint32_t)(unsigned char)ch);
    for (auto& c : cols) for (uint64_t v : c) { w.push_back((uint32_t)v); w.push_back((uint32_t)(v >> 32)); }
  }
  void instance(const Cell& c, uint32_t idx) { w.push_back(6); w.push_back((uint32_t)c.column.kind); w.push_back(c.column.index); w.push_back(c.row); w.push_back(idx); }
};

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s <out.bin> <case>\n", argv[0]); return 2; }
  const int tc = atoi(argv[2]);
  try {
    // TestCircuit::configure (lib.rs:408-433, 487-494)
    const uint32_t NUM_LOOKUP_ADVICE = 1, NUM_FIXED = 1, LOOKUP_BITS = 16, K = 17;
    RangeConfig range_config = RangeConfig::configure(NUM_LOOKUP_ADVICE, NUM_FIXED, LOOKUP_BITS, K);
    auto sha256 = Sha256DynamicConfig::configure({128, 128}, range_config, 8, 2, true, tc == 0 ? -1 : 0);
    if (sha256->layout().n_gate_cols != 3 || sha256->layout().n_lookup_cols != 1) { fprintf(stderr, "NUM_ADVICE = 3, NUM_LOOKUP_ADVICE = 1 expected (lib.rs:490-492)\n"); return 1; }
    std::vector<std::vector<uint8_t>> test_inputs;
    std::vector<size_t> pre = {0, 0};
    if (tc == 1 || tc == 0) test_inputs = {{'a', 'b', 'c'}, {}};
    else if (tc == 3) test_inputs = {std::vector<uint8_t>(56, 1), {0, 0, 0}};
    else {
      std::vector<uint8_t> m0(192), m1(192);
      for (int i = 0; i < 192; i++) { m0[i] = (uint8_t)i; m1[i] = (uint8_t)(i + 64); }
      test_inputs = {m0, m1};
      pre = {128, 128};
    }
    Recorder rec;
    // synthesize (lib.rs:435-484)
    RangeConfig range = sha256->range();
    sha256->range().load_lookup_table(rec);
    sha256->load(rec);
    Context ctx = sha256->new_context(rec);
    void *gate = nullptr, *lookup = nullptr, *spread = nullptr;
    if (tc == 0) ctx.shape_only();
    if (tc == 14) {
      // a prover that generated a whole batch up front: instance 1 of 3 holds this circuit's inputs
      const h2sha_layout_t& L = sha256->layout();
      cudaMalloc(&gate, 3 * L.gate_bytes); cudaMalloc(&lookup, 3 * L.lookup_bytes); cudaMalloc(&spread, 3 * L.spread_bytes);
      cudaMemset(gate, 0, 3 * L.gate_bytes); cudaMemset(lookup, 0, 3 * L.lookup_bytes); cudaMemset(spread, 0, 3 * L.spread_bytes);
      std::vector<uint8_t> blob;
      std::vector<uint64_t> offs;
      std::vector<uint32_t> lens, pl;
      auto push = [&](const std::vector<uint8_t>& m, uint32_t p) { offs.push_back(blob.size()); lens.push_back((uint32_t)m.size()); pl.push_back(p); blob.insert(blob.end(), m.begin(), m.end()); };
      push({'x'}, 0); push({}, 0); push(test_inputs[0], 128); push(test_inputs[1], 128); push(std::vector<uint8_t>(100, 7), 64); push({1, 2}, 0);
      h2sha_batch_t b{};
      b.n_instances = 3; b.msgs = blob.data(); b.msgs_bytes = blob.size(); b.offsets = offs.data(); b.lens = lens.data(); b.precomputed_lens = pl.data();
      b.gate = gate; b.lookup = lookup; b.spread = spread;
      if (h2sha_digest_batch(sha256->raw(), &b)) { fprintf(stderr, "batch: %s\n", h2sha_last_error()); return 1; }
      sha256->attach(gate, lookup, spread, 1);
    }
    std::vector<Cell> assigned_hash_cells;
    AssignedHashResult result0 = sha256->digest(ctx, test_inputs[0], &pre[0]);
    for (auto& v : result0.output_bytes) assigned_hash_cells.push_back(v.cell());
    AssignedHashResult result1 = sha256->digest(ctx, test_inputs[1], &pre[1]);
    for (auto& v : result1.output_bytes) assigned_hash_cells.push_back(v.cell());
    if (sha256->cur_hash_idx != 2) { fprintf(stderr, "cur_hash_idx must advance (lib.rs:347)\n"); return 1; }
    range.finalize(ctx);
    for (size_t i = 0; i < assigned_hash_cells.size(); i++) rec.instance(assigned_hash_cells[i], (uint32_t)i);   // lib.rs:480-482
    if (tc != 0) {
      // AssignedValue::value() is available once the region exists (what the reference's tests pin through the instance column)
      Fr acc{};
      for (auto& v : result0.output_bytes) for (int i = 0; i < 4; i++) acc[i] ^= v.value()[i];
      if ((acc[0] | acc[1] | acc[2] | acc[3]) == 0) { fprintf(stderr, "output byte values look unassigned\n"); return 1; }
    }
    // a third digest() has no max_variable_byte_size left: the reference indexes out of bounds (lib.rs:86)
    try { sha256->digest(ctx, {}, nullptr); fprintf(stderr, "third digest() must fail\n"); return 1; } catch (const h2sha::ReferencePanic&) {}
    printf("total advice cells: %zu\nmaximum rows used by a fixed column: %zu\nlookup cells used: %zu\n", ctx.total_advice, ctx.total_fixed + 1, ctx.cells_to_lookup.size());
    FILE* f = fopen(argv[1], "wb");
    if (!f) { perror("fopen"); return 1; }
    fwrite(rec.w.data(), 4, rec.w.size(), f);
    fclose(f);
    cudaFree(gate); cudaFree(lookup); cudaFree(spread);
    printf("ok %zu words\n", rec.w.size());
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "FAIL: %s\n", e.what());
    return 1;
  }
}
