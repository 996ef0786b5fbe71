// chip_api.hpp -- the reference's chip API (halo2-dynamic-sha256 src/lib.rs:38-369), in C++, over the C-ABI.
//
// The reference is used like this inside `Circuit::synthesize` (src/lib.rs:435-484, benches/digest.rs:73-99):
//
//     sha256.range().load_lookup_table(&mut layouter)?;          // lib.rs:442
//     sha256.load(&mut layouter)?;                               // lib.rs:443
//     layouter.assign_region(|| "...", |region| {
//         let ctx = &mut sha256.new_context(region);             // lib.rs:351-360
//         let r0 = sha256.digest(ctx, &input0, Some(pre0))?;     // lib.rs:71-76
//         let r1 = sha256.digest(ctx, &input1, Some(pre1))?;
//         range.finalize(ctx);                                   // lib.rs:469
//     })?;
//     layouter.constrain_instance(r0.output_bytes[i].cell(), hash_column, i)?;
//
// This header offers the same calls with the same names, argument meaning, state changes (`cur_hash_idx`, lib.rs:347) and
// panics (-> h2sha::ReferencePanic).  What differs is WHO computes the cells: `digest` has the engine generate the region on
// the GPU (h2sha_digest_batch with only_digest = this call: exactly the cells this `digest` call appends in the reference), or
// takes it from a batch generated earlier (`attach`), copies the columns out (h2sha_export_instance) and replays the call into
// the halo2 `Region`: every advice cell it owns, at the reference's (column,row), its selectors, the fixed cells it uses first,
// its copy constraints and its spread-table rows, in the order h2sha_get_shape reports them; the cells it looks up are pushed
// to `ctx.cells_to_lookup`, and `range.finalize(ctx)` copies them into the lookup advice column as halo2-base does.
// Keygen uses the same replay without values (`Context::shape_only`, works with a plan-only engine on a machine without a GPU).
//
// halo2_proofs itself is not available to a C++ build, so `Region` / `Layouter` are the two interfaces a binding
// implements (the Rust facade in rust/src/lib.rs implements them with halo2_proofs::circuit::{Region, Layouter});
// tests/cpp/test_chip_api.cc implements them with a recorder and hands the result to the MockProver-style checker.
#pragma once
#include <cuda_runtime.h>

#include <array>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "host_api.hpp"

namespace h2sha {
namespace chip {

using Fr = std::array<uint64_t, 4>;   // 4 x u64 little-endian limbs; advice cells in Montgomery form (halo2curves bn256::Fr layout)

// Columns in the allocation order of the reference's configure (lib.rs:409-428, spread.rs:39-52).
enum class ColumnKind : uint32_t { GateAdvice = 0, LookupAdvice = 1, SpreadDense = 2, SpreadSpread = 3, Fixed = 4, GateSelector = 5 };
struct Column {
  ColumnKind kind;
  uint32_t index;
};
struct Cell {   // halo2_proofs::circuit::Cell, region-relative
  Column column;
  uint32_t row;
};

// halo2_proofs::circuit::Region, as far as the chip uses it (spread.rs:203-227 and halo2-base's Context)
struct Region {
  virtual ~Region() {}
  virtual void assign_advice(Column column, uint32_t row, const Fr* value /* null: Value::unknown() (keygen) */) = 0;
  virtual void assign_fixed(Column column, uint32_t row, const Fr& canonical_value) = 0;
  virtual void enable_selector(Column selector, uint32_t row) = 0;
  virtual void constrain_equal(Cell a, Cell b) = 0;
};
// halo2_proofs::circuit::Layouter::assign_table, as far as `load` / `load_lookup_table` use it
struct Layouter {
  virtual ~Layouter() {}
  // one call per table: `columns[c][row]` canonical values
  virtual void assign_table(const std::string& name, const std::vector<std::vector<uint64_t>>& columns) = 0;
};

// halo2-base AssignedValue: the cell and its value
struct AssignedValue {
  Cell cell_;
  uint32_t stream_idx;                  // index in the FlexGate vertical stream of the Context
  std::shared_ptr<std::vector<Fr>> values;   // the Context's gate stream (shared; filled digest() call by digest() call)
  Cell cell() const { return cell_; }
  const Fr& value() const {
    if (!values || stream_idx >= values->size()) throw EngineError(H2SHA_EINVAL, "value() of a cell of a shape-only (keygen) context: Value::unknown()");
    return (*values)[stream_idx];
  }
};
struct AssignedHashResult {   // lib.rs:31-36
  AssignedValue input_len;
  std::vector<AssignedValue> input_bytes, output_bytes;
};

class Sha256DynamicConfig;

// halo2-base Context (lib.rs:351-360), as far as the chip and its users touch it
class Context {
 public:
  size_t total_advice = 0, total_fixed = 0;            // the numbers the reference prints under `display` (lib.rs:470-476)
  std::vector<uint32_t> cells_to_lookup;               // gate-stream indices, push order
  std::pair<uint32_t, uint32_t> advice_alloc{0, 0};    // (column, row) cursor of the FlexGate vertical stream
  // keygen: replay the shape only, every advice cell Value::unknown()
  void shape_only() { no_values = true; }

 private:
  friend class Sha256DynamicConfig;
  friend class RangeConfig;
  Region* region = nullptr;
  std::shared_ptr<std::vector<Fr>> values = std::make_shared<std::vector<Fr>>();   // the gate stream assigned so far
  std::vector<uint8_t> fixed_assigned;
  bool no_values = false, finalized = false;
};

// halo2-base RangeConfig (lib.rs:409-418), as far as the chip's users touch it
class RangeConfig {
 public:
  uint32_t lookup_bits = 16, k = 17, minimum_rows = 9, num_fixed = 1;
  struct Gate { uint32_t max_rows; } gate{(1u << 17) - 9};   // range.gate.max_rows (lib.rs:355)
  // RangeConfig::configure(meta, Vertical, &[num_advice], &[num_lookup_advice], num_fixed, lookup_bits, 0, k): the advice column
  // counts are outputs of the engine's plan here (layout().n_gate_cols / n_lookup_cols)
  static RangeConfig configure(uint32_t num_lookup_advice, uint32_t num_fixed, uint32_t lookup_bits, uint32_t k, uint32_t minimum_rows = 9) {
    RangeConfig r;
    r.num_lookup_advice = num_lookup_advice ? num_lookup_advice : 1;
    r.lookup_bits = lookup_bits; r.k = k; r.minimum_rows = minimum_rows; r.num_fixed = num_fixed ? num_fixed : 1;
    r.gate.max_rows = (1u << k) - minimum_rows;
    return r;
  }
  // lib.rs:442: the 2^lookup_bits-row range table
  void load_lookup_table(Layouter& layouter) const {
    std::vector<std::vector<uint64_t>> cols(1);
    cols[0].resize(1u << lookup_bits);
    for (uint32_t i = 0; i < (1u << lookup_bits); i++) cols[0][i] = i;
    layouter.assign_table("range lookup table", cols);
  }
  // lib.rs:469: copies cells_to_lookup into the lookup advice column(s) in push order, wrapping at max_rows, each copy-constrained
  // to the cell it copies (halo2-base RangeConfig::finalize; Table B of SURVEY.md 8a)
  void finalize(Context& ctx) const {
    for (size_t k = 0; k < ctx.cells_to_lookup.size(); k++) {
      const uint32_t col = (uint32_t)(k / gate.max_rows), row = (uint32_t)(k % gate.max_rows);
      if (col >= num_lookup_advice) throw ReferencePanic(H2SHA_EPANIC, "NOT ENOUGH LOOKUP ADVICE COLUMNS (halo2-base)");
      const Column lc{ColumnKind::LookupAdvice, col};
      const uint32_t src = ctx.cells_to_lookup[k];
      ctx.region->assign_advice(lc, row, ctx.no_values ? nullptr : &(*ctx.values)[src]);
      ctx.region->constrain_equal(Cell{lc, row}, gate_cell_of(src));
    }
    ctx.finalized = true;
  }
  uint32_t num_lookup_advice = 1;

 private:
  friend class Sha256DynamicConfig;
  std::vector<uint32_t> breaks_;   // filled by Sha256DynamicConfig::configure (gate-stream index -> column)
  Cell gate_cell_of(uint32_t stream_idx) const {
    size_t c = 0;
    while (c + 1 < breaks_.size() && breaks_[c + 1] <= stream_idx) c++;
    return Cell{Column{ColumnKind::GateAdvice, (uint32_t)c}, stream_idx - breaks_[c]};
  }
};

class Sha256DynamicConfig {
 public:
  std::vector<size_t> max_variable_byte_sizes;   // lib.rs:40
  size_t cur_hash_idx = 0;                       // lib.rs:43

  // lib.rs:49-69.  `meta` has no counterpart: the columns are implied by the plan (layout()).
  static std::unique_ptr<Sha256DynamicConfig> configure(std::vector<size_t> max_variable_byte_sizes, RangeConfig range, size_t num_bits_lookup,
                                                        size_t num_advice_columns, bool is_input_range_check, int device = 0) {
    std::unique_ptr<Sha256DynamicConfig> c(new Sha256DynamicConfig());
    for (size_t b : max_variable_byte_sizes)
      if (b % 64 != 0) throw ReferencePanic(H2SHA_EPANIC, "max_variable_byte_size must be a multiple of 64 (lib.rs:57-59)");
    c->max_variable_byte_sizes = max_variable_byte_sizes;
    c->range_ = range;
    c->spread_cols_ = (uint32_t)num_advice_columns;
    c->limb_bits_ = (uint32_t)num_bits_lookup;
    std::vector<uint32_t> sizes(max_variable_byte_sizes.begin(), max_variable_byte_sizes.end());
    h2sha_config_t cfg{};
    cfg.n_digests = (uint32_t)sizes.size(); cfg.max_variable_byte_sizes = sizes.data();
    cfg.max_rows = range.gate.max_rows; cfg.lookup_bits = range.lookup_bits; cfg.num_bits_lookup = (uint32_t)num_bits_lookup;
    cfg.num_advice_columns = (uint32_t)num_advice_columns; cfg.is_input_range_check = is_input_range_check;
    cfg.device = device; cfg.build_shape = 1; cfg.num_lookup_advice = range.num_lookup_advice;
    check(h2sha_create(&cfg, &c->engine_));
    check(h2sha_get_layout(c->engine_, &c->layout_));
    c->device_ = device;
    // the static shape, once
    const h2sha_layout_t& L = c->layout_;
    c->selectors_.resize(L.n_gate_cells); c->copies_.resize((size_t)L.n_copies * 4); c->fixed_.resize((size_t)L.n_fixed * 4);
    c->lookup_src_.resize(L.n_lookup_cells); c->limb_dense_src_.resize(L.n_spread_limbs); c->limb_spread_src_.resize(L.n_spread_limbs);
    check(h2sha_get_shape(c->engine_, c->selectors_.data(), c->copies_.data(), c->fixed_.data(), c->lookup_src_.data(), c->limb_dense_src_.data(),
                          c->limb_spread_src_.data()));
    c->breaks_.resize(L.n_gate_cols);
    check(h2sha_get_breaks(c->engine_, c->breaks_.data()));
    c->range_.breaks_ = c->breaks_;
    c->ranges_.resize(6 * sizes.size());
    check(h2sha_get_digest_ranges(c->engine_, c->ranges_.data()));
    return c;
  }
  ~Sha256DynamicConfig() { if (engine_) h2sha_destroy(engine_); }
  Sha256DynamicConfig(const Sha256DynamicConfig&) = delete;
  Sha256DynamicConfig& operator=(const Sha256DynamicConfig&) = delete;

  // lib.rs:351-360
  Context new_context(Region& region) const {
    Context ctx;
    ctx.region = &region;
    ctx.fixed_assigned.assign(layout_.n_fixed, 0);
    return ctx;
  }
  // Prover path: take the witness of the regions from a batch generated earlier with h2sha_digest_batch (device buffers) instead
  // of generating each digest() call's cells on demand; `instance` names the region inside the batch.
  void attach(const void* gate_dev, const void* lookup_dev, const void* spread_dev, uint64_t instance) {
    att_gate_ = gate_dev; att_lookup_ = lookup_dev; att_spread_ = spread_dev; att_instance_ = instance; attached_ = true;
  }
  // lib.rs:362-364
  const RangeConfig& range() const { return range_; }
  // lib.rs:366-368 -> SpreadConfig::load (spread.rs:165-194): the (dense, spread) table
  void load(Layouter& layouter) const {
    uint32_t n_rows = 0, n_range = 0;
    check(h2sha_get_lookup_tables(engine_, nullptr, nullptr, &n_rows, &n_range));
    std::vector<std::vector<uint64_t>> cols(2, std::vector<uint64_t>(n_rows));
    check(h2sha_get_lookup_tables(engine_, cols[0].data(), cols[1].data(), &n_rows, &n_range));
    layouter.assign_table("spread table", cols);
  }

  // lib.rs:71-349.  Same preconditions, same panics (lib.rs:86-90), same state change (lib.rs:347).
  AssignedHashResult digest(Context& ctx, const std::vector<uint8_t>& input, const size_t* precomputed_input_len /* Option<usize> */) {
    if (cur_hash_idx >= max_variable_byte_sizes.size())
      throw ReferencePanic(H2SHA_EPANIC, "digest() called more often than max_variable_byte_sizes has entries (index out of bounds, lib.rs:86)");
    if (ctx.finalized) throw EngineError(H2SHA_EINVAL, "digest() after range.finalize(ctx)");
    const size_t max_bytes = max_variable_byte_sizes[cur_hash_idx];
    const size_t pre = precomputed_input_len ? *precomputed_input_len : 0;
    if (pre % 64 != 0) throw ReferencePanic(H2SHA_EPANIC, "precomputed_input_len is not a multiple of 64 (lib.rs:89)");
    const size_t padded = (input.size() + 9 + 63) / 64 * 64;   // lib.rs:80-85
    if (padded < pre || padded - pre > max_bytes) throw ReferencePanic(H2SHA_EPANIC, "padded input does not fit max_variable_byte_size (lib.rs:90)");
    const uint32_t d = (uint32_t)cur_hash_idx;
    const uint32_t pre32 = (uint32_t)pre;
    assign_digest(ctx, d, input, precomputed_input_len ? &pre32 : nullptr);
    // AssignedHashResult (lib.rs:342-346)
    std::vector<uint32_t> in_idx(max_bytes), out_idx(32);
    uint32_t len_idx = 0;
    check(h2sha_get_handles(engine_, d, &len_idx, in_idx.data(), out_idx.data()));
    AssignedHashResult r;
    r.input_len = assigned(ctx, len_idx);
    for (uint32_t i : in_idx) r.input_bytes.push_back(assigned(ctx, i));
    for (uint32_t i : out_idx) r.output_bytes.push_back(assigned(ctx, i));
    cur_hash_idx += 1;   // lib.rs:347
    return r;
  }

  const h2sha_layout_t& layout() const { return layout_; }
  h2sha_engine_t* raw() const { return engine_; }
  Cell gate_cell(uint32_t stream_idx) const {
    size_t c = 0;
    while (c + 1 < breaks_.size() && breaks_[c + 1] <= stream_idx) c++;
    return Cell{Column{ColumnKind::GateAdvice, (uint32_t)c}, stream_idx - breaks_[c]};
  }
  Cell fixed_cell(uint32_t k) const { return Cell{Column{ColumnKind::Fixed, k % range_.num_fixed}, k / range_.num_fixed}; }   // Context::assign_fixed: round-robin over the fixed columns

 private:
  Sha256DynamicConfig() = default;
  static void check(int rc) {
    if (rc == H2SHA_OK) return;
    if (rc == H2SHA_EPANIC) throw ReferencePanic(rc, h2sha_last_error());
    throw EngineError(rc, h2sha_last_error());
  }
  static void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw EngineError(H2SHA_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
  }
  AssignedValue assigned(const Context& ctx, uint32_t idx) const { return AssignedValue{gate_cell(idx), idx, ctx.values}; }

  // what digest() call d appends to the region: its cells from the GPU, its share of the static shape from the plan
  void assign_digest(Context& ctx, uint32_t d, const std::vector<uint8_t>& input, const uint32_t* pre) {
    const h2sha_layout_t& L = layout_;
    const size_t D = max_variable_byte_sizes.size();
    const uint32_t* rg = &ranges_[6 * d];
    const uint32_t rows = 1u << range_.k;
    const uint32_t n_cols = L.n_gate_cols + L.n_lookup_cols + L.n_spread_cols;
    std::vector<std::vector<uint64_t>> cols;
    if (!ctx.no_values) {
      cuda_check(cudaSetDevice(device_), "cudaSetDevice");
      void *gate = nullptr, *spread = nullptr;
      const void *g = att_gate_, *s = att_spread_;
      uint64_t inst = att_instance_;
      if (!attached_) {
        cuda_check(cudaMalloc(&gate, L.gate_bytes), "cudaMalloc");
        if (cudaMalloc(&spread, L.spread_bytes) != cudaSuccess) { cudaFree(gate); throw EngineError(H2SHA_ECUDA, "cudaMalloc"); }
        cudaMemset(gate, 0, L.gate_bytes); cudaMemset(spread, 0, L.spread_bytes);
        // message d is this call's input; the other digest() calls of the region get the empty message (their cells are not generated)
        std::vector<uint64_t> offs(D, 0);
        std::vector<uint32_t> lens(D, 0), pl(D, 0);
        for (size_t k = d + 1; k < D; k++) offs[k] = input.size();
        lens[d] = (uint32_t)input.size();
        if (pre) pl[d] = *pre;
        h2sha_batch_t b{};
        b.n_instances = 1; b.msgs = input.empty() ? nullptr : input.data(); b.msgs_bytes = input.size(); b.offsets = offs.data(); b.lens = lens.data();
        b.precomputed_lens = pre ? pl.data() : nullptr;
        b.gate = gate; b.spread = spread; b.only_digest = d + 1;
        const int rc = h2sha_digest_batch(engine_, &b);
        if (rc) { cudaFree(gate); cudaFree(spread); check(rc); }
        g = gate; s = spread; inst = 0;
      }
      cols.assign(n_cols, std::vector<uint64_t>());
      std::vector<uint64_t*> ptrs(n_cols, nullptr);
      for (uint32_t c = 0; c < n_cols; c++)
        if (c < L.n_gate_cols || c >= L.n_gate_cols + L.n_lookup_cols) { cols[c].resize((size_t)rows * 4); ptrs[c] = cols[c].data(); }
      int rc = h2sha_export_instance(engine_, inst, g, nullptr, s, ptrs.data(), rows, nullptr);   // the lookup column is range.finalize's
      if (!rc && cudaStreamSynchronize(nullptr) != cudaSuccess) rc = H2SHA_ECUDA;
      cudaFree(gate); cudaFree(spread);
      check(rc);
    }
    Region& R = *ctx.region;
    auto fr_at = [&](uint32_t col, uint32_t row) {
      Fr v;
      memcpy(v.data(), &cols[col][(size_t)row * 4], 32);
      return v;
    };
    // (1) the gate cells this call owns, stream order (halo2-base Context::assign_region), + their gate selectors
    if (!ctx.no_values && ctx.values->size() < rg[1]) ctx.values->resize(rg[1]);
    for (uint32_t i = rg[0]; i < rg[1]; i++) {
      const Cell c = gate_cell(i);
      if (ctx.no_values) { R.assign_advice(c.column, c.row, nullptr); }
      else { const Fr v = fr_at(c.column.index, c.row); (*ctx.values)[i] = v; R.assign_advice(c.column, c.row, &v); }
      if (selectors_[i]) R.enable_selector(Column{ColumnKind::GateSelector, c.column.index}, c.row);
    }
    ctx.total_advice += rg[1] - rg[0];
    const Cell end = rg[1] < L.n_gate_cells ? gate_cell(rg[1]) : gate_cell(rg[1] - 1);
    ctx.advice_alloc = {end.column.index, rg[1] < L.n_gate_cells ? end.row : end.row + 1};
    // (2) its copy constraints (cell <-> earlier cell, cell <-> fixed); a fixed cell is assigned when it is first used (Context::assign_fixed)
    auto fixed_use = [&](uint32_t k) {
      if (!ctx.fixed_assigned[k]) {
        Fr v; memcpy(v.data(), &fixed_[(size_t)k * 4], 32);
        const Cell c = fixed_cell(k);
        R.assign_fixed(c.column, c.row, v);
        ctx.fixed_assigned[k] = 1; ctx.total_fixed++;
      }
      return fixed_cell(k);
    };
    for (uint32_t i = 0; i < L.n_copies; i++) {
      const uint32_t* p = &copies_[(size_t)i * 4];   // (a_kind, a_idx, b_kind, b_idx); a is always the newer gate cell
      if (p[0] != 0 || p[1] < rg[0] || p[1] >= rg[1]) continue;
      R.constrain_equal(gate_cell(p[1]), p[2] == 0 ? gate_cell(p[3]) : fixed_use(p[3]));
    }
    // (3) its spread-table rows: limb n -> column pair n % cols, row n / cols, copy-constrained to its gate cells (spread.rs:202-231)
    const uint32_t nc = spread_cols_;
    for (uint32_t n = rg[4]; n < rg[5]; n++) {
      const uint32_t col = n % nc, row = n / nc;
      const Column dc{ColumnKind::SpreadDense, col}, sc{ColumnKind::SpreadSpread, col};
      if (ctx.no_values) { R.assign_advice(dc, row, nullptr); R.assign_advice(sc, row, nullptr); }
      else {
        const Fr dv = fr_at(L.n_gate_cols + L.n_lookup_cols + col, row), sv = fr_at(L.n_gate_cols + L.n_lookup_cols + nc + col, row);
        R.assign_advice(dc, row, &dv); R.assign_advice(sc, row, &sv);
      }
      R.constrain_equal(Cell{dc, row}, gate_cell(limb_dense_src_[n]));
      R.constrain_equal(Cell{sc, row}, gate_cell(limb_spread_src_[n]));
    }
    // (4) the cells it looks up, push order: range.finalize(ctx) copies them into the lookup advice column (lib.rs:469)
    for (uint32_t k = rg[2]; k < rg[3]; k++) ctx.cells_to_lookup.push_back(lookup_src_[k]);
  }

  h2sha_engine_t* engine_ = nullptr;
  h2sha_layout_t layout_{};
  RangeConfig range_;
  int device_ = 0;
  uint32_t spread_cols_ = 2, limb_bits_ = 8;
  std::vector<uint8_t> selectors_;
  std::vector<uint32_t> copies_, lookup_src_, limb_dense_src_, limb_spread_src_, breaks_, ranges_;
  const void *att_gate_ = nullptr, *att_lookup_ = nullptr, *att_spread_ = nullptr;
  uint64_t att_instance_ = 0;
  bool attached_ = false;
  std::vector<uint64_t> fixed_;
};

}  // namespace chip
}  // namespace h2sha
