"""Host logic of the product (no GPU): the planner's static plan -- stream lengths, column breaks, selectors, copy
constraints (same order), fixed column, lookup sources, spread-limb sources, AssignedHashResult handles -- must equal
what the oracle records while walking the reference's program with concrete values."""
import numpy as np
import pytest

from oracle import oracle as O

CONFIGS = [
    dict(max_variable_byte_sizes=(64,)),
    dict(max_variable_byte_sizes=(128, 128)),                 # reference test circuit (lib.rs:487-494)
    dict(max_variable_byte_sizes=(320,)),                     # BASELINE config 4
    dict(max_variable_byte_sizes=(64, 192, 128)),
    dict(max_variable_byte_sizes=(128,), max_rows=4099),      # many column wraps
    dict(max_variable_byte_sizes=(64,), lookup_bits=8),
    dict(max_variable_byte_sizes=(64,), lookup_bits=12),
    dict(max_variable_byte_sizes=(64,), limb_bits=4),
    dict(max_variable_byte_sizes=(64,), limb_bits=2, spread_cols=3),
    dict(max_variable_byte_sizes=(128,), spread_cols=1),
    dict(max_variable_byte_sizes=(128,), is_input_range_check=False),
    dict(max_variable_byte_sizes=(128,), limb_bits=16),       # one 16-bit limb per spread (legal in the reference: 16 % num_bits_lookup == 0, spread.rs:37)
    dict(max_variable_byte_sizes=(6144, 64)),      # a 96-block digest: its prologue/epilogue job is cut into several job classes
]


def _engine(pkg, kw, **extra):
    return pkg.Sha256DynamicConfig.configure(list(kw["max_variable_byte_sizes"]), max_rows=kw.get("max_rows", (1 << 17) - 9),
                                             lookup_bits=kw.get("lookup_bits", 16), num_bits_lookup=kw.get("limb_bits", 8),
                                             num_advice_columns=kw.get("spread_cols", 2),
                                             is_input_range_check=kw.get("is_input_range_check", True), device=-1, build_shape=True, **extra)


@pytest.mark.parametrize("kw", CONFIGS, ids=[str(i) for i in range(len(CONFIGS))])
def test_plan_shape_equals_oracle_shape(pkg, kw):
    cfg = _engine(pkg, kw)
    lay, sh = cfg.layout, cfg.shape()
    D = len(kw["max_variable_byte_sizes"])
    reg = O.synthesize(O.OracleConfig(**kw), [bytes([d + 1]) * (7 * d + 3) for d in range(D)])
    assert (lay.n_gate_cells, lay.n_lookup_cells, lay.n_spread_limbs) == (reg.n_gate, len(reg.lookup_idx), reg.dense.shape[0])
    assert (cfg.breaks() == reg.breaks).all()
    assert (sh.selectors == reg.selectors).all()
    assert sh.copies.shape == reg.copies.shape and (sh.copies == reg.copies).all()
    assert (sh.lookup_src == reg.lookup_idx).all()
    assert (sh.limb_dense_src == reg.limb_gate_dense).all() and (sh.limb_spread_src == reg.limb_gate_spread).all()
    fixed = [int(a) | int(b) << 64 | int(c) << 128 | int(d) << 192 for a, b, c, d in sh.fixed]
    assert fixed == [O.mont_to_int(c) for c in reg.consts]
    olay = reg.layout()
    assert (lay.n_gate_cols, lay.gate_col_rows, lay.n_lookup_cols, lay.lookup_col_rows, lay.spread_rows) == \
        (olay.n_gate_cols, olay.gate_col_rows, olay.n_lookup_cols, olay.lookup_col_rows, olay.spread_rows)
    for d in range(D):
        h = cfg.handles(d)
        assert h.input_len == reg.input_len_idx[d]
        assert (h.input_bytes == reg.input_bytes_idx[d]).all() and (h.output_bytes == reg.output_bytes_idx[d]).all()
    cfg.close()


def test_cell_accounting_matches_survey(pkg):
    cfg = _engine(pkg, dict(max_variable_byte_sizes=(1024,)))  # benches/digest.rs shape
    lay = cfg.layout
    assert (lay.n_gate_cells, lay.n_lookup_cells, lay.n_spread_limbs // 2, lay.n_gate_cols) == (1116315, 53059, 32960, 9)
    assert lay.cells_per_instance == 1301214 and lay.n_blocks == 16


@pytest.mark.parametrize("bad", [dict(max_variable_byte_sizes=(100,)), dict(max_variable_byte_sizes=()), dict(max_variable_byte_sizes=(64,), limb_bits=3),
                                 dict(max_variable_byte_sizes=(64,), limb_bits=32), dict(max_variable_byte_sizes=(64,), lookup_bits=40)])
def test_rejected_configurations(pkg, bad):
    with pytest.raises(pkg.EngineError):
        _engine(pkg, bad)


def test_plan_only_engine_never_generates(pkg):
    cfg = _engine(pkg, dict(max_variable_byte_sizes=(64,)))
    with pytest.raises(pkg.EngineError):
        cfg.digest_batch([[b"abc"]])
    offs, lens = np.zeros(1, np.uint64), np.array([3], np.uint32)
    with pytest.raises(pkg.EngineError):  # straight through the C-ABI as well
        cfg.digest_batch_raw(1, 0, False, 0, offs, lens, None)


def test_wider_column_strides(pkg):
    cfg = _engine(pkg, dict(max_variable_byte_sizes=(64,)), gate_col_rows=1 << 17, lookup_col_rows=1 << 17, spread_rows=1 << 17)
    lay = cfg.layout
    assert lay.gate_bytes == (1 << 17) * 32 and lay.spread_bytes == 4 * (1 << 17) * 32
    with pytest.raises(pkg.EngineError):
        _engine(pkg, dict(max_variable_byte_sizes=(64,)), gate_col_rows=1000)


def test_random_configurations_shape_property(pkg):
    """Property test (seeded): for random chip configurations the planner's shape equals the oracle's."""
    rng = np.random.default_rng(2024)
    for _ in range(12):
        kw = dict(max_variable_byte_sizes=tuple(int(64 * rng.integers(1, 4)) for _ in range(int(rng.integers(1, 3)))),
                  lookup_bits=int(rng.choice([8, 9, 10, 11, 12, 13, 14, 16, 17, 18, 19, 20])), limb_bits=int(rng.choice([1, 2, 4, 8])), spread_cols=int(rng.integers(1, 5)),
                  is_input_range_check=bool(rng.integers(0, 2)), max_rows=int(rng.integers(3000, 200000)))
        cfg = _engine(pkg, kw)
        sh, lay = cfg.shape(), cfg.layout
        D = len(kw["max_variable_byte_sizes"])
        msgs = [bytes(rng.integers(0, 256, int(rng.integers(0, m - 8)), dtype=np.uint8)) for m in kw["max_variable_byte_sizes"]]
        reg = O.synthesize(O.OracleConfig(**kw), msgs)
        assert (lay.n_gate_cells, lay.n_lookup_cells, lay.n_spread_limbs) == (reg.n_gate, len(reg.lookup_idx), reg.dense.shape[0]), kw
        assert (cfg.breaks() == reg.breaks).all(), kw
        assert (sh.selectors == reg.selectors).all() and (sh.copies == reg.copies).all(), kw
        assert (sh.lookup_src == reg.lookup_idx).all(), kw
        assert [int(a) | int(b) << 64 | int(c) << 128 | int(d) << 192 for a, b, c, d in sh.fixed] == [O.mont_to_int(c) for c in reg.consts], kw
        cfg.close()


@pytest.mark.parametrize("limb_bits,lookup_bits", [(8, 16), (4, 12), (2, 8)])
def test_lookup_tables_follow_spread_config_load(pkg, limb_bits, lookup_bits):
    """SpreadConfig::load (spread.rs:165-194): row i = (i, bits of i moved to even positions); the range table has 2^lookup_bits rows."""
    e = _engine(pkg, dict(max_variable_byte_sizes=(64,), limb_bits=limb_bits, lookup_bits=lookup_bits))
    dense, spread, n_range = e.lookup_tables()
    assert n_range == 1 << lookup_bits
    assert dense.tolist() == list(range(1 << limb_bits))
    for i, s in enumerate(spread.tolist()):
        bits = [(i >> b) & 1 for b in range(limb_bits)]           # fe_to_bits_le (utils.rs:6-14)
        assert s == sum(bit << (2 * b) for b, bit in enumerate(bits))
        assert s == O.spread_bits(i) if hasattr(O, "spread_bits") else True
    e.close()
