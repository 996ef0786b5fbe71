"""Three restatements of cell placement -- the planner (csrc/planner.cc, through the C-ABI with a plan-only engine), the oracle
(oracle/h2sha_oracle.c) and the value-free cursor model (oracle/placement_model.py, written from SURVEY.md Table B alone) --
must agree on the gate-stream length, the column break points, the number of spread limbs and the gate-stream index of every
looked-up cell in push order, for the BASELINE shapes and for random chip configurations.  A disagreement means one of them
misreads a halo2-base op pattern; agreement does NOT pin the patterns against halo2-base itself (parity unpinned)."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import placement_model as PM

BASELINE_SHAPES = [
    dict(max_variable_byte_sizes=(128,)),                 # cfg 1
    dict(max_variable_byte_sizes=(64,)),                  # cfg 2
    dict(max_variable_byte_sizes=(1088,)),                # cfg 3
    dict(max_variable_byte_sizes=(320,)),                 # cfg 4
    dict(max_variable_byte_sizes=(2112,)),                # cfg 5
    dict(max_variable_byte_sizes=(128, 128)),             # the reference's TestCircuit (lib.rs:487-494)
    dict(max_variable_byte_sizes=(1024,)),                # the reference's bench circuit (benches/digest.rs:102-109)
    dict(max_variable_byte_sizes=(128,), limb_bits=16),   # num_bits_lookup = 16: one limb per spread (spread.rs:37 accepts it)
    dict(max_variable_byte_sizes=(128,), limb_bits=1),    # ... and 1: sixteen limbs per spread
]


def _three_ways(pkg, kw):
    sizes = tuple(kw["max_variable_byte_sizes"])
    max_rows = kw.get("max_rows", (1 << 17) - 9)
    lookup_bits, limb_bits, rc = kw.get("lookup_bits", 16), kw.get("limb_bits", 8), kw.get("is_input_range_check", True)
    spread_cols = kw.get("spread_cols", 2)
    m = PM.place(sizes, max_rows=max_rows, lookup_bits=lookup_bits, limb_bits=limb_bits, input_range_check=rc)
    eng = pkg.Sha256DynamicConfig.configure(list(sizes), max_rows=max_rows, lookup_bits=lookup_bits, num_bits_lookup=limb_bits,
                                            num_advice_columns=spread_cols, is_input_range_check=rc, device=-1, build_shape=True)
    lay, sh, brk = eng.layout, eng.shape(), eng.breaks()
    reg = O.synthesize(O.OracleConfig(max_variable_byte_sizes=sizes, max_rows=max_rows, lookup_bits=lookup_bits, limb_bits=limb_bits,
                                      spread_cols=spread_cols, is_input_range_check=rc), [b""] * len(sizes), record_shape=True)
    eng.close()
    return m, (lay, sh, brk), reg


def _assert_agree(m, planner, reg, what):
    lay, sh, brk = planner
    assert m.n == lay.n_gate_cells == reg.n_gate, f"{what}: gate-stream length model {m.n} planner {lay.n_gate_cells} oracle {reg.n_gate}"
    assert m.breaks == list(brk) == list(reg.breaks), f"{what}: column breaks differ"
    assert m.limbs == lay.n_spread_limbs == reg.dense.shape[0], f"{what}: spread limbs differ"
    assert len(m.lookups) == lay.n_lookup_cells == len(reg.lookup_idx), f"{what}: looked-up cells differ"
    ml = np.array(m.lookups, dtype=np.uint32)
    assert (ml == sh.lookup_src).all() and (ml == reg.lookup_idx).all(), f"{what}: lookup push order differs"


@pytest.mark.parametrize("kw", BASELINE_SHAPES, ids=lambda k: "x".join(str(s) for s in k["max_variable_byte_sizes"]) + (f"-limb{k['limb_bits']}" if "limb_bits" in k else ""))
def test_baseline_shapes(pkg, kw):
    m, planner, reg = _three_ways(pkg, kw)
    _assert_agree(m, planner, reg, str(kw))


def test_reference_column_budgets_from_the_model_alone():
    """3 gate columns for the TestCircuit, 9 (not 8) for the bench circuit -- from the cursor model on its own."""
    assert len(PM.place((128, 128)).breaks) == 3
    assert len(PM.place((1024,)).breaks) == 9
    assert PM.place((64,)).n == 70155 and len(PM.place((64,)).lookups) == 3379 and PM.place((64,)).limbs == 4120   # SURVEY.md 8a totals (cfg 2)


def test_fifty_random_configurations(pkg):
    rng = np.random.default_rng(20261018)
    done = 0
    while done < 50:
        nd = int(rng.integers(1, 4))
        kw = dict(max_variable_byte_sizes=tuple(int(64 * rng.integers(1, 5)) for _ in range(nd)),
                  max_rows=int(rng.choice([(1 << 17) - 9, (1 << 16) - 9, (1 << 15) - 10, 50000, 9973, 4096])),
                  lookup_bits=int(rng.choice([8, 9, 10, 12, 13, 16, 17, 20])), limb_bits=int(rng.choice([2, 4, 8, 8])),
                  spread_cols=int(rng.integers(1, 4)), is_input_range_check=bool(rng.integers(0, 2)))
        try:
            m, planner, reg = _three_ways(pkg, kw)
        except pkg.EngineError:
            continue   # a configuration the engine refuses (shared memory) -- refused, never mis-generated
        _assert_agree(m, planner, reg, str(kw))
        done += 1
