"""Pins the CPU oracle (oracle/h2sha_oracle.c) against everything the reference's own tests hold for this path
(SURVEY.md 8c): the known-answer digests of src/lib.rs:497-611, the column budgets of lib.rs:490 and
benches/digest.rs:105, constraint satisfaction as checked by MockProver (oracle/mock_prover.py), and the cell
accounting of SURVEY.md 8a.  Cell *placement* has no reference-side pin (the Rust crate cannot run here)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import mock_prover as MP
from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TEST_CFG = dict(max_variable_byte_sizes=(128, 128))  # lib.rs:487-494, 421-428 (k = 17 -> max_rows = 2^17 - 9)

# (inputs, precomputed lens, expected digests) -- src/lib.rs:497-527, 530-556, 559-584
ABC_BITS = [0b10111010, 0b01111000, 0b00010110, 0b10111111, 0b10001111, 0b00000001, 0b11001111, 0b11101010, 0b01000001, 0b01000001,
            0b01000000, 0b11011110, 0b01011101, 0b10101110, 0b00100010, 0b00100011, 0b10110000, 0b00000011, 0b01100001, 0b10100011,
            0b10010110, 0b00010111, 0b01111010, 0b10011100, 0b10110100, 0b00010000, 0b11111111, 0b01100001, 0b11110010, 0b00000000,
            0b00010101, 0b10101101]
EMPTY = "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"
REFERENCE_KATS = [
    ("correct1", [b"abc", b""], [0, 0], [bytes(ABC_BITS), bytes.fromhex(EMPTY)]),
    ("correct2", [b"\x00", b""], [0, 0], [bytes.fromhex("6e340b9cffb37a989ca544e6bb780a2c78901d3fb33738768511a30617afa01d"), bytes.fromhex(EMPTY)]),
    ("correct3", [b"\x01" * 56, b"\x00\x00\x00"], [0, 0], [bytes.fromhex("51e14a913680f24c85fe3b0e2e5b57f7202f117bb214f8ffdd4ea0f4e921fd52"),
                                                           bytes.fromhex("709e80c88487a2411e1ee4dfb9f22a861492d20c4765150c0c794abd70f8147c")]),
]


def test_field_constants_are_derived_correctly():
    L = O.lib()
    import ctypes as C
    r = (C.c_uint64 * 4)(); r2 = (C.c_uint64 * 4)(); inv = C.c_uint64()
    L.h2o_fr_consts(r, r2, C.byref(inv))
    to_int = lambda a: sum(int(a[i]) << (64 * i) for i in range(4))
    assert to_int(r) == (1 << 256) % O.P
    assert to_int(r2) == (1 << 512) % O.P
    assert (int(inv.value) * (O.P & ((1 << 64) - 1)) + 1) % (1 << 64) == 0
    for v in [0, 1, 2, 255, 0x55555555, (1 << 32) - 1, (1 << 64) - 1, 0x123456789ABCDEF0]:
        out = (C.c_uint64 * 4)()
        L.h2o_fr_from_u64(v, out)
        assert to_int(out) == v * ((1 << 256) % O.P) % O.P
        assert O.mont_to_int(out) == v


@pytest.mark.parametrize("name,msgs,pre,expected", REFERENCE_KATS)
def test_reference_known_answers_and_mock_prover(name, msgs, pre, expected):
    """test_sha256_correct1..3: digests equal the pinned vectors AND every gate / copy / lookup constraint holds."""
    reg = O.synthesize(O.OracleConfig(**TEST_CFG), msgs, pre)
    assert reg.digests == expected
    assert [hashlib.sha256(m).digest() for m in msgs] == expected
    stats = MP.verify_region(reg, expected)
    assert stats["cells"] == 279797 and stats["lookups"] == 13382
    # NUM_ADVICE = 3 (lib.rs:490): the gate stream must need exactly 3 columns at k = 17
    assert len(reg.breaks) == 3


def test_reference_correct4_precomputed_prefix():
    """test_sha256_correct4 (lib.rs:587-611): 192 random bytes, first 128 absorbed un-constrained."""
    rng = np.random.default_rng(4)
    msgs = [bytes(rng.integers(0, 256, 192, dtype=np.uint8)) for _ in range(2)]
    expected = [hashlib.sha256(m).digest() for m in msgs]
    reg = O.synthesize(O.OracleConfig(**TEST_CFG), msgs, [128, 128])
    assert reg.digests == expected
    MP.verify_region(reg, expected)


def test_bench_shape_needs_nine_columns():
    """benches/digest.rs:102-109,129: one digest of [0x01;56] with max 1024 fits NUM_ADVICE = 9 columns at k = 17 and not 8."""
    reg = O.synthesize(O.OracleConfig(max_variable_byte_sizes=(1024,)), [b"\x01" * 56], record_shape=False)
    assert reg.n_gate == 1116315 and len(reg.lookup_idx) == 53059 and reg.dense.shape[0] == 2 * 32960
    assert len(reg.breaks) == 9
    assert reg.digests[0] == bytes.fromhex("51e14a913680f24c85fe3b0e2e5b57f7202f117bb214f8ffdd4ea0f4e921fd52")


@pytest.mark.parametrize("max_bytes,gate,lookups,rows", [(64, 70155, 3379, 2060), (128, 139899, 6691, 4120), (320, 349131, 16627, 10300),
                                                          (1088, 1186059, 56371, 35020)])
def test_cell_accounting(max_bytes, gate, lookups, rows):
    """SURVEY.md 8a totals: 69 348 gate + 3 184 lookup cells + 2 060 spread rows per block, plus the digest overhead."""
    reg = O.synthesize(O.OracleConfig(max_variable_byte_sizes=(max_bytes,), max_rows=1 << 30), [b"x" * 10], record_shape=False)
    assert (reg.n_gate, len(reg.lookup_idx), reg.dense.shape[0] // 2) == (gate, lookups, rows)
    R = max_bytes // 64
    assert reg.n_gate == 1 + 69348 * R + 46 + max_bytes + 4 * max_bytes + 76 * (R + 1) + 288


@pytest.mark.parametrize("length", [0, 1, 55, 56, 63, 64, 119, 120, 183, 247])
def test_edge_lengths_dynamic(length):
    """Padding / length selection at the block boundaries (lib.rs:77-117, 294-310), max 256 bytes."""
    msg = bytes((7 * i + length) & 0xFF for i in range(length))
    reg = O.synthesize(O.OracleConfig(max_variable_byte_sizes=(256,)), [msg])
    assert reg.digests[0] == hashlib.sha256(msg).digest()
    MP.verify_region(reg, [hashlib.sha256(msg).digest()])


def test_panics_become_errors():
    cfg = O.OracleConfig(max_variable_byte_sizes=(128,))
    with pytest.raises(ValueError):  # lib.rs:89
        O.synthesize(cfg, [b"abc"], [32])
    with pytest.raises(ValueError):  # lib.rs:90: 120 + 9 > 128
        O.synthesize(cfg, [b"a" * 120])
    O.synthesize(cfg, [b"a" * 119])  # exactly fits


@pytest.mark.parametrize("kw", [dict(lookup_bits=8), dict(lookup_bits=12), dict(limb_bits=4), dict(spread_cols=1), dict(spread_cols=3),
                                dict(is_input_range_check=False), dict(max_rows=5000)])
def test_alternative_configurations_are_constraint_consistent(kw):
    cfg = O.OracleConfig(max_variable_byte_sizes=(64,), **kw)
    msg = b"The quick brown fox jumps over the lazy dog"
    reg = O.synthesize(cfg, [msg])
    MP.verify_region(reg, [hashlib.sha256(msg).digest()])


def test_golden_fixture_checksums():
    """Regression pin: the committed oracle checksums (tests/golden/make_golden.py) still reproduce."""
    with open(os.path.join(GOLDEN, "oracle_checksums.json")) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        cfg = O.OracleConfig(max_variable_byte_sizes=tuple(case["max_variable_byte_sizes"]))
        msgs = [bytes.fromhex(h) for h in case["msgs_hex"]]
        reg = O.synthesize(cfg, msgs, case["pre_lens"], record_shape=False)
        out = O.batch(cfg, reg.layout(), [msgs], [case["pre_lens"]])
        assert [d.hex() for d in reg.digests] == case["digests_hex"]
        assert [int(x) for x in out["checksums"][0]] == case["checksums"]
