"""Compares a dump of the REAL crate's advice columns (rust/tools/dump_witness) with the oracle, cell by cell.

    python tools/compare_rust_dump.py out.bin <hex msg 0> <pre 0> <hex msg 1> <pre 1>         # compare a dump
    python tools/compare_rust_dump.py --make-synthetic out.bin <hex msg 0> <pre 0> <hex msg 1> <pre 1>   # write the dump the oracle expects

This is the tool that turns "placement parity unpinned" (oracle header) into "pinned" on a machine that has
cargo + network; it needs no GPU.  Dump format (rust/tools/dump_witness/src/main.rs):

    magic "H2SHADMP" | u32 n_columns | u32 n_rows | n_columns x n_rows x 32-byte canonical little-endian field elements

Column order: gate advice [0,3), lookup advice, dense_0, dense_1, spread_0, spread_1 (allocation order of the reference's
configure, lib.rs:409-428 and spread.rs:39-52); unassigned cells are zero.  `--make-synthetic` writes that file from the
oracle (tests/test_compare_rust_dump.py uses it to keep this comparer known-good until the day someone runs the real one).
On a mismatch the report names the first differing cell of every column as (column, row) and, for gate columns, the
gate-stream index -- DESIGN.md §1b lists which halo2-base op pattern to suspect from there.
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

MAGIC = b"H2SHADMP"
SIZES = (128, 128)      # the reference's TestCircuit (lib.rs:487-494)
K_ROWS = 1 << 17


def _canon(a: np.ndarray) -> np.ndarray:
    """[.., 4] u64 Montgomery limbs -> canonical limbs (zeros stay zeros)."""
    flat = a.reshape(-1, 4)
    res = np.zeros_like(flat)
    for i in np.nonzero(flat.any(axis=1))[0]:
        v = O.mont_to_int(flat[i])
        res[i] = [(v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)]
    return res.reshape(a.shape)


def oracle_columns(msgs, pre, n_rows=K_ROWS):
    """[n_columns, n_rows, 4] u64 canonical: what the dump of the real crate should hold if the oracle's placement is right."""
    cfg = O.OracleConfig(max_variable_byte_sizes=SIZES)
    reg = O.synthesize(cfg, msgs, pre, record_shape=False)
    lay = O.Layout(len(reg.breaks), n_rows, 1, n_rows, n_rows)
    out = O.batch(cfg, lay, [msgs], [pre], want_cells=True)
    cols = np.concatenate([_canon(out["gate"][0]), _canon(out["lookup"][0]), _canon(out["spread"][0])], axis=0)
    return cols, reg


def write_dump(path, cols):
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", cols.shape[0], cols.shape[1]))
        f.write(np.ascontiguousarray(cols, dtype="<u8").tobytes())


def load_dump(path):
    raw = open(path, "rb").read()
    if raw[:8] != MAGIC:
        raise ValueError("bad magic: not a dump_witness file")
    n_cols, n_rows = struct.unpack("<II", raw[8:16])
    if len(raw) != 16 + n_cols * n_rows * 32:
        raise ValueError(f"truncated dump: {len(raw)} bytes for {n_cols} columns x {n_rows} rows")
    return np.frombuffer(raw, dtype="<u8", offset=16).reshape(n_cols, n_rows, 4)


def compare(path, msgs, pre):
    """-> dict(n_cols, n_rows, differing, per_column={name: (count, first_row)}, first_gate_stream_index)."""
    cells = load_dump(path)
    n_cols, n_rows = cells.shape[:2]
    want, reg = oracle_columns(msgs, pre, n_rows)
    if n_cols != want.shape[0]:
        raise ValueError(f"expected {want.shape[0]} advice columns (3 gate + 1 lookup + 2 dense + 2 spread), the dump has {n_cols}")
    ng = len(reg.breaks)
    names = [f"gate_{c}" for c in range(ng)] + ["lookup"] + ["dense_0", "dense_1", "spread_0", "spread_1"]
    per, total, first_stream = {}, 0, None
    for c, name in enumerate(names):
        bad = np.nonzero((cells[c] != want[c]).any(axis=-1))[0]
        if bad.size:
            per[name] = (int(bad.size), int(bad[0]))
            total += int(bad.size)
            if c < ng and first_stream is None:
                first_stream = int(reg.breaks[c]) + int(bad[0])
    return dict(n_cols=n_cols, n_rows=n_rows, differing=total, per_column=per, first_gate_stream_index=first_stream)


def main(argv):
    synth = argv and argv[0] == "--make-synthetic"
    if synth:
        argv = argv[1:]
    if len(argv) != 5:
        print(__doc__)
        return 2
    path, hex0, pre0, hex1, pre1 = argv
    msgs, pre = [bytes.fromhex(hex0), bytes.fromhex(hex1)], [int(pre0), int(pre1)]
    if synth:
        cols, _ = oracle_columns(msgs, pre)
        write_dump(path, cols)
        print(f"wrote {path}: {cols.shape[0]} columns x {cols.shape[1]} rows (from the oracle, not from the crate)")
        return 0
    r = compare(path, msgs, pre)
    print(f"{r['n_cols']} columns x {r['n_rows']} rows compared; {r['differing']} cells differ")
    for name, (cnt, row) in r["per_column"].items():
        print(f"  {name}: {cnt} cells differ, first at row {row}")
    if r["first_gate_stream_index"] is not None:
        print(f"  first differing gate-stream index: {r['first_gate_stream_index']} (see DESIGN.md §1b for the op patterns to suspect)")
    return 1 if r["differing"] else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
