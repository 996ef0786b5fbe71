"""Compares a dump of the REAL crate's advice columns (rust/tools/dump_witness) with the oracle, cell by cell.

    python tools/compare_rust_dump.py out.bin <hex msg 0> <pre 0> <hex msg 1> <pre 1>

This is the tool that turns "placement parity unpinned" (oracle header) into "pinned" on a machine that has
cargo + network; it needs no GPU.  Column order of the dump: gate advice [0,3), lookup advice, dense_0, dense_1,
spread_0, spread_1 (allocation order of the reference's configure, lib.rs:409-428 and spread.rs:39-52).
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def main():
    path, hex0, pre0, hex1, pre1 = sys.argv[1:6]
    raw = open(path, "rb").read()
    assert raw[:8] == b"H2SHADMP", "bad magic"
    n_cols, n_rows = struct.unpack("<II", raw[8:16])
    cells = np.frombuffer(raw, dtype="<u8", offset=16).reshape(n_cols, n_rows, 4)
    cfg = O.OracleConfig(max_variable_byte_sizes=(128, 128))
    msgs = [bytes.fromhex(hex0), bytes.fromhex(hex1)]
    reg = O.synthesize(cfg, msgs, [int(pre0), int(pre1)], record_shape=False)
    lay = O.Layout(len(reg.breaks), n_rows, 1, n_rows, n_rows)
    out = O.batch(cfg, lay, [msgs], [[int(pre0), int(pre1)]], want_cells=True)
    # oracle buffers are Montgomery form; the dump is canonical: convert the oracle side
    def canon(a):
        flat = a.reshape(-1, 4)
        res = np.zeros_like(flat)
        nz = np.nonzero(flat.any(axis=1))[0]
        for i in nz:
            v = O.mont_to_int(flat[i])
            res[i] = [(v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)]
        return res.reshape(a.shape)
    gate, lookup, spread = canon(out["gate"][0]), canon(out["lookup"][0]), canon(out["spread"][0])
    ng = gate.shape[0]
    assert n_cols == ng + 1 + 4, f"expected {ng + 5} advice columns, dump has {n_cols}"
    bad = 0
    for c in range(ng):
        bad += int((cells[c] != gate[c]).any(axis=-1).sum())
    bad += int((cells[ng] != lookup[0]).any(axis=-1).sum())
    for c in range(4):
        bad += int((cells[ng + 1 + c] != spread[c]).any(axis=-1).sum())
    print(f"{n_cols} columns x {n_rows} rows compared; {bad} cells differ")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
