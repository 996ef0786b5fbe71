"""Mini-MockProver (TEST INFRASTRUCTURE ONLY).

Re-checks, with independent Python big-int arithmetic, the four things halo2's
`MockProver::verify()` checks for the reference's test circuit (src/lib.rs:400-485, 525-526):

  (i)   every enabled gate row:  a + b*c - d == 0 over 4 consecutive rows of one column
        (halo2-base FlexGate, Vertical),
  (ii)  every copy constraint (advice<->advice, advice<->fixed constant, spread columns<->gate cells),
  (iii) every looked-up cell is in the 2^lookup_bits range table (lib.rs:442) and every
        (dense, spread) pair is a row of the spread table (spread.rs:56-62, 165-194),
  (iv)  the 32 output-byte cells equal the expected digest (instance column, lib.rs:480-482).

It is applied both to the oracle's own output and to the CUDA engine's output + shape plan.
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Sequence

import numpy as np

P = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
R_INV = pow((1 << 256) % P, -1, P)

CP_GATE, CP_FIXED, CP_LOOKUP, CP_DENSE, CP_SPREAD = 0, 1, 2, 3, 4


def canon_all(limbs: np.ndarray) -> List[int]:
    """[n,4] u64 Montgomery -> list of canonical ints."""
    out = []
    for a, b, c, d in limbs.tolist():
        out.append((a | (b << 64) | (c << 128) | (d << 192)) * R_INV % P)
    return out


def spread_bits(x: int) -> int:
    r = 0
    i = 0
    while x:
        if x & 1:
            r |= 1 << (2 * i)
        x >>= 1
        i += 1
    return r


def verify(*, gate: np.ndarray, selectors: np.ndarray, breaks: Sequence[int], lookup_idx: np.ndarray, dense: np.ndarray, spread: np.ndarray,
           limb_gate_dense: np.ndarray, limb_gate_spread: np.ndarray, copies: np.ndarray, consts: np.ndarray, lookup_bits: int, limb_bits: int,
           max_rows: int, output_bytes_idx: Sequence[np.ndarray], expected_digests: Sequence[bytes]) -> Dict[str, int]:
    """Raises AssertionError on the first violated constraint; returns counts of what was checked."""
    g = canon_all(gate)
    n = len(g)
    fixed = canon_all(consts)
    brk = [int(b) for b in breaks] + [n]
    assert brk[0] == 0
    # column of every stream index; columns must fit max_rows
    col_of = np.zeros(n, dtype=np.int32)
    for c in range(len(brk) - 1):
        assert brk[c + 1] - brk[c] <= max_rows, f"gate column {c} overflows max_rows"
        col_of[brk[c]:brk[c + 1]] = c
    # (i) gates
    rows = np.nonzero(selectors)[0]
    for i in rows.tolist():
        assert i + 3 < n and col_of[i] == col_of[i + 3], f"gate at stream {i} straddles a column"
        assert (g[i] + g[i + 1] * g[i + 2] - g[i + 3]) % P == 0, f"gate identity fails at stream {i}"
    # (ii) copies
    d_c = canon_all(dense)
    s_c = canon_all(spread)

    def val(kind, idx):
        if kind == CP_GATE:
            return g[idx]
        if kind == CP_FIXED:
            return fixed[idx]
        if kind == CP_DENSE:
            return d_c[idx]
        if kind == CP_SPREAD:
            return s_c[idx]
        raise AssertionError(f"bad copy kind {kind}")

    for ak, ai, bk, bi in copies.tolist():
        assert val(ak, ai) == val(bk, bi), f"copy constraint fails: ({ak},{ai}) != ({bk},{bi})"
    for k in range(len(d_c)):
        assert d_c[k] == g[int(limb_gate_dense[k])], f"dense limb {k} copy fails"
        assert s_c[k] == g[int(limb_gate_spread[k])], f"spread limb {k} copy fails"
    # (iii) lookups
    for i in lookup_idx.tolist():
        assert g[i] < (1 << lookup_bits), f"range lookup fails at stream {i}: {g[i]}"
    table = {v: spread_bits(v) for v in range(1 << limb_bits)}
    for k in range(len(d_c)):
        assert d_c[k] in table and table[d_c[k]] == s_c[k], f"spread lookup fails at limb {k}"
    # (iv) instance
    for idxs, exp in zip(output_bytes_idx, expected_digests):
        got = bytes(g[int(i)] for i in idxs)
        assert got == exp, f"digest mismatch: {got.hex()} != {exp.hex()}"
    return dict(gates=len(rows), copies=len(copies) + 2 * len(d_c), lookups=len(lookup_idx), spread_rows=len(d_c), cells=n)


def verify_region(reg, expected_digests: Sequence[bytes]) -> Dict[str, int]:
    """Verify an oracle.Region."""
    return verify(gate=reg.gate, selectors=reg.selectors, breaks=reg.breaks, lookup_idx=reg.lookup_idx, dense=reg.dense, spread=reg.spread,
                  limb_gate_dense=reg.limb_gate_dense, limb_gate_spread=reg.limb_gate_spread, copies=reg.copies, consts=reg.consts,
                  lookup_bits=reg.cfg.lookup_bits, limb_bits=reg.cfg.limb_bits, max_rows=reg.cfg.max_rows,
                  output_bytes_idx=reg.output_bytes_idx, expected_digests=expected_digests)


def sha256(msg: bytes) -> bytes:
    return hashlib.sha256(msg).digest()
