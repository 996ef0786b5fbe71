"""CPU restatement of halo2's lookup-argument pre-work for the two lookups of the SHA-256 chip.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py): imported by tests/ as the checker of
`h2sha_lookup_multiplicities` / `h2sha_permute_lookup`, never by the product package.

What it follows
  * the lookups themselves are the reference's:
      - "spread lookup" per column pair c: inputs (denses[c], spreads[c]) against (table_dense, table_spread),
        no selector, every row of the region is looked up (/root/reference/src/spread.rs:53-62); the table is
        SpreadConfig::load (spread.rs:165-194);
      - the range lookup of halo2-base's RangeConfig on the lookup advice column(s) that `range.finalize`
        fills (/root/reference/src/lib.rs:409-418, 442, 469) -- dependency zkmove/halo2-lib rev 40ba7e3, not vendored.
  * the permutation is PSE halo2_proofs `plonk/lookup/prover.rs`: `compress_expressions` (fold acc*theta + expr)
    and `permute_expression_pair` (sort the input, BTreeMap of leftover table elements, repeated rows filled by
    popping from the END of the repeated-row list).  The dependency is absent from /root/reference (Cargo.toml:10-17),
    so this is a restatement of the published algorithm: **parity unpinned**, like the cell placement.  What IS
    checked (tests/test_lookup_prework.py): the three properties the lookup argument's constraints need --
    A' is a permutation of A, S' a permutation of S, and on every row A'[i] == S'[i] or A'[i] == A'[i-1]
    (A'[0] == S'[0]) -- hold for the oracle's and the GPU's output.
  * `Ord` of bn256::Fr compares canonical (non-Montgomery) integers (halo2curves `impl Ord for Fr`).
  * table columns are padded to `usable_rows` with their first row's value (halo2 `SimpleTableLayouter`
    default value), so the range table holds 0..2^lookup_bits-1 then zeros, the spread table its
    2^num_bits_lookup rows then (0, 0); never-assigned advice rows hold 0.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

P = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def spread_bits(x: int, bits: int) -> int:
    """spread.rs:171-180: bit b of x -> bit 2b."""
    return sum(((x >> b) & 1) << (2 * b) for b in range(bits))


def compress(exprs: Sequence[int], theta: int) -> int:
    """halo2 lookup prover `compress_expressions`: fold(0, |acc, e| acc * theta + e)."""
    acc = 0
    for e in exprs:
        acc = (acc * theta + e) % P
    return acc


def permute_expression_pair(input_expression: Sequence[int], table_expression: Sequence[int], usable_rows: int) -> Tuple[List[int], List[int]]:
    """halo2 `permute_expression_pair` without the trailing blinding rows (those are random)."""
    permuted_input = sorted(input_expression[:usable_rows])
    leftover: Dict[int, int] = {}
    for t in table_expression[:usable_rows]:
        leftover[t] = leftover.get(t, 0) + 1
    permuted_table = [0] * usable_rows
    repeated_rows: List[int] = []
    for row, v in enumerate(permuted_input):
        if row == 0 or v != permuted_input[row - 1]:
            permuted_table[row] = v
            if leftover.get(v, 0) == 0:
                raise ValueError(f"input value {v} is not in the table (ConstraintSystemFailure)")
            leftover[v] -= 1
        else:
            repeated_rows.append(row)
    for coeff in sorted(leftover):            # BTreeMap iteration order
        for _ in range(leftover[coeff]):
            permuted_table[repeated_rows.pop()] = coeff
    assert not repeated_rows
    return permuted_input, permuted_table


def check_permuted(a: Sequence[int], s: Sequence[int], a_p: Sequence[int], s_p: Sequence[int]) -> None:
    """The properties the lookup argument constrains (product argument + the two row rules)."""
    assert sorted(a) == sorted(a_p), "A' is not a permutation of A"
    assert sorted(s) == sorted(s_p), "S' is not a permutation of S"
    assert a_p[0] == s_p[0]
    for i in range(1, len(a_p)):
        assert a_p[i] == s_p[i] or a_p[i] == a_p[i - 1], f"row {i}"


def range_lookup_columns(lookup_col: Sequence[int], usable_rows: int, lookup_bits: int) -> Tuple[List[int], List[int]]:
    """(input, table) expressions of the range lookup over the usable rows; `lookup_col` = the assigned prefix."""
    a = list(lookup_col[:usable_rows]) + [0] * max(0, usable_rows - len(lookup_col))
    n = 1 << lookup_bits
    s = list(range(n))[:usable_rows] + [0] * max(0, usable_rows - n)
    return a, s


def spread_lookup_columns(dense_col: Sequence[int], spread_col: Sequence[int], usable_rows: int, num_bits_lookup: int, theta: int) -> Tuple[List[int], List[int]]:
    """Compressed (input, table) expressions of one spread lookup over the usable rows."""
    assert len(dense_col) == len(spread_col)
    a = [compress((d, s), theta) for d, s in zip(dense_col[:usable_rows], spread_col[:usable_rows])]
    a += [0] * max(0, usable_rows - len(a))
    n = 1 << num_bits_lookup
    s = [compress((i, spread_bits(i, num_bits_lookup)), theta) for i in range(n)][:usable_rows] + [0] * max(0, usable_rows - n)
    return a, s


def multiplicities(values: Sequence[int], n_table: int, usable_rows: int) -> List[int]:
    """How often each table row is hit by the first usable_rows rows of a column whose assigned prefix is `values`."""
    m = [0] * n_table
    for v in values[:usable_rows]:
        m[v] += 1
    m[0] += max(0, usable_rows - len(values))
    return m
