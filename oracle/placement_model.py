"""Third, independent restatement of cell PLACEMENT only (TEST INFRASTRUCTURE ONLY).

Written from SURVEY.md 8a Table B (the recalled halo2-base v0.2.x op -> cell patterns) and the reference's call order
(/root/reference/src/lib.rs:71-349, src/compression.rs:19-882, src/spread.rs:76-233) alone, without values, symbols or
templates: a cursor that appends op lengths.  It yields the three things a shared misreading of Table B by the planner
(csrc/planner.cc) and the oracle (oracle/h2sha_oracle.c) would most likely get wrong together only if this file got it
wrong the same way: gate-stream length, column break points, and the gate-stream index of every cell pushed to
`cells_to_lookup`, in push order.  tests/test_placement_model.py asserts that all three restatements agree.
"""
from dataclasses import dataclass, field
from typing import List


@dataclass
class Cursor:
    max_rows: int
    lookup_bits: int
    n: int = 0                    # cells in the gate stream
    row: int = 0
    breaks: List[int] = field(default_factory=lambda: [0])
    lookups: List[int] = field(default_factory=list)     # gate-stream index of each pushed cell
    limbs: int = 0                # spread_limb calls
    zero: bool = False

    def op(self, length: int) -> int:
        """Append one op; returns the stream index of its first cell.  Table B: `if row + len >= max_rows {row = 0; col += 1}`."""
        if self.row + length >= self.max_rows:
            self.row = 0
            self.breaks.append(self.n)
        first = self.n
        self.n += length
        self.row += length
        return first

    # ---- halo2-base ops (Table B) ----
    def witness(self) -> int:
        return self.op(1)

    def load_zero(self):
        if not self.zero:
            self.op(1)
            self.zero = True

    def gate4(self, out_at: int) -> int:       # add [3], sub [0], neg [1], mul [3], mul_add [3]
        return self.op(4) + out_at

    def range_check(self, cell: int, bits: int) -> int:
        lb = self.lookup_bits
        k, rem = -(-bits // lb), bits % lb
        last = cell
        if k == 1:
            self.lookups.append(cell)
        else:
            first = self.op(1 + 3 * (k - 1))            # W l0, then (W l_i, C 2^(lb i), W acc) per further limb
            pos = [first] + [first + 1 + 3 * (i - 1) for i in range(1, k)]
            self.lookups.extend(pos)
            last = pos[-1]
        if rem:
            last = self.op(4) + 3                       # *C 0, last, C 2^(lb-rem), W shifted
            self.lookups.append(last)
        return last

    def is_zero(self) -> int:
        return self.op(8)

    def is_less_than_safe(self, a: int, b_bits: int = 7):
        lb = self.lookup_bits
        range_bits = -(-b_bits // lb) * lb
        self.range_check(a, range_bits)
        k = -(-range_bits // lb)
        first = self.op(7)                              # *W a+2^pb-b, C b, C 1, *W a+2^pb, C -2^pb, C 1, a
        self.range_check(first, k * lb + lb)
        self.is_zero()


def spread(c: Cursor, limb_bits: int):
    nl = 16 // limb_bits
    for _ in range(nl):
        c.witness()
    c.load_zero()
    for _ in range(nl):
        c.gate4(3)
    c.load_zero()
    for _ in range(nl):
        c.limbs += 1
        c.witness()                                     # spread_limb: gate.load_witness(spread)
        c.gate4(3)


def state_to_spread(c, lb):
    c.witness(); c.witness(); c.gate4(3)
    spread(c, lb); spread(c, lb)


def mod_u32(c):
    lo = c.witness(); c.witness()
    c.range_check(lo, 32)
    c.gate4(3)


def decompose_even_odd(c):
    e = c.witness(); o = c.witness()
    c.range_check(e, 16); c.range_check(o, 16)


def even_odd_check(c, lb):
    spread(c, lb); spread(c, lb); c.gate4(3)


def sigma(c, lb):
    for _ in range(4):
        c.witness()
    for _ in range(4):
        c.gate4(3)                                      # 3 recompose + x_composed
    c.load_zero()
    for _ in range(4):
        c.gate4(3)
    lo = c.witness(); hi = c.witness()
    c.range_check(lo, 32); c.range_check(hi, 32)
    c.gate4(3)
    decompose_even_odd(c); decompose_even_odd(c)
    even_odd_check(c, lb); even_odd_check(c, lb)
    c.gate4(3)


def ch(c, lb):
    c.gate4(3); c.gate4(3)                              # p_lo, p_hi
    c.gate4(1); c.gate4(1)                              # neg, neg
    for _ in range(4):
        c.gate4(3)                                      # two three_adds
    for _ in range(4):
        decompose_even_odd(c)
    for _ in range(4):
        even_odd_check(c, lb)
    c.gate4(3); c.gate4(3); c.gate4(3)


def maj(c, lb):
    for _ in range(4):
        c.gate4(3)
    decompose_even_odd(c); decompose_even_odd(c)
    even_odd_check(c, lb); even_odd_check(c, lb)
    c.gate4(3)


def compression(c, lb):
    for _ in range(16):
        c.load_zero()
        for _ in range(4):
            c.gate4(3)
    for _ in range(16):
        state_to_spread(c, lb)
    for _ in range(48):
        sigma(c, lb); sigma(c, lb)
        c.gate4(3); c.gate4(3); c.gate4(3)
        mod_u32(c)
        state_to_spread(c, lb)
    for _ in range(6):
        state_to_spread(c, lb)
    c.load_zero(); c.load_zero()
    for _ in range(64):
        sigma(c, lb); ch(c, lb)
        for _ in range(4):
            c.gate4(3)
        mod_u32(c)
        sigma(c, lb); maj(c, lb)
        c.gate4(3); mod_u32(c)
        c.gate4(3); mod_u32(c); state_to_spread(c, lb)
        c.gate4(3); mod_u32(c); state_to_spread(c, lb)
    for _ in range(8):
        c.gate4(3); mod_u32(c)


def digest(c: Cursor, max_bytes: int, limb_bits: int, input_range_check: bool):
    rounds = max_bytes // 64
    c.witness(); c.witness()                            # input length, num_round
    c.gate4(3)                                          # mul
    c.gate4(3)                                          # add 9
    padding = c.gate4(0)                                # sub
    c.is_less_than_safe(padding)
    c.witness()                                         # precomputed round
    c.gate4(0)                                          # target round
    for _ in range(8):
        c.witness()
    bytes_ = [c.witness() for _ in range(max_bytes)]
    if input_range_check:
        for b in bytes_:
            c.range_check(b, 8)
    for _ in range(rounds):
        compression(c, limb_bits)
    c.load_zero()
    for _ in range(rounds + 1):
        c.op(4); c.is_zero()                            # is_equal = sub + is_zero
        for _ in range(8):
            c.op(8)                                     # select
    for _ in range(8):
        for _ in range(4):
            b = c.witness()
            c.range_check(b, 8)
        c.load_zero()
        for _ in range(4):
            c.gate4(3)


def place(max_variable_byte_sizes, max_rows=(1 << 17) - 9, lookup_bits=16, limb_bits=8, input_range_check=True) -> Cursor:
    c = Cursor(max_rows=max_rows, lookup_bits=lookup_bits)
    for m in max_variable_byte_sizes:
        digest(c, m, limb_bits, input_range_check)
    return c
