"""ctypes front-end of the CPU oracle (oracle/h2sha_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package never imports it.

The oracle restates /root/reference's witness generation (src/lib.rs:71-349,
src/compression.rs:19-882, src/spread.rs:76-233) cell by cell; see the C file's header for the
parity status ("digests pinned, constraint-consistent, placement parity unpinned").
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libh2sha_oracle.so")

P = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
R = (1 << 256) % P
R_INV = pow(R, -1, P)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "h2sha_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Cfg(C.Structure):
    _fields_ = [
        ("n_digests", C.c_uint32),
        ("max_bytes", C.POINTER(C.c_uint32)),
        ("max_rows", C.c_uint32),
        ("lookup_bits", C.c_uint32),
        ("limb_bits", C.c_uint32),
        ("spread_cols", C.c_uint32),
        ("is_input_range_check", C.c_uint32),
    ]


class _Layout(C.Structure):
    _fields_ = [
        ("n_gate_cols", C.c_uint32),
        ("gate_col_rows", C.c_uint32),
        ("n_lookup_cols", C.c_uint32),
        ("lookup_col_rows", C.c_uint32),
        ("spread_rows", C.c_uint32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.h2o_synthesize.restype = C.c_void_p
        L.h2o_synthesize.argtypes = [C.POINTER(_Cfg), C.POINTER(C.c_char_p), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int, C.POINTER(C.c_int)]
        L.h2o_free.argtypes = [C.c_void_p]
        L.h2o_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        for name, rt in [("h2o_gate", C.c_uint64), ("h2o_selectors", C.c_uint8), ("h2o_breaks", C.c_uint32), ("h2o_lookup_idx", C.c_uint32),
                         ("h2o_dense", C.c_uint64), ("h2o_spread", C.c_uint64), ("h2o_limb_gate_dense", C.c_uint32),
                         ("h2o_limb_gate_spread", C.c_uint32), ("h2o_copies", C.c_uint32), ("h2o_consts", C.c_uint64)]:
            f = getattr(L, name)
            f.restype = C.POINTER(rt)
            f.argtypes = [C.c_void_p]
        L.h2o_digest.restype = C.POINTER(C.c_uint8)
        L.h2o_digest.argtypes = [C.c_void_p, C.c_uint32]
        L.h2o_input_len_idx.restype = C.c_uint32
        L.h2o_input_len_idx.argtypes = [C.c_void_p, C.c_uint32]
        L.h2o_input_bytes_idx.restype = C.POINTER(C.c_uint32)
        L.h2o_input_bytes_idx.argtypes = [C.c_void_p, C.c_uint32]
        L.h2o_output_bytes_idx.restype = C.POINTER(C.c_uint32)
        L.h2o_output_bytes_idx.argtypes = [C.c_void_p, C.c_uint32]
        L.h2o_emit.restype = C.c_int
        L.h2o_emit.argtypes = [C.c_void_p, C.POINTER(_Layout), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.h2o_batch.restype = C.c_int
        L.h2o_batch.argtypes = [C.POINTER(_Cfg), C.POINTER(_Layout), C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.h2o_fr_from_u64.argtypes = [C.c_uint64, C.POINTER(C.c_uint64)]
        L.h2o_fr_canon.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.h2o_fr_consts.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


@dataclass
class OracleConfig:
    """Mirrors the constructor arguments of the reference (lib.rs:49-56, 409-428)."""
    max_variable_byte_sizes: Sequence[int] = (64,)
    max_rows: int = (1 << 17) - 9           # range.gate.max_rows for k=17 (lib.rs:355, 417)
    lookup_bits: int = 16                   # LOOKUP_BITS (lib.rs:493)
    limb_bits: int = 8                      # num_bits_lookup (lib.rs:425)
    spread_cols: int = 2                    # num_advice_columns (lib.rs:426)
    is_input_range_check: bool = True       # lib.rs:427

    def _c(self):
        arr = (C.c_uint32 * len(self.max_variable_byte_sizes))(*self.max_variable_byte_sizes)
        cfg = _Cfg(len(self.max_variable_byte_sizes), arr, self.max_rows, self.lookup_bits, self.limb_bits, self.spread_cols,
                   1 if self.is_input_range_check else 0)
        cfg._keep = arr
        return cfg


@dataclass
class Layout:
    n_gate_cols: int
    gate_col_rows: int
    n_lookup_cols: int
    lookup_col_rows: int
    spread_rows: int

    def _c(self):
        return _Layout(self.n_gate_cols, self.gate_col_rows, self.n_lookup_cols, self.lookup_col_rows, self.spread_rows)


def _np(ptr, n, dtype, width=1):
    if n == 0:
        return np.zeros((0, width) if width > 1 else (0,), dtype=dtype)
    a = np.ctypeslib.as_array(ptr, shape=(n * width,)).copy()
    return a.reshape(n, width) if width > 1 else a


@dataclass
class Region:
    """Everything one synthesized Context holds (one `assign_region` of the reference's test circuit)."""
    cfg: OracleConfig
    gate: np.ndarray            # [n_gate,4] u64 Montgomery limbs, stream order
    selectors: np.ndarray       # [n_gate] u8
    breaks: np.ndarray          # stream index at which gate column c starts
    lookup_idx: np.ndarray      # cells_to_lookup as gate stream indices
    dense: np.ndarray           # [n_limb,4] spread-table dense column cells, limb order
    spread: np.ndarray          # [n_limb,4]
    limb_gate_dense: np.ndarray
    limb_gate_spread: np.ndarray
    copies: np.ndarray          # [n_copy,4] (a_kind,a_idx,b_kind,b_idx)
    consts: np.ndarray          # [n_const,4] fixed column, first-use order
    digests: List[bytes] = field(default_factory=list)
    input_len_idx: List[int] = field(default_factory=list)
    input_bytes_idx: List[np.ndarray] = field(default_factory=list)
    output_bytes_idx: List[np.ndarray] = field(default_factory=list)

    @property
    def n_gate(self):
        return self.gate.shape[0]

    def layout(self, align: int = 8) -> Layout:
        ends = list(self.breaks[1:]) + [self.n_gate]
        rows = max(int(e) - int(s) for s, e in zip(self.breaks, ends))
        rows = (rows + align - 1) // align * align
        n_lk = len(self.lookup_idx)
        lk_cols = max(1, -(-n_lk // self.cfg.max_rows))
        lk_rows = min(n_lk, self.cfg.max_rows)
        lk_rows = (lk_rows + align - 1) // align * align
        n_limb = self.dense.shape[0]
        srows = -(-n_limb // self.cfg.spread_cols)
        srows = (srows + align - 1) // align * align
        return Layout(len(self.breaks), rows, lk_cols, lk_rows, srows)


def synthesize(cfg: OracleConfig, msgs: Sequence[bytes], pre_lens: Optional[Sequence[int]] = None, record_shape: bool = True):
    """Run cfg.n_digests digest() calls in one Context; returns (Region, handle-free)."""
    L = lib()
    D = len(cfg.max_variable_byte_sizes)
    assert len(msgs) == D
    ccfg = cfg._c()
    bufs = [C.create_string_buffer(bytes(m), max(1, len(m))) for m in msgs]
    msg_ptrs = (C.c_char_p * D)(*[C.cast(b, C.c_char_p) for b in bufs])
    lens = (C.c_uint32 * D)(*[len(m) for m in msgs])
    pl = (C.c_uint32 * D)(*(pre_lens if pre_lens is not None else [0] * D))
    err = C.c_int(0)
    h = L.h2o_synthesize(C.byref(ccfg), msg_ptrs, lens, pl, 1 if record_shape else 0, C.byref(err))
    if not h:
        raise ValueError(f"reference would panic (oracle code {err.value})")
    try:
        sz = (C.c_uint64 * 6)()
        L.h2o_sizes(h, sz)
        n_gate, n_lk, n_limb, n_copy, n_const, n_cols = [int(x) for x in sz]
        reg = Region(
            cfg=cfg,
            gate=_np(L.h2o_gate(h), n_gate, np.uint64, 4),
            selectors=_np(L.h2o_selectors(h), n_gate, np.uint8),
            breaks=_np(L.h2o_breaks(h), n_cols, np.uint32),
            lookup_idx=_np(L.h2o_lookup_idx(h), n_lk, np.uint32),
            dense=_np(L.h2o_dense(h), n_limb, np.uint64, 4),
            spread=_np(L.h2o_spread(h), n_limb, np.uint64, 4),
            limb_gate_dense=_np(L.h2o_limb_gate_dense(h), n_limb, np.uint32),
            limb_gate_spread=_np(L.h2o_limb_gate_spread(h), n_limb, np.uint32),
            copies=_np(L.h2o_copies(h), n_copy, np.uint32, 4),
            consts=_np(L.h2o_consts(h), n_const, np.uint64, 4),
        )
        for d in range(D):
            reg.digests.append(bytes(np.ctypeslib.as_array(L.h2o_digest(h, d), shape=(32,))))
            reg.input_len_idx.append(int(L.h2o_input_len_idx(h, d)))
            reg.input_bytes_idx.append(_np(L.h2o_input_bytes_idx(h, d), cfg.max_variable_byte_sizes[d], np.uint32))
            reg.output_bytes_idx.append(_np(L.h2o_output_bytes_idx(h, d), 32, np.uint32))
        return reg
    finally:
        L.h2o_free(h)


def emit(cfg: OracleConfig, msgs: Sequence[bytes], pre_lens, layout: Layout):
    """Column-major witness of one instance + checksums, same contract as the CUDA engine."""
    out = batch(cfg, layout, [list(msgs)], [list(pre_lens) if pre_lens is not None else [0] * len(msgs)], want_cells=True, n_threads=1)
    return out


def batch(cfg: OracleConfig, layout: Layout, instances: Sequence[Sequence[bytes]], pre_lens: Optional[Sequence[Sequence[int]]] = None,
          want_cells: bool = False, n_threads: int = 1):
    """instances[i][d] = message d of instance i.  Returns dict(digests [n_msgs,32] u8, checksums [n_inst,4] u64,
    and, if want_cells, gate/lookup/spread arrays [n_inst, cols, rows, 4] u64)."""
    L = lib()
    D = len(cfg.max_variable_byte_sizes)
    n = len(instances)
    flat = [bytes(m) for inst in instances for m in inst]
    assert len(flat) == n * D
    lens = np.array([len(m) for m in flat], dtype=np.uint32)
    offs = np.zeros(n * D, dtype=np.uint64)
    if n * D > 1:
        offs[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    blob = np.frombuffer(b"".join(flat) + b"\0", dtype=np.uint8).copy()
    pl = np.array([p for inst in pre_lens for p in inst], dtype=np.uint32) if pre_lens is not None else np.zeros(n * D, dtype=np.uint32)
    return batch_packed(cfg, layout, n, blob, offs, lens, pl, want_cells, n_threads)


def batch_packed(cfg: OracleConfig, layout: Layout, n_inst: int, blob: np.ndarray, offs: np.ndarray, lens: np.ndarray,
                 pre_lens: np.ndarray, want_cells: bool = False, n_threads: int = 1):
    L = lib()
    D = len(cfg.max_variable_byte_sizes)
    ccfg, cl = cfg._c(), layout._c()
    digests = np.zeros((n_inst * D, 32), dtype=np.uint8)
    cks = np.zeros((n_inst, 4), dtype=np.uint64)
    gate = lookup = spread = None
    if want_cells:
        gate = np.zeros((n_inst, layout.n_gate_cols, layout.gate_col_rows, 4), dtype=np.uint64)
        lookup = np.zeros((n_inst, layout.n_lookup_cols, layout.lookup_col_rows, 4), dtype=np.uint64)
        spread = np.zeros((n_inst, 2 * cfg.spread_cols, layout.spread_rows, 4), dtype=np.uint64)
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    pre_lens = np.ascontiguousarray(pre_lens, dtype=np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rc = L.h2o_batch(C.byref(ccfg), C.byref(cl), n_inst, p(blob), p(offs), p(lens), p(pre_lens), p(digests), p(cks), p(gate), p(lookup),
                     p(spread), n_threads)
    if rc:
        raise ValueError(f"oracle batch failed (code {rc})")
    return dict(digests=digests, checksums=cks, gate=gate, lookup=lookup, spread=spread)


def mont_to_int(limbs) -> int:
    """[4] u64 Montgomery limbs -> canonical Python int."""
    v = int(limbs[0]) | (int(limbs[1]) << 64) | (int(limbs[2]) << 128) | (int(limbs[3]) << 192)
    return v * R_INV % P


def int_to_mont(v: int) -> np.ndarray:
    m = (v % P) * R % P
    return np.array([(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
