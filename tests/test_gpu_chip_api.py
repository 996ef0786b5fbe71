"""The reference-shaped chip API (csrc/chip_api.hpp: configure / new_context / digest / range / load / range.finalize with the
signatures of reference src/lib.rs:49-76, 351-368) driven like the reference's TestCircuit (lib.rs:400-494) from C++, with a
recording Region.  The region is rebuilt here from the recorder's file ALONE -- i.e. from what h2sha_get_shape,
h2sha_get_handles and h2sha_export_instance delivered through the facade -- and then (a) accepted by the MockProver-style
checker, (b) compared cell by cell, selector by selector, copy by copy with the oracle's synthesis of the same inputs."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from oracle import mock_prover as MP
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GATE, LOOKUP, DENSE, SPREAD, FIXED, SELECTOR = range(6)


def _build(pkg):
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_chip_api")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_chip_api.cc"),
                           "-L" + libdir, "-lh2sha_b200", "-Xlinker", "-rpath," + libdir, "-cudart", "shared"])
    return exe


def _parse(path):
    w = np.fromfile(path, dtype=np.uint32)
    ev = dict(advice={}, fixed={}, selectors=[], copies=[], tables={}, instance=[], order=[])
    i = 0
    while i < len(w):
        t = int(w[i])
        if t == 1:
            kind, idx, row, has = (int(x) for x in w[i + 1:i + 5]); i += 5
            val = None
            if has:
                val = w[i:i + 8].view(np.uint64).copy(); i += 8
            assert (kind, idx, row) not in ev["advice"], "advice cell assigned twice"
            ev["advice"][(kind, idx, row)] = val
            ev["order"].append((kind, idx, row))
        elif t == 2:
            idx, row = int(w[i + 1]), int(w[i + 2]); ev["fixed"][(idx, row)] = w[i + 3:i + 11].view(np.uint64).copy(); i += 11
        elif t == 3:
            ev["selectors"].append((int(w[i + 1]), int(w[i + 2]))); i += 3
        elif t == 4:
            ev["copies"].append(tuple(int(x) for x in w[i + 1:i + 7])); i += 7
        elif t == 5:
            ln, nc, nr = (int(x) for x in w[i + 1:i + 4]); i += 4
            name = bytes(int(x) for x in w[i:i + ln]).decode(); i += ln
            ev["tables"][name] = w[i:i + 2 * nc * nr].view(np.uint64).reshape(nc, nr).copy(); i += 2 * nc * nr
        elif t == 6:
            ev["instance"].append(tuple(int(x) for x in w[i + 1:i + 5])); i += 5
        else:
            raise AssertionError(f"bad event tag {t} at word {i}")
    return ev


def _rebuild(ev, with_values=True):
    """(column,row) -> stream arrays, from the recorded events only."""
    gcols = sorted({k[1] for k in ev["advice"] if k[0] == GATE})
    col_len = [1 + max(k[2] for k in ev["advice"] if k[0] == GATE and k[1] == c) for c in gcols]
    breaks = np.concatenate([[0], np.cumsum(col_len)[:-1]]).astype(np.uint32)
    n_gate = int(sum(col_len))
    sidx = lambda c, r: int(breaks[c]) + r
    gate = np.zeros((n_gate, 4), np.uint64)
    if with_values:
        for (k, c, r), v in ev["advice"].items():
            if k == GATE:
                gate[sidx(c, r)] = v
    selectors = np.zeros(n_gate, np.uint8)
    for c, r in ev["selectors"]:
        selectors[sidx(c, r)] = 1
    n_fixed = len(ev["fixed"])
    consts = np.array([ev["fixed"][(0, k)] for k in range(n_fixed)], dtype=np.uint64)
    nc = 1 + max(k[1] for k in ev["advice"] if k[0] == DENSE)
    n_limb = sum(1 for k in ev["advice"] if k[0] == DENSE)
    dense = np.zeros((n_limb, 4), np.uint64); spread = np.zeros((n_limb, 4), np.uint64)
    if with_values:
        for (k, c, r), v in ev["advice"].items():
            if k == DENSE:
                dense[r * nc + c] = v
            elif k == SPREAD:
                spread[r * nc + c] = v
    n_lk = sum(1 for k in ev["advice"] if k[0] == LOOKUP)
    lookup_idx = np.full(n_lk, -1, np.int64); limb_d = np.full(n_limb, -1, np.int64); limb_s = np.full(n_limb, -1, np.int64)
    copies = []
    for ak, ai, ar, bk, bi, br in ev["copies"]:
        if ak == LOOKUP:
            assert bk == GATE and ai == 0
            lookup_idx[ar] = sidx(bi, br)
        elif ak == DENSE:
            assert bk == GATE; limb_d[ar * nc + ai] = sidx(bi, br)
        elif ak == SPREAD:
            assert bk == GATE; limb_s[ar * nc + ai] = sidx(bi, br)
        else:
            a = (MP.CP_GATE, sidx(ai, ar)) if ak == GATE else (MP.CP_FIXED, ar)
            b = (MP.CP_GATE, sidx(bi, br)) if bk == GATE else (MP.CP_FIXED, br)
            copies.append(a + b)
    assert (lookup_idx >= 0).all() and (limb_d >= 0).all() and (limb_s >= 0).all()
    out_idx = [np.array([sidx(c, r) for (k, c, r, i) in ev["instance"][32 * d:32 * d + 32]], dtype=np.uint32) for d in range(len(ev["instance"]) // 32)]
    lookup_col = np.array([ev["advice"][(LOOKUP, 0, r)] for r in range(n_lk)], dtype=np.uint64) if with_values else None
    return dict(gate=gate, selectors=selectors, breaks=breaks, consts=consts, dense=dense, spread=spread, lookup_idx=lookup_idx.astype(np.uint32),
                limb_d=limb_d.astype(np.uint32), limb_s=limb_s.astype(np.uint32), copies=np.array(copies, dtype=np.uint32), out_idx=out_idx, lookup_col=lookup_col)


def _inputs(tc):
    if tc in (0, 1):
        return [b"abc", b""], [0, 0]
    if tc == 3:
        return [b"\x01" * 56, b"\x00\x00\x00"], [0, 0]
    return [bytes(range(192)), bytes((i + 64) & 255 for i in range(192))], [128, 128]


def _check_against_oracle(rb, tc, with_values=True):
    msgs, pre = _inputs(tc)
    reg = O.synthesize(O.OracleConfig(max_variable_byte_sizes=(128, 128)), msgs, pre)
    assert (rb["breaks"] == reg.breaks).all() and rb["gate"].shape == reg.gate.shape
    assert (rb["selectors"] == reg.selectors).all()
    assert (rb["lookup_idx"] == reg.lookup_idx).all()
    assert (rb["limb_d"] == reg.limb_gate_dense).all() and (rb["limb_s"] == reg.limb_gate_spread).all()
    canon = lambda a: [int(x[0]) | int(x[1]) << 64 | int(x[2]) << 128 | int(x[3]) << 192 for x in a]
    assert canon(rb["consts"]) == [O.mont_to_int(x) for x in reg.consts], "fixed column differs"
    assert {tuple(c) for c in rb["copies"].tolist()} == {tuple(c) for c in reg.copies.tolist()}, "copy constraints differ"
    for d in range(2):
        assert (rb["out_idx"][d] == reg.output_bytes_idx[d]).all()
    if with_values:
        assert (rb["gate"] == reg.gate).all(), "gate cells differ from the oracle"
        assert (rb["dense"] == reg.dense).all() and (rb["spread"] == reg.spread).all()
        assert (rb["lookup_col"] == reg.gate[reg.lookup_idx]).all()
    return reg


def test_chip_api_compiles_and_keygen_replay_matches_oracle_shape(pkg, tmp_path):
    """CPU: the facade builds; with a plan-only engine (no GPU) the shape-only replay (what keygen_vk / keygen_pk need,
    benches/digest.rs:136-137) emits the oracle's selectors, fixed cells, copy constraints and both tables."""
    exe = _build(pkg)
    out = subprocess.run([exe, str(tmp_path / "k.bin"), "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "total advice cells: 279797" in out.stdout and "lookup cells used: 13382" in out.stdout   # the reference's 3-column budget
    ev = _parse(tmp_path / "k.bin")
    assert all(v is None for v in ev["advice"].values())
    rb = _rebuild(ev, with_values=False)
    _check_against_oracle(rb, 0, with_values=False)
    t = ev["tables"]
    assert t["range lookup table"].shape == (1, 1 << 16) and (t["range lookup table"][0] == np.arange(1 << 16)).all()
    sp = t["spread table"]
    assert sp.shape == (2, 256) and (sp[0] == np.arange(256)).all() and all(int(sp[1, i]) == MP.spread_bits(i) for i in range(256))


@pytest.mark.gpu
@pytest.mark.parametrize("tc", [1, 3, 4, 14])
def test_reference_test_circuit_through_the_chip_api(pkg, tmp_path, tc):
    exe = _build(pkg)
    out = subprocess.run([exe, str(tmp_path / "r.bin"), str(tc)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    ev = _parse(tmp_path / "r.bin")
    rb = _rebuild(ev)
    tcase = 4 if tc == 14 else tc
    msgs, _ = _inputs(tcase)
    consts_mont = np.array([O.int_to_mont(int(a) | int(b) << 64 | int(c) << 128 | int(d) << 192) for a, b, c, d in rb["consts"]], dtype=np.uint64)
    stats = MP.verify(gate=rb["gate"], selectors=rb["selectors"], breaks=rb["breaks"], lookup_idx=rb["lookup_idx"], dense=rb["dense"], spread=rb["spread"],
                      limb_gate_dense=rb["limb_d"], limb_gate_spread=rb["limb_s"], copies=rb["copies"], consts=consts_mont, lookup_bits=16, limb_bits=8,
                      max_rows=(1 << 17) - 9, output_bytes_idx=rb["out_idx"], expected_digests=[hashlib.sha256(m).digest() for m in msgs])
    assert stats["gates"] > 0
    _check_against_oracle(rb, tcase)
