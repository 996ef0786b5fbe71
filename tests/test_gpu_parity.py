"""Parity tests proper (need a B200): the CUDA path, called through the C-ABI, against the oracle on the same inputs --
bit-exact for every advice, lookup and spread-column cell, the digests and the checksums -- plus a MockProver-style
pass over the GPU output using the product's own shape plan, and size-independent properties at BASELINE sizes."""
import hashlib
import os

import numpy as np
import pytest

from oracle import mock_prover as MP
from oracle import oracle as O

pytestmark = pytest.mark.gpu

P = O.P
NCPU = os.cpu_count() or 1


def _u64(t):
    return t.cpu().numpy().view(np.uint64)


def _oracle_cfg(kw):
    return O.OracleConfig(**kw)


def _engine(pkg, kw, **extra):
    return pkg.Sha256DynamicConfig.configure(list(kw["max_variable_byte_sizes"]), max_rows=kw.get("max_rows", (1 << 17) - 9),
                                             lookup_bits=kw.get("lookup_bits", 16), num_bits_lookup=kw.get("limb_bits", 8),
                                             num_advice_columns=kw.get("spread_cols", 2),
                                             is_input_range_check=kw.get("is_input_range_check", True), device=0, **extra)


def _compare(pkg, kw, instances, pre=None, threads=NCPU, **extra):
    cfg = _engine(pkg, kw, **extra)
    lay = cfg.layout
    res = cfg.digest_batch(instances, pre)
    olay = O.Layout(lay.n_gate_cols, lay.gate_col_rows, lay.n_lookup_cols, lay.lookup_col_rows, lay.spread_rows)
    ref = O.batch(_oracle_cfg(kw), olay, instances, pre, want_cells=True, n_threads=min(threads, len(instances)))
    assert (res.digests == ref["digests"]).all(), "digests differ"
    for name in ("gate", "lookup", "spread"):
        got = _u64(getattr(res, name))
        bad = np.argwhere((got != ref[name]).any(axis=-1))
        assert bad.size == 0, f"{name}: {len(bad)} cells differ, first at (instance, column, row) = {bad[0]}"
    assert (res.checksums == ref["checksums"]).all(), "checksums differ"
    cfg.close()
    return res, ref


def test_montgomery_conversion_matches_python_ints(pkg):
    """Device Barrett conversion v -> v * 2^256 mod p, against Python big ints: edge values and 2^20 random ones."""
    cfg = _engine(pkg, dict(max_variable_byte_sizes=(64,)))
    rng = np.random.default_rng(1)
    edge = [0, 1, 2, 255, 256, 0xFFFF, 0x10000, 0x55555555, 0xAAAAAAAA, 0xFFFFFFFF, 1 << 32, (1 << 35) - 1, 0x5555555555555555,
            0xAAAAAAAAAAAAAAAA, (1 << 63) - 1, 1 << 63, (1 << 64) - 1, (1 << 64) - 2]
    edge += [(1 << k) for k in range(64)] + [(1 << k) - 1 for k in range(1, 65)]
    vals = np.concatenate([np.array(edge, dtype=np.uint64), rng.integers(0, 1 << 63, size=1 << 20, dtype=np.uint64) * np.uint64(2) + np.uint64(1),
                           rng.integers(0, 1 << 32, size=1 << 16, dtype=np.uint64), rng.integers(0, 1 << 16, size=1 << 12, dtype=np.uint64)])
    out = cfg.mont_from_u64(vals)
    R = (1 << 256) % P
    idx = list(range(len(edge))) + list(rng.integers(0, len(vals), size=20000))
    for i in idx:
        got = sum(int(out[i, k]) << (64 * k) for k in range(4))
        assert got == int(vals[i]) * R % P, f"value {int(vals[i]):#x}"
    # all results are canonical (< p): compare the top limb cheaply for the whole batch
    assert (out[:, 3] <= np.uint64(P >> 192)).all()
    # the < 2^32 fast path: every 16-bit value, edge values, 2^20 random ones; and it must agree with the 64-bit path
    v32 = np.concatenate([np.arange(1 << 16, dtype=np.uint64), np.array([e for e in edge if e < (1 << 32)], dtype=np.uint64),
                          rng.integers(0, 1 << 32, size=1 << 20, dtype=np.uint64)])
    o32 = cfg.mont_from_u64(v32, path32=True)
    assert (o32 == cfg.mont_from_u64(v32)).all()
    for i in list(range(0, 1 << 16, 97)) + list(rng.integers(0, len(v32), size=5000)):
        assert sum(int(o32[i, k]) << (64 * k) for k in range(4)) == int(v32[i]) * R % P
    cfg.close()


def test_reference_test_vectors_two_digests_per_context(pkg):
    """The reference's TestCircuit (lib.rs:400-494): two digest() calls in one Context, max 128 each, k = 17."""
    kw = dict(max_variable_byte_sizes=(128, 128))
    instances = [[b"abc", b""], [b"\x00", b""], [b"\x01" * 56, b"\x00\x00\x00"]]
    res, _ = _compare(pkg, kw, instances)
    exp = ["ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad", "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855",
           "6e340b9cffb37a989ca544e6bb780a2c78901d3fb33738768511a30617afa01d", "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855",
           "51e14a913680f24c85fe3b0e2e5b57f7202f117bb214f8ffdd4ea0f4e921fd52", "709e80c88487a2411e1ee4dfb9f22a861492d20c4765150c0c794abd70f8147c"]
    assert [bytes(d).hex() for d in res.digests] == exp


def test_mock_prover_on_gpu_output(pkg):
    """MockProver pass on sampled instances: the GPU cells satisfy every gate, copy and lookup constraint of the
    product's own shape plan and the output-byte cells equal the digests (lib.rs:480-482, 525-526)."""
    kw = dict(max_variable_byte_sizes=(128, 128))
    cfg = _engine(pkg, kw, build_shape=True)
    rng = np.random.default_rng(5)
    instances = [[b"abc", b""], [bytes(rng.integers(0, 256, 119, dtype=np.uint8)), bytes(rng.integers(0, 256, 64, dtype=np.uint8))]]
    res = cfg.digest_batch(instances)
    lay, sh, brk = cfg.layout, cfg.shape(), cfg.breaks()
    gate = _u64(res.gate); lookup = _u64(res.lookup); spread = _u64(res.spread)
    ends = list(brk[1:]) + [lay.n_gate_cells]
    for i, inst in enumerate(instances):
        stream = np.concatenate([gate[i, c, : int(e) - int(s)] for c, (s, e) in enumerate(zip(brk, ends))])
        # the lookup column must hold copies of the looked-up gate cells, in push order (range.finalize)
        lk = np.concatenate([lookup[i, c] for c in range(lay.n_lookup_cols)])[: lay.n_lookup_cells]
        assert (lk == stream[sh.lookup_src]).all()
        nc = lay.n_spread_cols // 2
        n = np.arange(lay.n_spread_limbs)
        dense = spread[i, n % nc, n // nc]
        spr = spread[i, nc + n % nc, n // nc]
        consts_mont = np.array([O.int_to_mont(int(a) | int(b) << 64 | int(c) << 128 | int(d) << 192) for a, b, c, d in sh.fixed], dtype=np.uint64)
        stats = MP.verify(gate=stream, selectors=sh.selectors, breaks=brk, lookup_idx=sh.lookup_src, dense=dense, spread=spr,
                          limb_gate_dense=sh.limb_dense_src, limb_gate_spread=sh.limb_spread_src, copies=sh.copies, consts=consts_mont,
                          lookup_bits=16, limb_bits=8, max_rows=(1 << 17) - 9,
                          output_bytes_idx=[cfg.handles(d).output_bytes for d in range(2)],
                          expected_digests=[hashlib.sha256(m).digest() for m in inst])
        assert stats["gates"] == lay.n_selectors_on
    cfg.close()


def test_config2_1024_one_block_messages_bit_exact(pkg):
    """BASELINE configs[1]: 1024 random 55-byte messages, one block each, every cell compared with the oracle."""
    import __graft_entry__ as ge
    S = ge.load_package_module("synthetic")
    w = S.WORKLOADS["cfg2"]
    blob, offs, lens = S.generate(w, 0, w.n_instances)
    instances = [[bytes(blob[int(o):int(o) + int(l)])] for o, l in zip(offs, lens)]
    _compare(pkg, dict(max_variable_byte_sizes=w.max_variable_byte_sizes), instances)


def test_config1_single_message(pkg):
    """BASELINE configs[0]: one 64-byte message, max input 128 bytes; plus the bench's own [0x01;56] with max 1024."""
    _compare(pkg, dict(max_variable_byte_sizes=(128,)), [[bytes(range(64))]])
    res, _ = _compare(pkg, dict(max_variable_byte_sizes=(1024,)), [[b"\x01" * 56]])
    assert bytes(res.digests[0]).hex() == "51e14a913680f24c85fe3b0e2e5b57f7202f117bb214f8ffdd4ea0f4e921fd52"


def test_config3_dynamic_lengths_with_edges(pkg):
    """BASELINE configs[2] shape (max 1088 = 17 blocks): edge lengths 0, 55, 56, 63, 64, 119, 120, max-9 and random ones."""
    rng = np.random.default_rng(3)
    lens = [0, 1, 55, 56, 63, 64, 119, 120, 1023, 1024, 1079] + [int(x) for x in rng.integers(0, 1025, size=13)]
    instances = [[bytes(rng.integers(0, 256, n, dtype=np.uint8))] for n in lens]
    res, _ = _compare(pkg, dict(max_variable_byte_sizes=(1088,)), instances)
    for inst, d in zip(instances, res.digests):
        assert hashlib.sha256(inst[0]).digest() == bytes(d)


def test_precomputed_prefix(pkg):
    """test_sha256_correct4 (lib.rs:587-611): 192 bytes with precomputed_input_len = 128, and a target_round == 0 corner."""
    rng = np.random.default_rng(9)
    kw = dict(max_variable_byte_sizes=(128, 128))
    instances = [[bytes(rng.integers(0, 256, 192, dtype=np.uint8)) for _ in range(2)] for _ in range(3)]
    pre = [[128, 128], [192, 128], [256, 192]]
    res, _ = _compare(pkg, kw, instances, pre)
    for inst, ds in zip(instances, res.digests.reshape(3, 2, 32)):
        assert [hashlib.sha256(m).digest() for m in inst] == [bytes(x) for x in ds]
    # everything inside the prefix: padded size == precomputed length, so the selected round is 0
    _compare(pkg, dict(max_variable_byte_sizes=(64,)), [[b"q" * 50]], [[64]])


@pytest.mark.parametrize("kw", [dict(max_variable_byte_sizes=(128,), max_rows=4099), dict(max_variable_byte_sizes=(64,), lookup_bits=8),
                                dict(max_variable_byte_sizes=(64,), lookup_bits=12), dict(max_variable_byte_sizes=(64,), limb_bits=4),
                                dict(max_variable_byte_sizes=(64,), limb_bits=2, spread_cols=3), dict(max_variable_byte_sizes=(128,), spread_cols=1),
                                dict(max_variable_byte_sizes=(128,), is_input_range_check=False), dict(max_variable_byte_sizes=(64, 192, 128))],
                         ids=lambda kw: "-".join(f"{k}={v}" for k, v in kw.items() if k != "max_variable_byte_sizes") or "multi")
def test_alternative_configurations(pkg, kw):
    rng = np.random.default_rng(11)
    D = len(kw["max_variable_byte_sizes"])
    instances = [[bytes(rng.integers(0, 256, int(rng.integers(0, m - 8)), dtype=np.uint8)) for m in kw["max_variable_byte_sizes"]] for _ in range(3)]
    assert all(len(i) == D for i in instances)
    _compare(pkg, kw, instances)


@pytest.mark.parametrize("kw", [dict(max_variable_byte_sizes=(64,), limb_bits=16), dict(max_variable_byte_sizes=(192, 128), limb_bits=16, spread_cols=1),
                                dict(max_variable_byte_sizes=(128,), limb_bits=16, spread_cols=3, lookup_bits=12)],
                         ids=["1block", "2digests-1col", "3cols-12bit"])
def test_sixteen_bit_spread_limbs(pkg, kw):
    """num_bits_lookup = 16 is legal in the reference (`16 % num_bits_lookup == 0`, spread.rs:37): one limb per `spread`, no shared-memory
    spread table -- the limb's spread is the 32-bit spread slot itself.  Cells bit-exact against the oracle, a corrupted spread cell is
    caught by the device-side spread-lookup check, and the two calls that keep table bins in shared memory refuse cleanly."""
    import torch
    rng = np.random.default_rng(16)
    instances = [[bytes(rng.integers(0, 256, int(rng.integers(0, m - 8)), dtype=np.uint8)) for m in kw["max_variable_byte_sizes"]] for _ in range(3)]
    res, _ = _compare(pkg, kw, instances)
    cfg = _engine(pkg, kw)
    res = cfg.digest_batch(instances)
    assert all(v == 0 for v in cfg.check_batch(res).values())
    res.spread[1, kw.get("spread_cols", 2), 5, 0] ^= 4      # spread half of column pair 0, row 5 of instance 1: no longer the spread of its dense cell
    torch.cuda.synchronize()
    bad = cfg.check_batch(res)
    assert bad["spread_lookups"] >= 1 and bad["gates"] == 0
    with pytest.raises(pkg.EngineError):
        cfg.lookup_multiplicities(res, (1 << 17) - 6)
    cfg.close()


@pytest.mark.parametrize("parts", [1, 6, 12, 24])
def test_job_granularity_does_not_change_the_cells(pkg, parts):
    """h2sha_config_t.block_parts only decides how a compression is cut into GPU jobs (>= 12: the latency setting, 8-instance jobs);
    cells, digests and checksums stay bit-exact, for one digest per call and for a batch with two digests per context."""
    rng = np.random.default_rng(parts)
    _compare(pkg, dict(max_variable_byte_sizes=(128,)), [[bytes(range(64))]], block_parts=parts)
    kw = dict(max_variable_byte_sizes=(192, 1088))
    instances = [[bytes(rng.integers(0, 256, int(rng.integers(0, m - 8)), dtype=np.uint8)) for m in kw["max_variable_byte_sizes"]] for _ in range(5)]
    _compare(pkg, kw, instances, block_parts=parts)


def test_wider_strides_leave_unassigned_cells_untouched(pkg):
    import torch
    kw = dict(max_variable_byte_sizes=(64,))
    cfg = _engine(pkg, kw, gate_col_rows=1 << 17, lookup_col_rows=4096, spread_rows=2100)
    lay = cfg.layout
    sentinel = -0x0123456789ABCDEF
    outs = cfg.alloc_outputs(2, zero=False)
    for t in outs:
        t.fill_(sentinel)
    res = cfg.digest_batch([[b"abc"], [b"abcd" * 13]], outputs=outs)
    olay = O.Layout(lay.n_gate_cols, lay.gate_col_rows, lay.n_lookup_cols, lay.lookup_col_rows, lay.spread_rows)
    ref = O.batch(_oracle_cfg(kw), olay, [[b"abc"], [b"abcd" * 13]], None, want_cells=True)
    g = _u64(res.gate)
    assert (g[:, :, : lay.n_gate_cells] == ref["gate"][:, :, : lay.n_gate_cells]).all()
    assert (res.gate[:, :, lay.n_gate_cells:] == sentinel).all(), "cells beyond the assigned rows must not be written"
    assert (res.lookup[:, :, lay.n_lookup_cells:] == sentinel).all()
    assert (res.spread[:, :, lay.n_spread_limbs // 2:] == sentinel).all()
    # zero-fill of just the never-assigned ranges
    lib = pkg.load_library()
    rc = lib.h2sha_zero_outputs(cfg._h, 2, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), 1, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert (_u64(res.gate) == ref["gate"]).all() and (_u64(res.lookup) == ref["lookup"]).all() and (_u64(res.spread) == ref["spread"]).all()
    cfg.close()


def test_reference_panics_are_errors(pkg):
    cfg = _engine(pkg, dict(max_variable_byte_sizes=(128,)))
    with pytest.raises(pkg.ReferencePanic):   # lib.rs:89
        cfg.digest_batch([[b"abc"]], [[32]])
    with pytest.raises(pkg.ReferencePanic):   # lib.rs:90
        cfg.digest_batch([[b"a" * 120]])
    cfg.digest_batch([[b"a" * 119]])
    with pytest.raises(pkg.EngineError):
        cfg.digest_batch([[b"a", b"b"]])
    cfg.close()


def test_config4_sample_properties_at_full_shape(pkg):
    """BASELINE configs[3] shape (256-byte messages, max 320, 5 blocks): 512 instances from the synthetic stream --
    all digests against hashlib, checksum-of-checksums against the oracle on a 32-instance sample, idempotence."""
    import __graft_entry__ as ge
    S = ge.load_package_module("synthetic")
    w = S.WORKLOADS["cfg4"]
    n = 512
    blob, offs, lens = S.generate(w, 4096, n)
    instances = [[bytes(blob[int(o):int(o) + int(l)])] for o, l in zip(offs, lens)]
    cfg = _engine(pkg, dict(max_variable_byte_sizes=w.max_variable_byte_sizes))
    lay = cfg.layout
    assert lay.n_gate_cols == 3 and lay.cells_per_instance == 406958
    outs = cfg.alloc_outputs(n)
    res = cfg.digest_batch(instances, outputs=outs)
    for inst, d in zip(instances, res.digests):
        assert hashlib.sha256(inst[0]).digest() == bytes(d)
    first = res.checksums.copy()
    res2 = cfg.digest_batch(instances, outputs=outs)  # same buffers, same inputs: nothing may change
    assert (res2.checksums == first).all()
    olay = O.Layout(lay.n_gate_cols, lay.gate_col_rows, lay.n_lookup_cols, lay.lookup_col_rows, lay.spread_rows)
    sample = list(range(0, n, 16))
    ref = O.batch(_oracle_cfg(dict(max_variable_byte_sizes=w.max_variable_byte_sizes)), olay, [instances[i] for i in sample], None,
                  want_cells=False, n_threads=NCPU)
    assert (first[sample] == ref["checksums"]).all()
    assert int(first[sample, 3].sum()) == int(ref["checksums"][:, 3].sum())
    cfg.close()


def test_config5_shape_33_blocks_18_columns(pkg):
    """BASELINE configs[4] shape (max 2112 = 33 blocks, 18 gate columns at k = 17): column wraps inside units, lookup
    column wrap, long digest traces.  Edge lengths 1, 2047, 2048, max-9 and a random one."""
    rng = np.random.default_rng(55)
    lens = [1, 2047, 2048, 2103, int(rng.integers(1, 2049))]
    instances = [[bytes(rng.integers(0, 256, n, dtype=np.uint8))] for n in lens]
    res, _ = _compare(pkg, dict(max_variable_byte_sizes=(2112,)), instances)
    for inst, d in zip(instances, res.digests):
        assert hashlib.sha256(inst[0]).digest() == bytes(d)


def test_checksum_only_device_inputs_and_input_reuse(pkg):
    """The other entry modes of h2sha_digest_batch: no cell buffers (checksums only), message bytes already on the
    device, and reuse_inputs (inputs of the previous call still resident in HBM)."""
    import torch
    kw = dict(max_variable_byte_sizes=(128,))
    cfg = _engine(pkg, kw)
    rng = np.random.default_rng(77)
    instances = [[bytes(rng.integers(0, 256, int(n), dtype=np.uint8))] for n in rng.integers(0, 120, size=64)]
    full = cfg.digest_batch(instances)
    only = cfg.digest_batch(instances, want_cells=False)
    assert (only.checksums == full.checksums).all() and (only.digests == full.digests).all()
    blob, offs, lens = pkg.pack_messages(instances)
    d_blob = torch.from_numpy(np.concatenate([blob, np.zeros(16, np.uint8)])).cuda()
    dig = np.zeros((64, 32), np.uint8); cks = np.zeros((64, 4), np.uint64)
    st = torch.cuda.current_stream().cuda_stream
    cfg.digest_batch_raw(64, d_blob.data_ptr(), True, int(blob.size), offs, lens, None, digests_host_ptr=dig.ctypes.data,
                         checksums_host_ptr=cks.ctypes.data, stream=st)
    torch.cuda.synchronize()
    assert (cks == full.checksums).all() and (dig == full.digests).all()
    dig[:] = 0; cks[:] = 0
    cfg.digest_batch_raw(64, 0, True, 0, offs, lens, None, digests_host_ptr=dig.ctypes.data, checksums_host_ptr=cks.ctypes.data, stream=st,
                         reuse_inputs=True)
    torch.cuda.synchronize()
    assert (cks == full.checksums).all() and (dig == full.digests).all()
    with pytest.raises(pkg.EngineError):   # more instances than are resident
        cfg.digest_batch_raw(65, 0, True, 0, offs, lens, None, reuse_inputs=True)
    cfg.close()


def test_random_configurations_cells_property(pkg):
    """Property test (seeded): random chip configurations and random messages, every cell against the oracle."""
    rng = np.random.default_rng(4242)
    for _ in range(6):
        kw = dict(max_variable_byte_sizes=tuple(int(64 * rng.integers(1, 4)) for _ in range(int(rng.integers(1, 3)))),
                  lookup_bits=int(rng.choice([8, 9, 10, 11, 12, 13, 14, 16, 17, 18, 19, 20])), limb_bits=int(rng.choice([2, 4, 8])), spread_cols=int(rng.integers(1, 4)),
                  is_input_range_check=bool(rng.integers(0, 2)), max_rows=int(rng.integers(3000, 200000)))
        instances = [[bytes(rng.integers(0, 256, int(rng.integers(0, m - 8)), dtype=np.uint8)) for m in kw["max_variable_byte_sizes"]]
                     for _ in range(3)]
        _compare(pkg, kw, instances)


def test_no_out_of_bounds_writes_guard_bands(pkg):
    """compute-sanitizer is closed on this pool, so bound the writes ourselves: every output buffer sits between two
    sentinel bands that must stay intact, for a multi-column configuration and a batch larger than the CTA count."""
    import torch
    kw = dict(max_variable_byte_sizes=(128,), max_rows=20011)
    cfg = _engine(pkg, kw)
    lay = cfg.layout
    n, band = 200, 4096
    rng = np.random.default_rng(99)
    instances = [[bytes(rng.integers(0, 256, int(rng.integers(0, 120)), dtype=np.uint8))] for _ in range(n)]
    blob, offs, lens = pkg.pack_messages(instances)
    sentinel = 0x5A5A5A5A5A5A5A5A
    bufs = []
    for nbytes in (lay.gate_bytes, lay.lookup_bytes, lay.spread_bytes):
        t = torch.full((band + n * nbytes // 8 + band,), sentinel, dtype=torch.int64, device="cuda")
        bufs.append(t)
    dig = np.zeros((n, 32), np.uint8); cks = np.zeros((n, 4), np.uint64)
    cfg.digest_batch_raw(n, blob.ctypes.data, False, int(blob.size), offs, lens, None, gate_ptr=bufs[0].data_ptr() + band * 8,
                         lookup_ptr=bufs[1].data_ptr() + band * 8, spread_ptr=bufs[2].data_ptr() + band * 8, digests_host_ptr=dig.ctypes.data,
                         checksums_host_ptr=cks.ctypes.data, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for t in bufs:
        assert (t[:band] == sentinel).all() and (t[-band:] == sentinel).all(), "write outside the output buffer"
    olay = O.Layout(lay.n_gate_cols, lay.gate_col_rows, lay.n_lookup_cols, lay.lookup_col_rows, lay.spread_rows)
    ref = O.batch(_oracle_cfg(kw), olay, instances, None, want_cells=False, n_threads=NCPU)
    assert (cks == ref["checksums"]).all() and (dig == ref["digests"]).all()
    # assigned cells differ from the sentinel; everything unassigned still holds it
    g = bufs[0][band:-band].view(n, lay.n_gate_cols, lay.gate_col_rows, 4)
    brk = list(cfg.breaks()) + [lay.n_gate_cells]
    for c in range(lay.n_gate_cols):
        used = int(brk[c + 1]) - int(brk[c])
        assert (g[:, c, used:] == sentinel).all(), "unassigned gate rows were written"
    cfg.close()


def test_export_instance_matches_halo2_witness_layout(pkg):
    """Prover hand-off (h2sha_export_instance): one instance's advice columns as zero-padded 2^k-row vectors in the
    reference's column order -- compared with the oracle's column-major output."""
    kw = dict(max_variable_byte_sizes=(128, 128))
    cfg = _engine(pkg, kw)
    lay = cfg.layout
    instances = [[b"abc", b""], [b"\x01" * 56, b"\x00\x00\x00"]]
    res = cfg.digest_batch(instances)
    cols = cfg.export_instance(res, 1, 1 << 17)
    assert cols.shape == (3 + 1 + 4, 1 << 17, 4)
    olay = O.Layout(lay.n_gate_cols, lay.gate_col_rows, lay.n_lookup_cols, lay.lookup_col_rows, lay.spread_rows)
    ref = O.batch(_oracle_cfg(kw), olay, instances, None, want_cells=True)
    assert (cols[0:3, : lay.gate_col_rows] == ref["gate"][1]).all() and not cols[0:3, lay.gate_col_rows:].any()
    assert (cols[3, : lay.lookup_col_rows] == ref["lookup"][1][0]).all() and not cols[3, lay.lookup_col_rows:].any()
    assert (cols[4:8, : lay.spread_rows] == ref["spread"][1]).all() and not cols[4:8, lay.spread_rows:].any()
    with pytest.raises(pkg.EngineError):
        cfg.export_instance(res, 0, 1024)   # fewer rows than a column has assigned
    cfg.close()


def test_long_digests_up_to_122_blocks(pkg):
    """maximum sizes: a 4096-byte digest (64 blocks, 35 gate columns) runs as one digest job; a 7808-byte digest (122 blocks,
    the largest the inverse table of is_zero covers) has its prologue/epilogue job cut into several job classes.  Beyond that
    the configuration is refused, never mis-generated."""
    rng = np.random.default_rng(77)
    _compare(pkg, dict(max_variable_byte_sizes=(4096,)), [[bytes(rng.integers(0, 256, int(n), dtype=np.uint8))] for n in (0, 4096 - 9, 2051)], threads=3)
    _compare(pkg, dict(max_variable_byte_sizes=(7808, 128)), [[bytes(rng.integers(0, 256, int(n), dtype=np.uint8)), b"abc"] for n in (7808 - 9, 6000)], threads=2)
    with pytest.raises(pkg.EngineError) as ei:
        pkg.Sha256DynamicConfig.configure([7872], device=0)
    assert ei.value.code == pkg.H2SHA_EINVAL
