"""Lookup-argument pre-work: table-row multiplicities and the permuted (A', S') pair of the chip's
range and spread lookups (reference src/spread.rs:53-62,165-194; src/lib.rs:409-418,469; halo2's
`permute_expression_pair`, restated in oracle/lookup_prework.py).

CPU part: the restatement itself (lookup-argument properties; a multiplicity-driven reconstruction -- the algorithm the
kernels use -- gives the same rows as the literal sort + BTreeMap walk) and the host side of the C-ABI.
GPU part: h2sha_lookup_multiplicities / h2sha_permute_lookup against the oracle on the witness of real batches."""
import ctypes as C

import numpy as np
import pytest

from oracle import lookup_prework as LP
from oracle import mock_prover as MP
from oracle import oracle as O

P = O.P


def _from_multiplicities(m, order, vals, usable_rows):
    """What k_permute_scan + k_permute_fill compute, in numpy: exclusive scans over the table rows in sorted order,
    then per row two binary searches."""
    n = len(order)
    mk = np.array([m[r] for r in order], dtype=np.int64)
    present = (mk > 0).astype(np.int64)
    t = np.ones(n, dtype=np.int64)
    t[[i for i, r in enumerate(order) if r == 0]] += usable_rows - n
    left = t - present
    start = np.concatenate([[0], np.cumsum(mk)[:-1]])
    dpre = np.concatenate([[0], np.cumsum(present)[:-1]])
    lpre = np.concatenate([[0], np.cumsum(left)[:-1]])
    assert mk.sum() == usable_rows
    n_rep = usable_rows - present.sum()
    rows = np.arange(usable_rows)
    k = np.searchsorted(start, rows, side="right") - 1
    a_p = [vals[i] for i in k]
    rank = rows - dpre[k] - 1
    j = n_rep - 1 - rank
    w = np.searchsorted(lpre, np.where(rows == start[k], 0, j), side="right") - 1
    s_p = [vals[k[r]] if r == start[k[r]] else vals[w[r]] for r in range(usable_rows)]
    return a_p, s_p


@pytest.mark.parametrize("seed", range(6))
def test_permute_expression_pair_properties_and_scan_formulation(seed):
    rng = np.random.default_rng(seed)
    bits = int(rng.integers(2, 7))
    n = 1 << bits
    usable = n + int(rng.integers(0, 3 * n))
    assigned = int(rng.integers(0, usable + 1))
    hot = rng.integers(0, n, size=max(1, n // 3))          # some table rows are never hit, some many times
    col = [int(hot[i]) for i in rng.integers(0, len(hot), size=assigned)]
    a, s = LP.range_lookup_columns(col, usable, bits)
    a_p, s_p = LP.permute_expression_pair(a, s, usable)
    LP.check_permuted(a, s, a_p, s_p)
    m = LP.multiplicities(col, n, usable)
    assert sum(m) == usable
    a_q, s_q = _from_multiplicities(m, list(range(n)), list(range(n)), usable)
    assert a_q == a_p and s_q == s_p


@pytest.mark.parametrize("seed", range(4))
def test_spread_lookup_compression_and_scan_formulation(seed):
    rng = np.random.default_rng(100 + seed)
    bits = 4
    n = 1 << bits
    theta = int.from_bytes(rng.bytes(32), "little") % P
    usable = 3 * n + 5
    dense = [int(x) for x in rng.integers(0, n, size=2 * n)]
    spread = [LP.spread_bits(d, bits) for d in dense]
    a, s = LP.spread_lookup_columns(dense, spread, usable, bits, theta)
    a_p, s_p = LP.permute_expression_pair(a, s, usable)
    LP.check_permuted(a, s, a_p, s_p)
    table = [LP.compress((i, LP.spread_bits(i, bits)), theta) for i in range(n)]
    order = sorted(range(n), key=lambda i: table[i])
    assert order[0] == 0                                    # (0, 0) compresses to 0: the padding row sorts first for every theta
    a_q, s_q = _from_multiplicities(LP.multiplicities(dense, n, usable), order, [table[i] for i in order], usable)
    assert a_q == a_p and s_q == s_p


def test_input_outside_the_table_is_rejected():
    with pytest.raises(ValueError):
        LP.permute_expression_pair([0, 1, 9], [0, 1, 2, 3], 3)


def test_lookup_info_and_argument_checks_without_a_device(pkg):
    cfg = pkg.Sha256DynamicConfig.configure([128, 128], device=-1)
    info = cfg.lookup_info()
    lay = cfg.layout
    assert info["n_range_lookups"] == lay.n_lookup_cols and info["n_spread_lookups"] == lay.n_spread_cols // 2
    assert info["range_table_rows"] == 1 << 16 and info["spread_table_rows"] == 256
    assert info["mult_words_per_instance"] == lay.n_lookup_cols * 65536 + 2 * 256
    assert info["min_usable_rows"] == max(65536, lay.n_lookup_cells, (lay.n_spread_limbs + 1) // 2)
    L = pkg.load_library()
    # no CPU path: a plan-only engine refuses the compute calls
    buf = (C.c_uint32 * 4)()
    assert L.h2sha_lookup_multiplicities(cfg._h, 1, buf, buf, 1 << 17, buf, None, None) == pkg.H2SHA_ECUDA
    assert L.h2sha_permute_lookup(cfg._h, 1, 0, buf, 1 << 17, None, buf, buf, None, None) == pkg.H2SHA_ECUDA
    cfg.close()


# ----------------------------------------------------------------------------------------------------------------
# GPU
# ----------------------------------------------------------------------------------------------------------------
def _canon(t):
    return MP.canon_all(t.cpu().numpy().view(np.uint64).reshape(-1, 4))


def _run(pkg, sizes, msgs, usable, **kw):
    cfg = pkg.Sha256DynamicConfig.configure(sizes, device=0, **kw)
    res = cfg.digest_batch(msgs)
    mult, bad = cfg.lookup_multiplicities(res, usable)
    return cfg, res, mult, bad


def _assigned_columns(cfg, res, inst):
    """canonical ints of the assigned prefix of every lookup / dense / spread column of one instance"""
    lay = cfg.layout
    max_rows = cfg._max_rows
    lk = []
    for c in range(lay.n_lookup_cols):
        used = max(0, min(max_rows, lay.n_lookup_cells - c * max_rows))
        lk.append(_canon(res.lookup[inst, c, :used]))
    cols = lay.n_spread_cols // 2
    dn, sp = [], []
    for c in range(cols):
        used = (lay.n_spread_limbs - c + cols - 1) // cols if lay.n_spread_limbs > c else 0
        dn.append(_canon(res.spread[inst, c, :used]))
        sp.append(_canon(res.spread[inst, cols + c, :used]))
    return lk, dn, sp


@pytest.mark.gpu
@pytest.mark.parametrize("max_rows", [(1 << 17) - 9, 5000])
def test_multiplicities_and_permuted_pairs_match_the_oracle(pkg, max_rows):
    """two 128-byte digests per instance (the reference's test shape); max_rows = 5000 forces three lookup columns"""
    rng = np.random.default_rng(5)
    sizes = [128, 128]
    msgs = [[bytes(rng.integers(0, 256, size=int(n), dtype=np.uint8)) for n in pair] for pair in ((0, 119), (55, 64), (1, 100))]
    usable = (1 << 17) - 6                                   # k = 17, 5 blinding factors
    cfg, res, mult, bad = _run(pkg, sizes, msgs, usable, max_rows=max_rows)
    cfg._max_rows = max_rows
    assert bad == 0
    info = cfg.lookup_info()
    n_r, n_s = info["n_range_lookups"], info["n_spread_lookups"]
    if max_rows == 5000:
        assert n_r == 3
    m = mult.cpu().numpy().astype(np.int64)
    theta = int.from_bytes(rng.bytes(32), "little") % P
    theta_m = O.int_to_mont(theta)
    perm = [cfg.permute_lookup(mult, l, usable, None if l < n_r else theta_m) for l in range(n_r + n_s)]
    for inst in range(len(msgs)):
        lk, dn, sp = _assigned_columns(cfg, res, inst)
        for l in range(n_r):
            want = LP.multiplicities(lk[l], 1 << 16, usable)
            assert (m[inst, l * 65536:(l + 1) * 65536] == np.array(want)).all(), f"range multiplicities, instance {inst}, column {l}"
            a, s = LP.range_lookup_columns(lk[l], usable, 16)
            a_p, s_p = LP.permute_expression_pair(a, s, usable)
            assert _canon(perm[l][0][inst]) == a_p, "A' (range)"
            assert _canon(perm[l][1][inst]) == s_p, "S' (range)"
        for c in range(n_s):
            want = LP.multiplicities(dn[c], 256, usable)
            off = n_r * 65536 + c * 256
            assert (m[inst, off:off + 256] == np.array(want)).all(), f"spread multiplicities, instance {inst}, pair {c}"
            a, s = LP.spread_lookup_columns(dn[c], sp[c], usable, 8, theta)
            a_p, s_p = LP.permute_expression_pair(a, s, usable)
            got_a, got_s = _canon(perm[n_r + c][0][inst]), _canon(perm[n_r + c][1][inst])
            assert got_a == a_p, "A' (spread)"
            assert got_s == s_p, "S' (spread)"
            LP.check_permuted(a, s, got_a, got_s)
    cfg.close()


@pytest.mark.gpu
def test_corrupted_witness_is_counted_not_binned(pkg):
    """a lookup cell >= 2^16 and a (dense, spread) pair that is no table row are reported; the multiplicities then no
    longer cover the usable rows and the permutation refuses the instance"""
    import torch
    rng = np.random.default_rng(6)
    usable = (1 << 17) - 6
    msgs = [[bytes(rng.integers(0, 256, size=40, dtype=np.uint8))] for _ in range(2)]
    cfg = pkg.Sha256DynamicConfig.configure([64], device=0)
    res = cfg.digest_batch(msgs)
    res.lookup[1, 0, 7] = torch.from_numpy(O.int_to_mont(1 << 16).view(np.int64)).to(res.lookup.device)
    res.spread[1, 2, 3] = torch.from_numpy(O.int_to_mont(2).view(np.int64)).to(res.spread.device)   # spread half: 2 is not a spread value
    mult, bad = cfg.lookup_multiplicities(res, usable)
    assert bad == 2
    a, s = cfg.permute_lookup(mult[:1], 0, usable)           # instance 0 is intact
    assert a.shape == (1, usable, 4)
    with pytest.raises(pkg.EngineError):
        cfg.permute_lookup(mult, 0, usable)
    cfg.close()


@pytest.mark.gpu
def test_permuted_pairs_at_batch_size_hold_the_lookup_rules(pkg):
    """size-independent check on a larger batch (64 instances x 17 blocks): on the device, for every row A'[i] == S'[i]
    or A'[i] == A'[i-1]; A' is sorted and has the multiplicities' histogram; S' holds every table row"""
    import torch
    rng = np.random.default_rng(8)
    n = 64
    usable = (1 << 17) - 6
    msgs = [[bytes(rng.integers(0, 256, size=int(rng.integers(0, 1025)), dtype=np.uint8))] for _ in range(n)]
    cfg = pkg.Sha256DynamicConfig.configure([1088], device=0)
    res = cfg.digest_batch(msgs)
    mult, bad = cfg.lookup_multiplicities(res, usable)
    assert bad == 0
    info = cfg.lookup_info()
    assert (mult.to(torch.int64).view(n, -1)[:, :65536 * info["n_range_lookups"]].view(n, info["n_range_lookups"], 65536).sum(-1) == usable).all()
    theta_m = O.int_to_mont(int.from_bytes(rng.bytes(32), "little") % P)
    import os
    for l in range(info["n_range_lookups"] + info["n_spread_lookups"]):
        os.environ["H2SHA_TUNE"] = "lkchunk=24"          # three chunks of instances share the scan workspace
        try:
            a, s = cfg.permute_lookup(mult, l, usable, None if l < info["n_range_lookups"] else theta_m)
        finally:
            del os.environ["H2SHA_TUNE"]
        a1, s1 = cfg.permute_lookup(mult[40:41], l, usable, None if l < info["n_range_lookups"] else theta_m)
        assert bool((a1[0] == a[40]).all()) and bool((s1[0] == s[40]).all()), "chunked and single-instance results differ"
        same = (a == s).all(-1)
        prev = torch.zeros_like(same)
        prev[:, 1:] = (a[:, 1:] == a[:, :-1]).all(-1)
        assert bool((same | prev).all()) and bool(same[:, 0].all())
        if l < info["n_range_lookups"]:
            # S' is a permutation of the padded range table: every value 1..2^16-1 once (checked through a checksum of the
            # low limb over distinct rows), and A' runs follow the multiplicities
            runs = (~prev).sum(-1)
            assert (runs == (mult[:, l * 65536:(l + 1) * 65536] > 0).sum(-1)).all()
            tab = cfg.mont_from_u64(np.arange(65536, dtype=np.uint64), path32=True)
            want = int(tab[:, 0].astype(object).sum() + (usable - 65536) * int(tab[0, 0])) & ((1 << 64) - 1)
            got = s[..., 0].cpu().numpy().view(np.uint64)
            for i in range(0, n, 16):
                assert int(got[i].astype(object).sum()) & ((1 << 64) - 1) == want
    cfg.close()


@pytest.mark.gpu
def test_prework_writes_stay_inside_their_buffers(pkg):
    """guard bands around the multiplicity and the permuted buffers stay untouched (compute-sanitizer is closed on this pool)"""
    import torch
    rng = np.random.default_rng(9)
    usable = (1 << 16) + 777                                 # not a multiple of the 256-row tiles
    n = 5
    msgs = [[bytes(rng.integers(0, 256, size=int(rng.integers(0, 56)), dtype=np.uint8))] for _ in range(n)]
    cfg = pkg.Sha256DynamicConfig.configure([64], device=0)
    res = cfg.digest_batch(msgs)
    info = cfg.lookup_info()
    L = pkg.load_library()
    dev = res.gate.device
    G = 4096
    words = n * info["mult_words_per_instance"]
    mbuf = torch.full((words + 2 * G,), 0x5A5A5A5A, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(0).cuda_stream
    assert L.h2sha_lookup_multiplicities(cfg._h, n, res.lookup.data_ptr(), res.spread.data_ptr(), usable, mbuf[G:].data_ptr(), None, st) == 0
    cells = n * usable * 4
    theta = O.int_to_mont(12345678901234567890 % P)
    for l in range(info["n_range_lookups"] + info["n_spread_lookups"]):
        a = torch.full((cells + 2 * G,), 0x6B6B6B6B6B6B6B6B, dtype=torch.int64, device=dev)
        s = torch.full((cells + 2 * G,), 0x6B6B6B6B6B6B6B6B, dtype=torch.int64, device=dev)
        th = theta.ctypes.data if l >= info["n_range_lookups"] else None
        assert L.h2sha_permute_lookup(cfg._h, n, l, mbuf[G:].data_ptr(), usable, th, a[G:].data_ptr(), s[G:].data_ptr(), None, st) == 0
        torch.cuda.synchronize()
        for t in (a, s):
            assert bool((t[:G] == 0x6B6B6B6B6B6B6B6B).all()) and bool((t[G + cells:] == 0x6B6B6B6B6B6B6B6B).all()), "guard band overwritten"
            assert not bool((t[G:G + cells].view(-1, 4) == 0x6B6B6B6B6B6B6B6B).all(-1).any()), "a row was not written"
    assert bool((mbuf[:G] == 0x5A5A5A5A).all()) and bool((mbuf[G + words:] == 0x5A5A5A5A).all()), "multiplicity guard band overwritten"
    assert bool((mbuf[G:G + words].view(n, -1).to(torch.int64).sum(-1) == usable * (info["n_range_lookups"] + info["n_spread_lookups"])).all())
    cfg.close()


@pytest.mark.gpu
def test_multiplicities_follow_widened_column_strides(pkg):
    """2^k-row column strides (the layout a prover holds): the histogram reads the assigned prefix of each column, wherever the stride puts it"""
    rng = np.random.default_rng(10)
    usable = (1 << 17) - 6
    msgs = [[bytes(rng.integers(0, 256, size=int(n), dtype=np.uint8))] for n in (0, 55, 119)]
    tight = pkg.Sha256DynamicConfig.configure([128], device=0)
    wide = pkg.Sha256DynamicConfig.configure([128], device=0, gate_col_rows=1 << 17, lookup_col_rows=1 << 17, spread_rows=1 << 17)
    m_t, bad_t = tight.lookup_multiplicities(tight.digest_batch(msgs), usable)
    res_w = wide.digest_batch(msgs)
    m_w, bad_w = wide.lookup_multiplicities(res_w, usable)
    assert bad_t == 0 and bad_w == 0 and bool((m_t == m_w).all())
    a_t, s_t = tight.permute_lookup(m_t, 0, usable)
    a_w, s_w = wide.permute_lookup(m_w, 0, usable)
    assert bool((a_t == a_w).all()) and bool((s_t == s_w).all())
    tight.close(); wide.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [
    dict(sizes=[64], n=48),                                                        # config 2 shape
    dict(sizes=[128, 128], n=5),                                                   # the reference's TestCircuit: two digests per region
    dict(sizes=[320], n=7, max_rows=9973),                                         # lookup column wraps: 2 lookup advice columns, wraps inside chunks
    dict(sizes=[192], n=3, lookup_bits=9, num_bits_lookup=4, num_advice_columns=3),  # other table sizes, 3 spread column pairs (not a power of two)
    dict(sizes=[64], n=4, lookup_bits=20, num_bits_lookup=2, num_advice_columns=1),
], ids=lambda k: "x".join(map(str, k["sizes"])) + "_" + "_".join(f"{a}{b}" for a, b in k.items() if a not in ("sizes", "n")))
def test_multiplicities_counted_while_the_cells_are_written(pkg, kw):
    """h2sha_batch_t.lookup_mult_dev: the table-row multiplicities come out of the expansion kernel itself and are bit-identical to
    the second pass over the finished witness (h2sha_lookup_multiplicities), which the tests above pin against the oracle."""
    import torch
    kw = dict(kw)
    sizes, n = kw.pop("sizes"), kw.pop("n")
    cfg = pkg.Sha256DynamicConfig.configure(sizes, device=0, **kw)
    info = cfg.lookup_info()
    usable = info["min_usable_rows"] + 11
    rng = np.random.default_rng(9)
    msgs = [[bytes(rng.integers(0, 256, int(rng.integers(0, s - 8)), dtype=np.uint8)) for s in sizes] for _ in range(n)]
    blob, offs, lens = pkg.pack_messages(msgs)
    gate, lookup, spread = cfg.alloc_outputs(n)
    dev = gate.device
    fused = torch.full((n, info["mult_words_per_instance"]), 123456, dtype=torch.int32, device=dev)   # must be overwritten, not accumulated into
    bad = torch.full((1,), 77, dtype=torch.int32, device=dev)
    try:
        cfg.digest_batch_raw(n, blob.ctypes.data if blob.size else 0, False, int(blob.size), offs, lens, None, gate_ptr=gate.data_ptr(),
                             lookup_ptr=lookup.data_ptr(), spread_ptr=spread.data_ptr(), lookup_mult_ptr=fused.data_ptr(), mult_usable_rows=usable,
                             stream=torch.cuda.current_stream(0).cuda_stream)
    except pkg.EngineError as ex:
        # refused, never mis-counted: configurations whose templates leave no shared memory for the raw-value scratch
        assert "no shared memory for the fused multiplicity count" in str(ex) and kw.get("num_bits_lookup", 8) < 8
        cfg.close()
        return
    for _ in range(2):   # twice: every call rewrites all bins
        cfg.digest_batch_raw(n, blob.ctypes.data if blob.size else 0, False, int(blob.size), offs, lens, None, gate_ptr=gate.data_ptr(),
                             lookup_ptr=lookup.data_ptr(), spread_ptr=spread.data_ptr(), lookup_mult_ptr=fused.data_ptr(), mult_usable_rows=usable,
                             mult_bad_ptr=bad.data_ptr(), stream=torch.cuda.current_stream(0).cuda_stream)
    torch.cuda.synchronize()
    second_pass, bad2 = cfg.lookup_multiplicities(pkg.BatchResult(None, None, gate, lookup, spread), usable)
    assert int(bad.item()) == 0 and bad2 == 0
    assert torch.equal(fused, second_pass), f"{int((fused != second_pass).sum())} bins differ"
    assert int(fused.sum(dim=1)[0]) == usable * (info["n_range_lookups"] + info["n_spread_lookups"])
    # argument checks
    L = pkg.load_library()
    with pytest.raises(pkg.EngineError):
        cfg.digest_batch_raw(n, blob.ctypes.data if blob.size else 0, False, int(blob.size), offs, lens, None, gate_ptr=gate.data_ptr(),
                             lookup_mult_ptr=fused.data_ptr(), mult_usable_rows=usable)
    with pytest.raises(pkg.EngineError):
        cfg.digest_batch_raw(n, blob.ctypes.data if blob.size else 0, False, int(blob.size), offs, lens, None, gate_ptr=gate.data_ptr(),
                             lookup_ptr=lookup.data_ptr(), spread_ptr=spread.data_ptr(), lookup_mult_ptr=fused.data_ptr(), mult_usable_rows=10)
    cfg.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(sizes=[64], n=5), dict(sizes=[128, 64], n=3, max_rows=4099), dict(sizes=[192], n=3, lookup_bits=9, num_bits_lookup=4, num_advice_columns=3)],
                         ids=["cfg2", "2digests-wraps", "9bit-3cols"])
def test_permuted_pairs_from_the_raw_value_lists(pkg, kw):
    """h2sha_batch_t.keep_lookup_raw + h2sha_permute_lookup_from_raw: no dense multiplicity array is written at all; the permuted pair of
    every lookup (range lookups and, with a theta, spread lookups), for the whole batch and for a slice of it, is bit-identical to
    h2sha_permute_lookup on the second-pass multiplicities (which the tests above pin against the oracle)."""
    import torch
    kw = dict(kw)
    sizes, n = kw.pop("sizes"), kw.pop("n")
    cfg = pkg.Sha256DynamicConfig.configure(sizes, device=0, **kw)
    info = cfg.lookup_info()
    usable = info["min_usable_rows"] + 5
    rng = np.random.default_rng(23)
    msgs = [[bytes(rng.integers(0, 256, int(rng.integers(0, s - 8)), dtype=np.uint8)) for s in sizes] for _ in range(n)]
    blob, offs, lens = pkg.pack_messages(msgs)
    gate, lookup, spread = cfg.alloc_outputs(n)
    with pytest.raises(pkg.EngineError):   # nothing generated with keep_lookup_raw yet
        cfg.permute_lookup_from_raw(0, n, 0, usable)
    try:
        cfg.digest_batch_raw(n, blob.ctypes.data if blob.size else 0, False, int(blob.size), offs, lens, None, gate_ptr=gate.data_ptr(),
                             lookup_ptr=lookup.data_ptr(), spread_ptr=spread.data_ptr(), keep_lookup_raw=True, stream=torch.cuda.current_stream(0).cuda_stream)
    except pkg.EngineError as ex:
        assert "no shared memory for the fused multiplicity count" in str(ex) and kw.get("num_bits_lookup", 8) < 8
        cfg.close()
        return
    torch.cuda.synchronize()
    mult, bad = cfg.lookup_multiplicities(pkg.BatchResult(None, None, gate, lookup, spread), usable)
    assert bad == 0
    theta = O.int_to_mont(int.from_bytes(rng.bytes(32), "little") % O.P)
    for l in range(info["n_range_lookups"] + info["n_spread_lookups"]):
        th = None if l < info["n_range_lookups"] else theta
        want_a, want_s = cfg.permute_lookup(mult, l, usable, th)
        got_a, got_s = cfg.permute_lookup_from_raw(0, n, l, usable, th)
        assert torch.equal(got_a, want_a) and torch.equal(got_s, want_s), f"lookup {l}"
        got_a, got_s = cfg.permute_lookup_from_raw(1, n - 1, l, usable, th)      # a slice of the batch
        assert torch.equal(got_a, want_a[1:]) and torch.equal(got_s, want_s[1:]), f"lookup {l}, instances 1.."
    with pytest.raises(pkg.EngineError):
        cfg.permute_lookup_from_raw(1, n, 0, usable)                               # beyond the batch
    cfg.close()
