"""Device-side MockProver-style check of a whole batch (h2sha_check_batch): all five violation counters are zero on the
engine's own output -- the acceptance criterion of the reference's tests (MockProver verify == Ok, src/lib.rs:525-526) --
and each class of corruption is caught by the class of constraint that covers it."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O


def test_plan_only_engine_refuses(pkg):
    cfg = pkg.Sha256DynamicConfig.configure([64], device=-1)
    buf = (C.c_uint64 * 8)()
    assert pkg.load_library().h2sha_check_batch(cfg._h, 1, buf, buf, buf, None, buf, None) == pkg.H2SHA_ECUDA
    cfg.close()


def _mont(v):
    import torch
    return torch.from_numpy(O.int_to_mont(v).view(np.int64))


@pytest.mark.gpu
def test_clean_batch_has_no_violations_and_corruptions_are_caught(pkg):
    import torch
    rng = np.random.default_rng(11)
    sizes = [128, 128]
    lens = [(0, 119), (55, 64), (1, 100), (56, 63), (119, 3)]
    msgs = [[bytes(rng.integers(0, 256, size=int(n), dtype=np.uint8)) for n in pair] for pair in lens]
    cfg = pkg.Sha256DynamicConfig.configure(sizes, device=0)
    res = cfg.digest_batch(msgs)
    dig = torch.from_numpy(res.digests).cuda()
    clean = cfg.check_batch(res, dig.data_ptr())
    assert clean == dict(gates=0, copies=0, range_lookups=0, spread_lookups=0, digest_bytes=0), clean

    shp = pkg.Sha256DynamicConfig.configure(sizes, device=-1, build_shape=True)
    sh, brk, lay = shp.shape(), shp.breaks(), shp.layout
    on = np.nonzero(sh.selectors)[0]

    def pos(idx):
        c = int(np.searchsorted(brk, idx, side="right") - 1)
        return c, int(idx - brk[c])

    # (a) the output cell of a gate in instance 2: the gate fails (and whatever copies that cell)
    g = int(on[len(on) // 2])
    c, r = pos(g + 3)
    saved = res.gate[2, c, r].clone()
    res.gate[2, c, r] = _mont(123456789).cuda()
    v = cfg.check_batch(res, dig.data_ptr())
    assert v["gates"] >= 1 and v["range_lookups"] == 0 and v["spread_lookups"] == 0 and v["digest_bytes"] == 0, v
    res.gate[2, c, r] = saved
    # (b) a constant cell (copy-constrained to the fixed column): copies + the gate it sits in
    fixed_copies = [cp for cp in sh.copies.tolist() if cp[2] == 1 or cp[0] == 1]
    cp = fixed_copies[len(fixed_copies) // 3]
    gidx = cp[1] if cp[0] == 0 else cp[3]
    c, r = pos(gidx)
    saved = res.gate[0, c, r].clone()
    res.gate[0, c, r] = _mont(77).cuda() if O.mont_to_int(saved.cpu().numpy().view(np.uint64)) != 77 else _mont(78).cuda()
    v = cfg.check_batch(res, dig.data_ptr())
    assert v["copies"] >= 1, v
    res.gate[0, c, r] = saved
    # (c) a lookup-column cell >= 2^16: the range lookup and its link to the gate cell
    saved = res.lookup[4, 0, 9].clone()
    res.lookup[4, 0, 9] = _mont((1 << 16) + 5).cuda()
    v = cfg.check_batch(res, dig.data_ptr())
    assert v["range_lookups"] == 1 and v["copies"] >= 1 and v["gates"] == 0, v
    res.lookup[4, 0, 9] = saved
    # (d) the spread half of a table row
    ncol = lay.n_spread_cols // 2
    saved = res.spread[1, ncol, 5].clone()
    res.spread[1, ncol, 5] = _mont(2).cuda()
    v = cfg.check_batch(res, dig.data_ptr())
    assert v["spread_lookups"] == 1 and v["copies"] >= 1 and v["gates"] == 0, v
    res.spread[1, ncol, 5] = saved
    # (e) a digest byte
    dig2 = dig.clone()
    dig2[3, 7] ^= 1
    v = cfg.check_batch(res, dig2.data_ptr())
    assert v == dict(gates=0, copies=0, range_lookups=0, spread_lookups=0, digest_bytes=1), v
    assert cfg.check_batch(res, dig.data_ptr()) == clean
    cfg.close(); shp.close()


@pytest.mark.gpu
def test_every_instance_of_config2_and_a_dynamic_batch_passes(pkg):
    """size-independent property at BASELINE size: all 1024 instances of config 2, and 96 dynamic-length 17-block instances
    (config 3 shape, 18 gate columns ... 2 lookup-relevant wraps), satisfy every constraint"""
    import torch
    import __graft_entry__ as ge
    S = ge.load_package_module("synthetic")
    for name, n in (("cfg2", 1024), ("cfg3", 96)):
        w = S.WORKLOADS[name]
        cfg = pkg.Sha256DynamicConfig.configure(list(w.max_variable_byte_sizes), device=0)
        blob, offs, lens = S.generate(w, 0, n)
        msgs = [[bytes(blob[int(o):int(o) + int(l)])] for o, l in zip(offs, lens)]
        res = cfg.digest_batch(msgs)
        dig = torch.from_numpy(res.digests).cuda()
        v = cfg.check_batch(res, dig.data_ptr())
        assert sum(v.values()) == 0, (name, v)
        cfg.close()


@pytest.mark.gpu
def test_check_follows_widened_column_strides(pkg):
    """the checker's cell positions come from the engine's own strides (2^k-row columns instead of tight ones)"""
    import torch
    rng = np.random.default_rng(12)
    cfg = pkg.Sha256DynamicConfig.configure([128], device=0, gate_col_rows=1 << 17, lookup_col_rows=1 << 17, spread_rows=1 << 17)
    msgs = [[bytes(rng.integers(0, 256, size=int(n), dtype=np.uint8))] for n in (3, 119, 64)]
    res = cfg.digest_batch(msgs)
    assert res.gate.shape[2] == 1 << 17
    dig = torch.from_numpy(res.digests).cuda()
    assert sum(cfg.check_batch(res, dig.data_ptr()).values()) == 0
    res.gate[1, 0, 1000] = _mont(99).cuda()
    assert sum(cfg.check_batch(res, dig.data_ptr()).values()) >= 1
    cfg.close()
