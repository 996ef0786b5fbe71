"""Prover hand-off of whole batches: h2sha_export_batch (the cells, three strided copies) and the compact
hand-off (dictionary of distinct values + static cell -> entry map + host-side expander).

CPU part: the map and the expander against the oracle's witness (a dictionary rebuilt from the oracle's cells through the map must
expand to exactly those cells, and cells that share an entry must hold equal values in the oracle).
GPU part: expander(dictionary written by the kernel) == h2sha_export_batch == the batch buffers, bit for bit."""
import numpy as np
import pytest

from oracle import oracle as O

CONFIGS = [
    dict(sizes=(64,), msgs=[[b"a" * 55], [b""]]),
    dict(sizes=(128, 128), msgs=[[b"abc", b""], [bytes(range(100)), b"\x01" * 56]]),
    dict(sizes=(320,), msgs=[[bytes(range(200))]], max_rows=9973),
    dict(sizes=(192,), msgs=[[b"xyz" * 40]], lookup_bits=9, num_bits_lookup=4, num_advice_columns=3),
]


def _oracle_columns(kw, rows):
    cfg = O.OracleConfig(max_variable_byte_sizes=tuple(kw["sizes"]), max_rows=kw.get("max_rows", (1 << 17) - 9), lookup_bits=kw.get("lookup_bits", 16),
                         limb_bits=kw.get("num_bits_lookup", 8), spread_cols=kw.get("num_advice_columns", 2))
    reg = O.synthesize(cfg, kw["msgs"][0], record_shape=False)
    n_lk_cols = max(1, -(-len(reg.lookup_idx) // cfg.max_rows))
    lay = O.Layout(len(reg.breaks), rows, n_lk_cols, rows, rows)
    out = O.batch(cfg, lay, kw["msgs"], None, want_cells=True)
    return np.concatenate([out["gate"], out["lookup"], out["spread"]], axis=1)   # [n, columns, rows, 4]


def _engine_kw(kw):
    return {k: v for k, v in kw.items() if k not in ("sizes", "msgs")}


@pytest.mark.parametrize("kw", CONFIGS, ids=lambda k: "x".join(map(str, k["sizes"])))
def test_compact_map_and_expander_against_the_oracle(pkg, kw):
    eng = pkg.Sha256DynamicConfig.configure(list(kw["sizes"]), device=-1, **_engine_kw(kw))
    lay, ci = eng.layout, eng.compact_info()
    gmap, lmap, dmap, smap, consts = eng.compact_map()
    assert ci["cells_per_instance"] == lay.cells_per_instance and ci["dict_bytes_per_instance"] == 32 * ci["dict_cells_per_instance"]
    assert ci["dict_cells_per_instance"] < 0.4 * ci["cells_per_instance"]            # the point of the format: ~27 % of the cells
    rows = 1 << 14 if kw.get("max_rows", 1 << 17) < (1 << 14) else 1 << 17
    want = _oracle_columns(kw, rows)
    n = want.shape[0]
    brk = list(eng.breaks()) + [lay.n_gate_cells]
    max_rows = kw.get("max_rows", (1 << 17) - 9)
    nc = lay.n_spread_cols // 2
    # every cell's value in the oracle, keyed by its map entry
    dicts = np.zeros((n, ci["dict_cells_per_instance"], 4), np.uint64)
    seen = np.zeros((n, ci["dict_cells_per_instance"]), bool)

    def feed(i, m, vals):
        is_c = (m & 0x80000000) != 0
        assert (vals[is_c] == consts[m[is_c] & 0x7FFFFFFF]).all(), "a cell mapped to a constant holds another value in the oracle"
        e = m[~is_c]
        first = ~seen[i, e]
        # cells sharing a dictionary entry must agree
        dicts[i, e[first]] = vals[~is_c][first]
        seen[i, e] = True
        assert (dicts[i, e] == vals[~is_c]).all(), "two cells that share a dictionary entry differ in the oracle"

    for i in range(n):
        for c in range(lay.n_gate_cols):
            feed(i, gmap[brk[c]:brk[c + 1]], want[i, c, : brk[c + 1] - brk[c]])
        for c in range(lay.n_lookup_cols):
            lo, hi = min(lay.n_lookup_cells, c * max_rows), min(lay.n_lookup_cells, (c + 1) * max_rows)
            feed(i, lmap[lo:hi], want[i, lay.n_gate_cols + c, : hi - lo])
        nl = np.arange(lay.n_spread_limbs)
        base = lay.n_gate_cols + lay.n_lookup_cols
        feed(i, dmap, want[i, base + nl % nc, nl // nc])
        feed(i, smap, want[i, base + nc + nl % nc, nl // nc])
    assert seen.all(), "dictionary entries no cell uses"
    got = eng.expand_compact(dicts, n, rows, n_threads=3)
    assert got.shape == want.shape and (got == want).all()
    # single-threaded and without zero fill into a pre-zeroed buffer: same result
    out = np.zeros_like(want)
    eng.expand_compact(dicts, n, rows, out=out, n_threads=1, zero_fill=False)
    assert (out == want).all()
    with pytest.raises(pkg.EngineError):
        eng.expand_compact(dicts, n, 64)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CONFIGS, ids=lambda k: "x".join(map(str, k["sizes"])))
def test_compact_dictionary_from_the_kernel_expands_to_the_batch(pkg, kw):
    import torch
    eng = pkg.Sha256DynamicConfig.configure(list(kw["sizes"]), device=0, **_engine_kw(kw))
    lay, ci = eng.layout, eng.compact_info()
    rng = np.random.default_rng(3)
    msgs = kw["msgs"] + [[bytes(rng.integers(0, 256, int(rng.integers(0, s - 8)), dtype=np.uint8)) for s in kw["sizes"]] for _ in range(9)]
    n = len(msgs)
    rows = 1 << 14 if kw.get("max_rows", 1 << 17) < (1 << 14) else 1 << 17
    res = eng.digest_batch(msgs)
    full = eng.export_batch(res, 0, n, rows).numpy().view(np.uint64)
    # export_batch == the batch buffers (assigned prefix of every column) and zero elsewhere
    g = res.gate.cpu().numpy().view(np.uint64)
    r0 = min(rows, lay.gate_col_rows)
    assert (full[:, : lay.n_gate_cols, :r0] == g[:, :, :r0]).all() and not full[:, : lay.n_gate_cols, lay.gate_col_rows:].any()
    part = eng.export_batch(res, 3, 4, rows).numpy().view(np.uint64)
    assert (part == full[3:7]).all()
    # the kernel's dictionary, with and without the cells being written in the same launch
    blob, offs, lens = pkg.pack_messages(msgs)
    stream = torch.cuda.current_stream(0).cuda_stream
    dicts = []
    for with_cells in (True, False):
        d = torch.zeros((n, ci["dict_cells_per_instance"], 4), dtype=torch.int64, device="cuda:0")
        cks = torch.zeros((n, 4), dtype=torch.int64, device="cuda:0")
        outs = eng.alloc_outputs(n) if with_cells else (None, None, None)
        eng.digest_batch_raw(n, blob.ctypes.data, False, int(blob.size), offs, lens, None, gate_ptr=outs[0].data_ptr() if with_cells else 0,
                             lookup_ptr=outs[1].data_ptr() if with_cells else 0, spread_ptr=outs[2].data_ptr() if with_cells else 0,
                             checksums_dev_ptr=cks.data_ptr(), compact_dict_ptr=d.data_ptr(), stream=stream)
        torch.cuda.synchronize()
        assert (cks.cpu().numpy().view(np.uint64) == res.checksums).all(), "checksums change with the compact hand-off"
        if with_cells:
            assert torch.equal(outs[0], res.gate) and torch.equal(outs[1], res.lookup) and torch.equal(outs[2], res.spread)
        dicts.append(d.cpu().numpy().view(np.uint64))
    assert (dicts[0] == dicts[1]).all()
    got = eng.expand_compact(dicts[1], n, rows)
    assert (got == full).all(), f"{int((got != full).any(axis=-1).sum())} cells differ between expander(compact) and the full export"
    with pytest.raises(pkg.EngineError):   # one extra output per launch
        eng.digest_batch_raw(n, blob.ctypes.data, False, int(blob.size), offs, lens, None, compact_dict_ptr=1, lookup_mult_ptr=1, mult_usable_rows=1 << 17)
    eng.close()
