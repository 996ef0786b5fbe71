"""Soak test on one GPU: many seeded random chip configurations and message sets; every cell, digest and checksum against
the oracle, then the device-side MockProver pass and the lookup multiplicities of the same batch.
python tools/soak.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import lookup_prework as LP  # noqa: E402
from oracle import mock_prover as MP  # noqa: E402
from oracle import oracle as O  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pkg = ge.load_package()
rng = np.random.default_rng(seed)
t0 = time.perf_counter()
n_cfg = n_cells = 0
while time.perf_counter() - t0 < budget:
    max_blocks = int(os.environ.get("SOAK_MAX_BLOCKS", "6"))     # e.g. 122: long digests (split digest jobs), fewer configurations per minute
    sizes = tuple(int(64 * rng.integers(1, max_blocks + 1)) for _ in range(int(rng.integers(1, 4))))
    kw = dict(max_variable_byte_sizes=sizes, lookup_bits=int(rng.choice([8, 10, 12, 14, 16, 17, 18])), limb_bits=int(rng.choice([1, 2, 4, 8, 8, 16])),
              spread_cols=int(rng.integers(1, 5)), is_input_range_check=bool(rng.integers(0, 2)), max_rows=int(rng.integers(2500, 150000)))
    n_inst = int(rng.integers(1, 40 if max_blocks <= 8 else 4))
    use_pre = bool(rng.integers(0, 3) == 0)
    instances, pre = [], []
    for _ in range(n_inst):
        row, prow = [], []
        for m in sizes:
            p = int(64 * rng.integers(0, 3)) if use_pre else 0
            # padded length minus the prefix has to fit max: total length < p + m - 8, and at least the prefix itself
            lo = p if p else 0
            length = int(rng.integers(lo, p + m - 8))
            row.append(bytes(rng.integers(0, 256, length, dtype=np.uint8))); prow.append(p)
        instances.append(row); pre.append(prow)
    try:
        cfg = pkg.Sha256DynamicConfig.configure(list(sizes), max_rows=kw["max_rows"], lookup_bits=kw["lookup_bits"], num_bits_lookup=kw["limb_bits"],
                                                num_advice_columns=kw["spread_cols"], is_input_range_check=kw["is_input_range_check"], device=0)
    except pkg.EngineError as e:
        print("skipped (engine refuses the configuration):", kw, e, flush=True)
        continue
    lay = cfg.layout
    res = cfg.digest_batch(instances, pre if use_pre else None)
    olay = O.Layout(lay.n_gate_cols, lay.gate_col_rows, lay.n_lookup_cols, lay.lookup_col_rows, lay.spread_rows)
    ref = O.batch(O.OracleConfig(**kw), olay, instances, pre if use_pre else None, want_cells=True, n_threads=min(8, n_inst))
    ok = (res.digests == ref["digests"]).all() and (res.checksums == ref["checksums"]).all()
    for name in ("gate", "lookup", "spread"):
        ok = ok and bool((getattr(res, name).cpu().numpy().view(np.uint64) == ref[name]).all())
    dig = torch.from_numpy(res.digests).cuda()
    viol = cfg.check_batch(res, dig.data_ptr())
    ok = ok and sum(viol.values()) == 0
    if kw["limb_bits"] == 16:   # the lookup pre-work keeps spread-table bins in shared memory: it refuses 16-bit limbs (DESIGN.md limits)
        n_cfg += 1; n_cells += n_inst * lay.cells_per_instance
        if not ok:
            print("MISMATCH", kw, "instances", n_inst, "pre", use_pre, "violations", viol, flush=True)
            sys.exit(1)
        cfg.close()
        continue
    info = cfg.lookup_info()
    usable = max(info["min_usable_rows"], 1 << 10) + int(rng.integers(0, 1000))
    mult, bad = cfg.lookup_multiplicities(res, usable)
    ok = ok and bad == 0
    m = mult.cpu().numpy().astype(np.int64)
    nr, ns, rt, stt = info["n_range_lookups"], info["n_spread_lookups"], info["range_table_rows"], info["spread_table_rows"]
    ok = ok and bool((m[:, :nr * rt].reshape(n_inst, nr, rt).sum(-1) == usable).all()) and bool((m[:, nr * rt:].reshape(n_inst, ns, stt).sum(-1) == usable).all())
    # one permuted pair against the literal restatement (instance 0, a random lookup)
    l = int(rng.integers(0, nr + ns))
    theta = int.from_bytes(rng.bytes(32), "little") % O.P
    a, s = cfg.permute_lookup(mult[:1], l, usable, None if l < nr else O.int_to_mont(theta))
    canon = lambda t: MP.canon_all(t.cpu().numpy().view(np.uint64).reshape(-1, 4))
    if l < nr:
        used = max(0, min(kw["max_rows"], lay.n_lookup_cells - l * kw["max_rows"]))
        ia, it = LP.range_lookup_columns(canon(res.lookup[0, l, :used]), usable, kw["lookup_bits"])
    else:
        c = l - nr
        used = (lay.n_spread_limbs - c + ns - 1) // ns if lay.n_spread_limbs > c else 0
        ia, it = LP.spread_lookup_columns(canon(res.spread[0, c, :used]), canon(res.spread[0, ns + c, :used]), usable, kw["limb_bits"], theta)
    wa, ws = LP.permute_expression_pair(ia, it, usable)
    ok = ok and canon(a[0]) == wa and canon(s[0]) == ws
    n_cfg += 1; n_cells += n_inst * lay.cells_per_instance
    if not ok:
        print("MISMATCH", kw, "instances", n_inst, "pre", use_pre, "violations", viol, "bad", bad, "lookup", l, flush=True)
        sys.exit(1)
    cfg.close()
print(f"soak ok: {n_cfg} random configurations, {n_cells / 1e6:.1f} M cells bit-exact vs the oracle, device checks clean, seed {seed}, {time.perf_counter() - t0:.0f} s")
