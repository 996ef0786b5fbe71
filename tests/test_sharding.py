"""N>1 host logic on CPU (gloo, world_size 2): instance ranges per rank, synthetic messages keyed by the GLOBAL instance
index, and the post-hot-path gather of digests + checksums.  The per-rank witness work is stood in for by the oracle
(tests may call it; the product path never does) -- what is under test is the sharding + gather plumbing."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    from oracle import oracle as O
    S = ge.load_package_module("synthetic")
    sh = ge.load_package_module("sharding")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = S.WORKLOADS["cfg2"]
    first, last = sh.shard_range(n_total, rank, world)
    blob, offs, lens = S.generate(w, first, last - first)
    cfg = O.OracleConfig(max_variable_byte_sizes=w.max_variable_byte_sizes)
    reg = O.synthesize(cfg, [bytes(blob[: int(lens[0])])], record_shape=False)
    out = O.batch_packed(cfg, reg.layout(), last - first, np.concatenate([blob, np.zeros(1, np.uint8)]), offs, lens,
                         np.zeros(last - first, np.uint32), n_threads=2)
    d = torch.from_numpy(out["digests"])
    c = torch.from_numpy(out["checksums"].view(np.int64))
    alld, allc, job = sh.gather_results(d, c, world)
    if rank == 0:
        np.save(os.path.join(out_dir, "digests.npy"), alld.numpy())
        np.save(os.path.join(out_dir, "cks.npy"), allc.numpy())
        with open(os.path.join(out_dir, "job.txt"), "w") as f:
            f.write(str(job))
    dist.destroy_process_group()


def test_shard_ranges_cover_everything(pkg):
    import __graft_entry__ as ge
    sh = ge.load_package_module("sharding")
    for n in (1, 7, 8, 1024, 65536, 65537):
        for world in (1, 2, 4, 8):
            rs = [sh.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_range(8, 2, 2)


def test_two_rank_gather_matches_single_process(pkg, tmp_path):
    import __graft_entry__ as ge
    from oracle import oracle as O
    S = ge.load_package_module("synthetic")
    n_total, world = 8, 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    digests = np.load(tmp_path / "digests.npy")
    cks = np.load(tmp_path / "cks.npy").view(np.uint64)
    job = int((tmp_path / "job.txt").read_text())
    w = S.WORKLOADS["cfg2"]
    blob, offs, lens = S.generate(w, 0, n_total)
    for i in range(n_total):
        assert bytes(digests[i]) == hashlib.sha256(bytes(blob[int(offs[i]):int(offs[i]) + int(lens[i])])).digest()
    cfg = O.OracleConfig(max_variable_byte_sizes=w.max_variable_byte_sizes)
    reg = O.synthesize(cfg, [bytes(blob[: int(lens[0])])], record_shape=False)
    ref = O.batch_packed(cfg, reg.layout(), n_total, np.concatenate([blob, np.zeros(1, np.uint8)]), offs, lens, np.zeros(n_total, np.uint32))
    assert (cks == ref["checksums"]).all()
    assert job == int(ref["checksums"][:, 3].astype(object).sum()) % (1 << 64)
