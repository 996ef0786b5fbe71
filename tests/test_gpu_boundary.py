"""Behaviour of the drop-in boundary itself (needs a B200): h2sha_digest_batch only enqueues, back-to-back calls with host
buffers deliver the same results as isolated ones, two engines of different configurations coexist, stale libraries are refused."""
import hashlib
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _msgs(n, length, seed):
    rng = np.random.default_rng(seed)
    return [bytes(rng.integers(0, 256, length, dtype=np.uint8)) for _ in range(n)]


def test_digest_batch_only_enqueues(pkg):
    """include/h2sha_b200.h: "the call only enqueues".  Eight calls with host buffers are issued back to back; the host must be
    back before the first batch's kernels have finished (round 1 waited for the device inside every call)."""
    import torch
    cfg = pkg.Sha256DynamicConfig.configure([64], device=0)
    n = 4096                                   # ~2 ms of expansion per call
    blob, offs, lens = pkg.pack_messages([[m] for m in _msgs(n, 55, 3)])
    h_blob = torch.from_numpy(blob).pin_memory()
    outs = cfg.alloc_outputs(n)
    hd = [torch.zeros((n, 32), dtype=torch.uint8).pin_memory() for _ in range(8)]
    hc = [torch.zeros((n, 4), dtype=torch.int64).pin_memory() for _ in range(8)]
    stream = torch.cuda.current_stream(0)

    def call(k):
        cfg.digest_batch_raw(n, h_blob.data_ptr(), False, int(blob.size), offs, lens, None, gate_ptr=outs[0].data_ptr(), lookup_ptr=outs[1].data_ptr(),
                             spread_ptr=outs[2].data_ptr(), digests_host_ptr=hd[k].data_ptr(), checksums_host_ptr=hc[k].data_ptr(), stream=stream.cuda_stream)

    call(0); call(1); call(2)                 # workspaces, pinned ring and events exist from here on
    torch.cuda.synchronize()
    first_done = torch.cuda.Event()
    t0 = time.perf_counter()
    for k in range(8):
        call(k)
        if k == 0:
            first_done.record(stream)
    host_s = time.perf_counter() - t0
    still_running = not first_done.query()
    torch.cuda.synchronize()
    total_s = time.perf_counter() - t0
    assert still_running, f"the host took {host_s * 1e3:.2f} ms for 8 calls and the first batch had already finished: the call blocks"
    assert host_s < 0.5 * total_s, f"host {host_s * 1e3:.2f} ms of {total_s * 1e3:.2f} ms: the calls do not run ahead of the device"
    # every call delivered the same (correct) results
    want = np.array([np.frombuffer(hashlib.sha256(bytes(blob[int(o):int(o) + int(l)])).digest(), np.uint8) for o, l in zip(offs[:64], lens[:64])])
    for k in range(8):
        assert (hd[k].numpy()[:64] == want).all()
        assert (hc[k].numpy() == hc[0].numpy()).all() and (hd[k].numpy() == hd[0].numpy()).all()
    cfg.close()


def test_pipelined_host_calls_match_isolated_calls(pkg):
    """Different batches issued back to back (the trace kernel of batch i+1 overlaps the expansion of batch i on the engine's copy
    stream): digests, checksums and cells equal those of the same batches run one at a time."""
    import torch
    cfg = pkg.Sha256DynamicConfig.configure([128], device=0)
    n = 96
    batches = [[[m] for m in _msgs(n, length, 10 + k)] for k, length in enumerate([0, 55, 56, 64, 119, 100, 1, 63])]
    iso = [cfg.digest_batch(b) for b in batches]
    iso_ck = [r.checksums.copy() for r in iso]
    iso_dg = [r.digests.copy() for r in iso]
    iso_gate = [r.gate.clone() for r in (iso[2], iso[7])]
    del iso
    stream = torch.cuda.current_stream(0)
    outs = [cfg.alloc_outputs(n) for _ in range(len(batches))]
    hd = [torch.zeros((n, 32), dtype=torch.uint8).pin_memory() for _ in batches]
    hc = [torch.zeros((n, 4), dtype=torch.int64).pin_memory() for _ in batches]
    keep = []
    for k, b in enumerate(batches):
        blob, offs, lens = pkg.pack_messages(b)
        keep.append(blob)
        cfg.digest_batch_raw(n, blob.ctypes.data if blob.size else 0, False, int(blob.size), offs, lens, None, gate_ptr=outs[k][0].data_ptr(),
                             lookup_ptr=outs[k][1].data_ptr(), spread_ptr=outs[k][2].data_ptr(), digests_host_ptr=hd[k].data_ptr(),
                             checksums_host_ptr=hc[k].data_ptr(), stream=stream.cuda_stream)
        blob[:] = 0xEE                             # host arrays are consumed before the call returns
        offs[:] = 0; lens[:] = 0
    torch.cuda.synchronize()
    for k in range(len(batches)):
        assert (hd[k].numpy() == iso_dg[k]).all(), f"batch {k}: digests differ"
        assert (hc[k].numpy().view(np.uint64) == iso_ck[k]).all(), f"batch {k}: checksums differ"
    assert torch.equal(outs[2][0], iso_gate[0]) and torch.equal(outs[7][0], iso_gate[1])
    cfg.close()


def test_reuse_inputs_after_host_call_and_on_another_stream(pkg):
    import torch
    cfg = pkg.Sha256DynamicConfig.configure([64], device=0)
    n = 64
    b = [[m] for m in _msgs(n, 40, 77)]
    r0 = cfg.digest_batch(b)
    blob, offs, lens = pkg.pack_messages(b)
    s2 = torch.cuda.Stream(0)
    d_cks = torch.zeros((n, 4), dtype=torch.int64, device="cuda:0")
    g, l, s = cfg.alloc_outputs(n)
    cfg.digest_batch_raw(n, 0, False, 0, offs, lens, None, gate_ptr=g.data_ptr(), lookup_ptr=l.data_ptr(), spread_ptr=s.data_ptr(),
                         checksums_dev_ptr=d_cks.data_ptr(), stream=s2.cuda_stream, reuse_inputs=True)
    s2.synchronize()
    assert (d_cks.cpu().numpy().view(np.uint64) == r0.checksums).all()
    assert torch.equal(g, r0.gate)
    cfg.close()


def test_large_then_small_configuration_then_large_again(pkg):
    """the dynamic shared-memory attribute belongs to the kernel function, not to an engine.  A second engine with a
    smaller plan must not make the first one's launches fail."""
    big = pkg.Sha256DynamicConfig.configure([1088], device=0)
    b = [[m] for m in _msgs(3, 700, 5)]
    r0 = big.digest_batch(b)
    small = pkg.Sha256DynamicConfig.configure([64], device=0)
    rs = small.digest_batch([[m] for m in _msgs(4, 30, 6)])
    r1 = big.digest_batch(b)
    assert (r0.checksums == r1.checksums).all() and (r0.digests == r1.digests).all()
    assert rs.digests.shape == (4, 32)
    big.close(); small.close()


def test_digests_only_batch_reports_no_expand_time(pkg):
    """h2sha_last_kernel_ms after a timed batch without an expansion step."""
    import torch
    cfg = pkg.Sha256DynamicConfig.configure([64], device=0)
    blob, offs, lens = pkg.pack_messages([[m] for m in _msgs(8, 20, 1)])
    dg = np.zeros((8, 32), dtype=np.uint8)
    cfg.digest_batch_raw(8, blob.ctypes.data, False, int(blob.size), offs, lens, None, digests_host_ptr=dg.ctypes.data, time_kernels=True)
    torch.cuda.synchronize()
    t_ms, x_ms = cfg.last_kernel_ms()
    assert t_ms > 0 and x_ms == 0.0
    cfg.close()


def test_zero_outputs_handles_more_than_65535_instances(pkg):
    """grid.y limit.  Uses a tiny row stride so that 70 000 instances fit easily."""
    import torch
    cfg = pkg.Sha256DynamicConfig.configure([64], device=0)
    lay = cfg.layout
    n = 70000
    # only the lookup buffer (smallest) is exercised at this size
    lookup = torch.full((n, lay.n_lookup_cols, lay.lookup_col_rows, 4), -1, dtype=torch.int64, device="cuda:0")
    rc = pkg.load_library().h2sha_zero_outputs(cfg._h, n, None, lookup.data_ptr(), None, 1, None)
    assert rc == 0
    torch.cuda.synchronize()
    tail = lookup[:, :, lay.n_lookup_cells:, :]
    if tail.numel():
        assert int(tail.abs().sum().item()) == 0
    assert int((lookup[n - 1, 0, 0] == -1).all().item()) == 1
    cfg.close()


def test_warp_per_message_trace_kernel_equals_thread_per_message(pkg, monkeypatch):
    """Small batches take k_trace_warp (one warp per message: latency), large ones k_trace (one thread per message).  Both must
    leave the same traces: every cell, digest and checksum of a batch with dynamic lengths, the padding edges (0, 55, 56, 63,
    64, 119 ...), a precomputed prefix and two digests per context is identical under either kernel."""
    sizes = [192, 1088]
    rng = np.random.default_rng(11)
    lens_a = [0, 1, 55, 56, 63, 64, 119, 120, 183, 100, 7, 64]
    msgs = [[bytes(rng.integers(0, 256, la, dtype=np.uint8)), bytes(rng.integers(0, 256, int(rng.integers(0, 1080)), dtype=np.uint8))] for la in lens_a]
    pre = [[0, 64 * int(rng.integers(0, 1 + len(m[1]) // 64))] for m in msgs]
    results = []
    for knob in ("tracewarp=0", "tracewarp=100000"):
        monkeypatch.setenv("H2SHA_TUNE", knob)
        cfg = pkg.Sha256DynamicConfig.configure(sizes, device=0)
        res = cfg.digest_batch(msgs, pre)
        results.append((res.digests.copy(), res.checksums.copy(), res.gate.cpu().numpy().copy(), res.lookup.cpu().numpy().copy(), res.spread.cpu().numpy().copy()))
        cfg.close()
    for a, b in zip(*results):
        assert (a == b).all()
    for k, m in enumerate(msgs):
        for d in range(2):
            assert bytes(results[0][0][2 * k + d]) == hashlib.sha256(m[d]).digest()
