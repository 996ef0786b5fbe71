"""Sweeps the engine's tuning knobs (H2SHA_TUNE) on one GPU and prints k_expand time per setting.
usage: python tools/tune.py [workload] "parts=3,fill=128,cons=16,prod=4" "..." """
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
if os.environ.get("TUNE_LIB"):   # A/B runs: time another build of the engine (e.g. tools/ab/libh2sha_base.so) with the same harness
    pkg.LIB_PATH = os.path.abspath(os.environ["TUNE_LIB"])
S = ge.load_package_module("synthetic")


def run(workload, tune, n_inst=None, steps=10):
    os.environ["H2SHA_TUNE"] = tune
    w = S.WORKLOADS[workload]
    try:
        cfg = pkg.Sha256DynamicConfig.configure(list(w.max_variable_byte_sizes), device=0)
    except Exception as e:
        return f"{tune:40s} ERROR {e}"
    lay = cfg.layout
    n = n_inst or min(w.n_instances, int(0.5 * torch.cuda.mem_get_info()[0]) // lay.bytes_per_instance)
    blob, offs, lens = S.generate(w, 0, n)
    gate, lookup, spread = cfg.alloc_outputs(n, zero=False)
    d_blob = torch.from_numpy(np.concatenate([blob, np.zeros(16, np.uint8)])).cuda()
    dd = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    dc = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    kw = dict(gate_ptr=gate.data_ptr(), lookup_ptr=lookup.data_ptr(), spread_ptr=spread.data_ptr(), digests_dev_ptr=dd.data_ptr(),
              checksums_dev_ptr=dc.data_ptr(), stream=sp)
    cfg.digest_batch_raw(n, d_blob.data_ptr(), True, int(blob.size), offs, lens, None, **kw)
    for _ in range(3):
        cfg.digest_batch_raw(n, 0, True, 0, offs, lens, None, reuse_inputs=True, **kw)
    ts, tr = [], []
    for _ in range(steps):
        cfg.digest_batch_raw(n, 0, True, 0, offs, lens, None, reuse_inputs=True, time_kernels=True, **kw)
        ts.append(cfg.last_kernel_ms()[1]); tr.append(cfg.last_kernel_ms()[0])
    torch.cuda.synchronize()
    sustained = ""
    if os.environ.get("TUNE_SUSTAIN"):
        # back-to-back launches for ~1.5 s (power-cap regime), then the median kernel time of 20 timed launches
        import time
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 1.5:
            for _ in range(50):
                cfg.digest_batch_raw(n, 0, True, 0, offs, lens, None, reuse_inputs=True, **kw)
            torch.cuda.synchronize()
        t2 = []
        for _ in range(20):
            for _ in range(5):
                cfg.digest_batch_raw(n, 0, True, 0, offs, lens, None, reuse_inputs=True, **kw)
            cfg.digest_batch_raw(n, 0, True, 0, offs, lens, None, reuse_inputs=True, time_kernels=True, **kw)
            t2.append(cfg.last_kernel_ms()[1])
        sustained = f"  sustained {float(np.median(t2)):.4f} ms"
    ck = int(dc.sum().item()) & ((1 << 64) - 1)
    cfg.close()
    del gate, lookup, spread
    torch.cuda.empty_cache()
    ms = float(np.median(ts))
    gbs = n * lay.cells_per_instance * 32 / ms / 1e6
    return f"{tune:40s} n={n:6d} k_expand {ms:8.4f} ms  {n * lay.n_blocks / ms / 1e3:8.3f} Mblk/s  {gbs:7.1f} GB/s  ck={ck:#x}{sustained}  k_trace {float(np.median(tr)) * 1e3:.1f} us"


if __name__ == "__main__":
    wl = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1].startswith("cfg") else "cfg2"
    tunes = [a for a in sys.argv[1:] if not a.startswith("cfg")] or ["parts=3,fill=128,cons=16,prod=4"]
    for t in tunes:
        print(run(wl, t, n_inst=int(os.environ.get("TUNE_N", "0")) or None), flush=True)
