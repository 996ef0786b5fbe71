"""Host-side cost of h2sha_digest_batch per call (the call only enqueues): wall time of each of 12 back-to-back calls with host
buffers, for several batch sizes, with and without the copy-stream overlap (H2SHA_TUNE overlap=0)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
for n in (1024, 4096, 16384):
    cfg = pkg.Sha256DynamicConfig.configure([64], device=0)
    rng = np.random.default_rng(1)
    blob, offs, lens = pkg.pack_messages([[bytes(rng.integers(0, 256, 55, dtype=np.uint8))] for _ in range(n)])
    h_blob = torch.from_numpy(blob).pin_memory()
    outs = cfg.alloc_outputs(n)
    hd = torch.zeros((n, 32), dtype=torch.uint8).pin_memory(); hc = torch.zeros((n, 4), dtype=torch.int64).pin_memory()
    st = torch.cuda.current_stream(0)

    def call():
        cfg.digest_batch_raw(n, h_blob.data_ptr(), False, int(blob.size), offs, lens, None, gate_ptr=outs[0].data_ptr(), lookup_ptr=outs[1].data_ptr(),
                             spread_ptr=outs[2].data_ptr(), digests_host_ptr=hd.data_ptr(), checksums_host_ptr=hc.data_ptr(), stream=st.cuda_stream)

    for _ in range(3):
        call()
    torch.cuda.synchronize()
    ts = []
    t00 = time.perf_counter()
    for k in range(12):
        t0 = time.perf_counter(); call(); ts.append((time.perf_counter() - t0) * 1e6)
    host = time.perf_counter() - t00
    torch.cuda.synchronize()
    total = time.perf_counter() - t00
    print(f"n={n} tune={os.environ.get('H2SHA_TUNE')}: per-call host us = {[int(t) for t in ts]}; host {host*1e3:.2f} ms of {total*1e3:.2f} ms", flush=True)
    cfg.close()
    del outs
    torch.cuda.empty_cache()
