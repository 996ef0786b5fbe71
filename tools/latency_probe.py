"""One-instance latency timeline: runs a single digest a few times with a -DH2SHA_DEBUG_TIMING build (TUNE_LIB) so that the kernel prints
its producer / consumer timestamps, and reports wall / device times.  usage: TUNE_LIB=tools/ab/libh2sha_dbg.so python tools/latency_probe.py [max_bytes]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
if os.environ.get("TUNE_LIB"):
    pkg.LIB_PATH = os.path.abspath(os.environ["TUNE_LIB"])
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = pkg.Sha256DynamicConfig.configure([mb], device=0)
gate, lookup, spread = cfg.alloc_outputs(1)
blob, offs, lens = pkg.pack_messages([[b"\x01" * 56]])
hb = torch.from_numpy(np.concatenate([blob, np.zeros(16, np.uint8)])).pin_memory()
hd = torch.zeros((1, 32), dtype=torch.uint8).pin_memory(); hc = torch.zeros((1, 4), dtype=torch.int64).pin_memory()
st = torch.cuda.current_stream(0)
n = int(os.environ.get("REPS", "3"))
for i in range(n):
    t0 = time.perf_counter()
    cfg.digest_batch_raw(1, hb.data_ptr(), False, int(blob.size), offs, lens, None, gate_ptr=gate.data_ptr(), lookup_ptr=lookup.data_ptr(), spread_ptr=spread.data_ptr(),
                         digests_host_ptr=hd.data_ptr(), checksums_host_ptr=hc.data_ptr(), stream=st.cuda_stream, time_kernels=bool(int(os.environ.get("TIMED", "1"))))
    st.synchronize()
    dt = time.perf_counter() - t0
    km = cfg.last_kernel_ms() if int(os.environ.get("TIMED", "1")) else (0, 0)
    print(f"--- run {i}: wall {dt*1e6:.1f} us, k_trace {km[0]*1e3:.1f} us, k_expand {km[1]*1e3:.1f} us", flush=True)
