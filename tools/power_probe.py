"""Power / clock under sustained load for: the expansion kernel, the same without its global stores (checksums only), the
plain store probe and the integer probe.  python tools/power_probe.py"""
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
if os.environ.get("TUNE_LIB"):   # another build of the engine (experiment variants under tools/ab/)
    pkg.LIB_PATH = os.path.abspath(os.environ["TUNE_LIB"])
S = ge.load_package_module("synthetic")
w = S.WORKLOADS["cfg2"]
n = int(os.environ.get("TUNE_N", "1024"))
dev = torch.device("cuda", 0)
cfg = pkg.Sha256DynamicConfig.configure(list(w.max_variable_byte_sizes), device=0)
lay = cfg.layout
blob, offs, lens = S.generate(w, 0, n)
d_blob = torch.from_numpy(np.concatenate([blob, np.zeros(16, np.uint8)])).cuda()
gate, lookup, spread = cfg.alloc_outputs(n, zero=True)
dd = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
dc = torch.zeros((n, 4), dtype=torch.int64, device=dev)
st = torch.cuda.current_stream(dev)
sp = st.cuda_stream
cfg.digest_batch_raw(n, d_blob.data_ptr(), True, int(blob.size), offs, lens, None, gate_ptr=gate.data_ptr(), lookup_ptr=lookup.data_ptr(),
                     spread_ptr=spread.data_ptr(), digests_dev_ptr=dd.data_ptr(), checksums_dev_ptr=dc.data_ptr(), stream=sp)
probe = torch.empty(n * lay.bytes_per_instance, dtype=torch.uint8, device=dev)
scratch = torch.zeros(148 * 8 * 256 * 2, dtype=torch.int32, device=dev)


def full():
    cfg.digest_batch_raw(n, 0, True, 0, offs, lens, None, reuse_inputs=True, gate_ptr=gate.data_ptr(), lookup_ptr=lookup.data_ptr(),
                         spread_ptr=spread.data_ptr(), digests_dev_ptr=dd.data_ptr(), checksums_dev_ptr=dc.data_ptr(), stream=sp)


def nostore():
    cfg.digest_batch_raw(n, 0, True, 0, offs, lens, None, reuse_inputs=True, digests_dev_ptr=dd.data_ptr(), checksums_dev_ptr=dc.data_ptr(), stream=sp)


def store():
    cfg.store_probe(probe.data_ptr(), probe.numel(), sp)


def ints():
    cfg.int_probe(scratch.data_ptr(), 1 << 12, sp)


Q = "clocks.sm,power.draw,clocks_event_reasons.sw_power_cap"
modes = (("k_expand", full), ("k_expand without stores", nostore), ("store probe", store), ("int probe", ints), ("k_expand", full))
if os.environ.get("POWER_MODES") == "expand":
    modes = (("k_expand", full), ("k_expand without stores", nostore))
for name, fn in modes:
    p = subprocess.Popen(["nvidia-smi", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-i", "0", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < 3.0:
        for _ in range(20):
            fn(); k += 1
        st.synchronize()
    dt = time.perf_counter() - t0
    p.terminate()
    out = p.communicate()[0].strip().splitlines()
    rows = [[x.strip() for x in ln.split(",")] for ln in out][len(out) // 2:]     # second half: settled
    clk = np.median([float(r[0]) for r in rows]); pw = np.median([float(r[1]) for r in rows]); cap = sum(r[2].startswith("Active") for r in rows) / max(1, len(rows))
    print(f"{name:26s} {1e3 * dt / k:8.4f} ms/launch  sm {clk:6.0f} MHz  power {pw:6.0f} W  power-cap active {100 * cap:3.0f}% of samples", flush=True)
    time.sleep(1.0)
