"""Does alternating two engine handles on two streams (consecutive batches overlap at launch boundaries) beat one stream?
python tools/two_stream_probe.py [workload] [instances]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
pkg = ge.load_package()
S = ge.load_package_module("synthetic")
w = S.WORKLOADS[wl]
dev = torch.device("cuda", 0)
blob, offs, lens = S.generate(w, 0, n)
d_blob = torch.from_numpy(np.concatenate([blob, np.zeros(16, np.uint8)])).cuda()
h_blob = torch.from_numpy(np.concatenate([blob, np.zeros(16, np.uint8)])).pin_memory()


class Lane:
    def __init__(self):
        self.cfg = pkg.Sha256DynamicConfig.configure(list(w.max_variable_byte_sizes), device=0)
        self.stream = torch.cuda.Stream(dev)
        self.out = self.cfg.alloc_outputs(n, zero=True)
        self.dd = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
        self.dc = torch.zeros((n, 4), dtype=torch.int64, device=dev)
        self.hd = torch.zeros((n, 32), dtype=torch.uint8).pin_memory()
        self.hc = torch.zeros((n, 4), dtype=torch.int64).pin_memory()
        self.kw = dict(gate_ptr=self.out[0].data_ptr(), lookup_ptr=self.out[1].data_ptr(), spread_ptr=self.out[2].data_ptr(), stream=self.stream.cuda_stream)
        self.cfg.digest_batch_raw(n, d_blob.data_ptr(), True, int(blob.size), offs, lens, None, digests_dev_ptr=self.dd.data_ptr(), checksums_dev_ptr=self.dc.data_ptr(), **self.kw)
        self.stream.synchronize()

    def resident(self):
        self.cfg.digest_batch_raw(n, 0, True, 0, offs, lens, None, reuse_inputs=True, digests_dev_ptr=self.dd.data_ptr(), checksums_dev_ptr=self.dc.data_ptr(), **self.kw)

    def e2e(self):
        self.cfg.digest_batch_raw(n, h_blob.data_ptr(), False, int(blob.size), offs, lens, None, digests_host_ptr=self.hd.data_ptr(), checksums_host_ptr=self.hc.data_ptr(), **self.kw)


lanes = [Lane(), Lane()]
blocks = n * lanes[0].cfg.layout.n_blocks


def run(kind, n_lanes, steps=200, sustain=1.5):
    t_end = time.perf_counter() + sustain
    i = 0
    while time.perf_counter() < t_end:     # power-capped regime first
        for _ in range(20):
            ln = lanes[i % n_lanes]; i += 1
            if kind == "e2e":
                ln.stream.synchronize()
            getattr(ln, kind)()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        ln = lanes[k % n_lanes]
        if kind == "e2e":
            ln.stream.synchronize()        # the step that used this lane's host buffers before has been read
        getattr(ln, kind)()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return steps * blocks / dt


for kind in ("resident", "e2e"):
    for nl in (1, 2, 1, 2):
        print(f"{wl} n={n} {kind:8s} lanes={nl}: {run(kind, nl) / 1e6:.3f} Mblk/s", flush=True)
