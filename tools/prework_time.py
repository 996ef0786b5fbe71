"""Times the lookup pre-work calls on one GPU: python tools/prework_time.py [workload] [instances] [permute instances]."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
n_perm = int(sys.argv[3]) if len(sys.argv) > 3 else 256
ge.build()
pkg = ge.load_package()
S = ge.load_package_module("synthetic")
w = S.WORKLOADS[wl]
cfg = pkg.Sha256DynamicConfig.configure(list(w.max_variable_byte_sizes), device=0)
lay = cfg.layout
blob, offs, lens = S.generate(w, 0, n)
msgs = [[bytes(blob[int(o):int(o) + int(l)])] for o, l in zip(offs, lens)]
res = cfg.digest_batch(msgs)
usable = (1 << 17) - 6
info = cfg.lookup_info()
L = pkg.load_library()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev)
mult = torch.empty((n, info["mult_words_per_instance"]), dtype=torch.int32, device=dev)
bad = torch.zeros(1, dtype=torch.int32, device=dev)
a = torch.empty((n_perm, usable, 4), dtype=torch.int64, device=dev)
s = torch.empty((n_perm, usable, 4), dtype=torch.int64, device=dev)


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); fn(); e1.record(st); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:])


t_m = timed(lambda: L.h2sha_lookup_multiplicities(cfg._h, n, res.lookup.data_ptr(), res.spread.data_ptr(), usable, mult.data_ptr(), bad.data_ptr(), st.cuda_stream))
print(f"multiplicities: {n} instances {t_m:.3f} ms, reads {n * (lay.n_lookup_cells + 2 * lay.n_spread_limbs) * 32 / t_m / 1e6:.0f} GB/s, bad={int(bad.item())}")
theta = np.array([3, 5, 7, 11], dtype=np.uint64)
for l in range(info["n_range_lookups"] + info["n_spread_lookups"]):
    th = theta.ctypes.data if l >= info["n_range_lookups"] else None
    t_p = timed(lambda: L.h2sha_permute_lookup(cfg._h, n_perm, l, mult.data_ptr(), usable, th, a.data_ptr(), s.data_ptr(), None, st.cuda_stream))
    print(f"permute lookup {l}: {n_perm} instances {t_p:.3f} ms, writes {n_perm * usable * 64 / t_p / 1e6:.0f} GB/s")
