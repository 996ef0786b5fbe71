"""Measures the write-only HBM bandwidth of this GPU (the real ceiling of a pure witness-store stream)."""
import torch
n = 2_684_354_560 // 8
x = torch.empty(n, dtype=torch.int64, device="cuda")
y = torch.empty(n, dtype=torch.int64, device="cuda")
for name, fn in [("fill_", lambda: x.fill_(7)), ("zero_ (memset)", lambda: x.zero_()), ("copy_ (read+write)", lambda: y.copy_(x))]:
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = min(ts)
    bytes_ = n * 8 * (2 if "copy" in name else 1)
    print(f"{name:22s} {ms:7.3f} ms  {bytes_ / ms / 1e6:8.1f} GB/s")
