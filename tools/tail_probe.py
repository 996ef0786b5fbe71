"""Head / tail of one k_expand launch: runs tools/tune.py with a -DH2SHA_DEBUG_TIMING=2 build and summarises, for the last launch, when the
first and the last consumer warp of every CTA finished and how long they waited for their producers.
usage: bash tools/build_ab.sh WORK dbg2 -DH2SHA_DEBUG_TIMING=2 && python tools/tail_probe.py [workload] [H2SHA_TUNE string]"""
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
tune = sys.argv[2] if len(sys.argv) > 2 else "parts=3"
env = dict(os.environ, TUNE_LIB=os.path.join(ROOT, "tools/ab/libh2sha_dbg2.so"))
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools/tune.py"), wl, tune], env=env, capture_output=True, text=True).stdout
rows = [(int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5)))
        for m in re.finditer(r"E cta\s+(\d+) warp\s+(\d+): done at \+(\d+) ns, waited (\d+) ns in (\d+) waits", out)]
n_cta = len({r[0] for r in rows})
last = rows[-2 * n_cta:]
done = np.array([r[2] for r in last]) / 1e3
waited = np.array([r[3] for r in last]) / 1e3
per_cta = {}
for r in last:
    per_cta.setdefault(r[0], []).append(r[2] / 1e3)
cta_done = np.array([max(v) for v in per_cta.values()])
print([l for l in out.splitlines() if l.startswith(tune)][-1])
print(f"{n_cta} CTAs; consumer warps done at (us after the trace kernel): min {done.min():.1f} median {np.median(done):.1f} max {done.max():.1f}")
print(f"CTA finish times: min {cta_done.min():.1f}  p10 {np.percentile(cta_done, 10):.1f}  median {np.median(cta_done):.1f}  p90 {np.percentile(cta_done, 90):.1f}  max {cta_done.max():.1f}"
      f"  -> idle SM time at the tail {100 * (cta_done.max() - cta_done.mean()) / cta_done.max():.1f} % of the launch")
print(f"time a consumer warp spent waiting for producers: median {np.median(waited):.1f} us, max {waited.max():.1f} us")
