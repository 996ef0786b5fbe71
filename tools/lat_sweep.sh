# latency of one digest() (BASELINE config 1 and the bench circuit's shape) under a few engine tunings
python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -x -q -m gpu 2>&1 | tail -2 | cut -c1-400
H2SHA_TUNE="parts=12" python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2 | cut -c1-400
for t in "parts=3" "parts=17" "parts=24"; do echo "== $t"; H2SHA_TUNE="$t" REPS=6 python tools/latency_probe.py 128 2>&1 | tail -2; H2SHA_TUNE="$t" REPS=6 python tools/latency_probe.py 1024 2>&1 | tail -1; H2SHA_TUNE="$t" TIMED=0 REPS=6 python tools/latency_probe.py 128 2>&1 | tail -1; done
bash tools/ab.sh cfg2 3
