"""bench.py's arms: the CPU reference arm prints its one JSON line without a GPU; on a GPU the whole-configuration mode
streams a configuration through HBM in chunks and checks what it produced."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.strip()]
    assert len(lines) == 1, f"stdout must be exactly one JSON line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "8")
    assert d["impl"] == "reference" and d["unit"] == "blocks/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0


@pytest.mark.gpu
def test_full_workload_mode_streams_config3_in_chunks():
    d = _run("--workload", "cfg3", "--full-workload", "--steps", "1")
    assert d["config"]["instances_total"] == 4096 and d["config"]["blocks_per_instance"] == 17
    assert d["config"]["chunks_per_gpu"] >= 2                      # 181 GB of witness do not fit in one pass
    assert sum(d["checked"]["violations"].values()) == 0 and d["checked"]["oracle_instances"] == 8
    assert d["value"] > 1e5 and 0 < d["roofline"]["frac"] < 1.2
