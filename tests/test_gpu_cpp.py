"""Builds and runs the C++ host-side parity test (tests/cpp/test_reference_vectors.cc) that mirrors the reference's own
test_sha256_correct1..4 over the C-ABI: no Python between the test and libh2sha_b200.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(pkg):
    from oracle import oracle as O
    O.build()
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_reference_vectors")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_reference_vectors.cc")
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O1", "-o", exe, src, "-L" + libdir, "-lh2sha_b200", "-ldl",
                           "-Xlinker", "-rpath," + libdir, "-cudart", "shared"])
    return exe


def test_cpp_host_test_compiles_and_links(pkg):
    """CPU part: the C++ host API (csrc/host_api.hpp) compiles against include/h2sha_b200.h and links the library."""
    assert os.path.exists(_build(pkg))


@pytest.mark.gpu
def test_reference_tests_through_cpp_host_api(pkg):
    exe = _build(pkg)
    out = subprocess.run([exe, os.path.join(ROOT, "oracle", "_build", "libh2sha_oracle.so")], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok ") == 5


def test_host_field_helpers():
    """csrc/fr_host.h: the CIOS Montgomery product (spread-table compression of h2sha_permute_lookup) against the
    double-and-add product; plain g++, no GPU."""
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_fr_host")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_fr_host.cc")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


def _build_native_runner(pkg):
    exe = os.path.join(ROOT, "tools", "native_runner")
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(ROOT, "tools", "native_runner.cc"),
                           "-L" + libdir, "-lh2sha_b200", "-lnccl", "-Xlinker", "-rpath,$ORIGIN/../halo2-dynamic-sha256_b200", "-cudart", "shared"])
    return exe


def test_native_runner_compiles_and_links(pkg):
    """The no-Python host program (one thread per GPU, NCCL gather) builds against the C-ABI; without a GPU it refuses to run."""
    exe = _build_native_runner(pkg)
    import torch
    if not torch.cuda.is_available():
        out = subprocess.run([exe, "--gpus", "1"], capture_output=True, text=True, timeout=60)
        assert out.returncode == 3 and "no CPU path" in out.stderr


@pytest.mark.gpu
def test_native_runner_one_gpu(pkg):
    import json
    exe = _build_native_runner(pkg)
    out = subprocess.run([exe, "--workload", "cfg3", "--instances", "64", "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["digest_mismatches"] == 0 and line["blocks_per_instance"] == 17 and line["value"] > 0
