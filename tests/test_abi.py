"""The C-ABI library loads and exports every symbol include/h2sha_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "h2sha_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    funcs = set(re.findall(r"\b(h2sha_[a-z0-9_]+)\s*\(", src))
    data = set(re.findall(r"extern\s+const\s+\w+\s+(H2SHA_[A-Z0-9_]+)\s*\[", src))
    return funcs | data


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 12
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/h2sha_b200.h but not exported"
    assert declared == set(pkg.EXPORTED_SYMBOLS)


def test_library_has_no_python_or_torch_dependency(pkg):
    import subprocess
    out = subprocess.run(["ldd", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcudart" in out
    needed = [ln.split("=>")[0].strip() for ln in out.splitlines()]
    assert not any(n.startswith(("libtorch", "libpython", "libc10")) for n in needed), needed


def test_checksum_multipliers_match_oracle(pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    m = (ctypes.c_uint32 * 8).in_dll(lib, "H2SHA_CK_M")
    src = open(os.path.join(ROOT, "oracle", "h2sha_oracle.c")).read()
    vals = re.search(r"CK_M\[8\] = \{([^}]*)\}", src).group(1)
    assert [int(x.strip().rstrip("u"), 16) for x in vals.split(",")] == list(m)


def test_no_device_means_error_not_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        return
    try:
        pkg.Sha256DynamicConfig.configure([64], device=0)
    except pkg.EngineError as e:
        assert e.code == pkg.H2SHA_ECUDA
    else:
        raise AssertionError("engine creation must fail without a CUDA device")


def test_gather_rejects_a_null_communicator(pkg):
    """h2sha_gather resolves NCCL at run time; argument errors come back as codes, never as crashes"""
    L = pkg.load_library()
    assert L.h2sha_gather(None, 1, 1, None, None, None, None, None) == pkg.H2SHA_EINVAL
    assert b"communicator" in L.h2sha_last_error()


def test_header_is_plain_c_and_usable_from_c(pkg):
    """include/h2sha_b200.h compiles as C99 (what bindgen / cgo / JNI parse) and a C program can drive the plan queries"""
    import subprocess
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "abi_c_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "cpp", "abi_c_check.c"),
                           "-L" + libdir, "-lh2sha_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("ok blocks=4 gate_cols=3"), out.stdout + out.stderr
