"""The kernels' pure integer helpers (fastkmer_b200/csrc/fkm_math.h: reverse complements, the closed-form norm,
canonical records, k-mer extraction from a record, table slots) compiled for the host and checked against
string-level definitions — the same source the device code is built from.  CPU only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_device_math_on_host(tmp_path):
    exe = tmp_path / "host_math_check"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unused-function", "-I", os.path.join(ROOT, "fastkmer_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_math_check.cpp"), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout[-2000:]
