"""CPU check of the NTT passes (stark_brainfuck_b200/csrc/ntt4.cuh + ntt4_plan.h): the kernel's
__host__ __device__ phase functions are run thread by thread on the host and compared with the C
oracle for every length 2^4 .. 2^16 (forward/inverse, coset offsets, zero padding, ragged n_in,
several planes)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ntt4_passes_on_host(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    exe = str(tmp_path / "ntt4_hostcheck")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "stark_brainfuck_b200", "csrc"),
                           os.path.join(ROOT, "tests", "ntt4_hostcheck.cpp"), "-L", os.path.join(ROOT, "oracle"),
                           "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", exe])
    out = subprocess.run([exe, "16"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failed" in out.stdout


def test_montgomery_and_lazy_arithmetic_on_host(tmp_path):
    """glmont.cuh compiled for the host: mont_mul / ladd / lsub / lcanon / x_mul_mont / x_fma_mont and the cached-power
    bookkeeping of the quotient kernel against plain 128-bit arithmetic (tests/glmont_hostcheck.cpp)"""
    exe = str(tmp_path / "glmont_hostcheck")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "stark_brainfuck_b200", "csrc"),
                           os.path.join(ROOT, "tests", "glmont_hostcheck.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "0 failed" in out.stdout, out.stdout + out.stderr
