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


def _quotient_hostcheck_input():
    """cases for tests/quotient_hostcheck.cpp: the reference's real AIR (tests/golden/air.json) with the base columns
    flagged and unflagged, and random programs with the shapes the compiler special-cases"""
    import random
    import numpy as np
    from util import P, golden, quotient_program
    R = random.Random(77)
    cases = []

    def add(width, kinds, prog, n_points, stage):
        off, coeffs, factors = prog
        vals = []
        for _ in range(n_points):
            for v in range(2 * width):
                x = [R.randrange(P), R.randrange(P), R.randrange(P)]
                if R.random() < 0.1:
                    x[R.randrange(3)] = R.choice([0, 1, P - 1])
                if kinds[v % width]:
                    x[1] = x[2] = 0
                vals += x
        words = [width, len(off) - 1, factors.shape[1], n_points, int(stage)] + list(kinds) + off.tolist() + \
            coeffs.reshape(-1).tolist() + factors.reshape(-1).tolist() + vals
        cases.append(" ".join(str(int(w)) for w in words))
    for t in golden("air.json")["tables"]:
        W = t["full_width"]
        for name in ("boundary", "transition", "terminal"):
            prog = quotient_program(t[name])
            for kinds in ([int(j < t["base_width"]) for j in range(W)], [0] * W):
                add(W, kinds, prog, 4, 0)
    # random programs: repeated variables, equal exponent vectors, constants, the zero polynomial, exponents to 9,
    # and a wide table (2 * 40 variables)
    for width, n_cons, n_mono, max_f in ((3, 4, 12, 3), (2, 3, 40, 4), (40, 2, 60, 5), (1, 2, 6, 2), (6, 5, 25, 6)):
        kinds = [int(R.random() < 0.5) for _ in range(width)]
        off, coeffs, facs = [0], [], []
        for c in range(n_cons):
            k = 0 if (c == 1 and width == 1) else R.randrange(1, n_mono)
            for _ in range(k):
                coeffs.append([R.randrange(P), R.choice([0, R.randrange(P)]), R.choice([0, R.randrange(P)])])
                nf = R.randrange(0, max_f + 1)
                facs.append([(R.randrange(2 * width) << 8) | R.randrange(1, 10 if max_f < 4 else 4) for _ in range(nf)])
            if k > 2:  # the same exponent vector twice, and a bare constant
                coeffs.append([R.randrange(P), 0, 0])
                facs.append(list(facs[-1]))
                coeffs.append([R.randrange(P), R.randrange(P), 0])
                facs.append([])
            off.append(len(coeffs))
        factors = np.zeros((len(facs), max_f), dtype=np.uint32)
        for m, f in enumerate(facs):
            factors[m, :len(f)] = f
        prog = (np.asarray(off, dtype=np.uint32), np.asarray(coeffs, dtype=np.uint64).reshape(-1, 3), factors)
        add(width, kinds, prog, 4, 1)
    return "%d\n%s\n" % (len(cases), "\n".join(cases))


def test_quotient_programs_on_host(tmp_path):
    """quotient_prog.h (Horner compiler + the kernel's interpreter) compiled for the host against a direct
    monomial-by-monomial evaluation, on the reference's AIR and on random programs"""
    exe = str(tmp_path / "quotient_hostcheck")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "stark_brainfuck_b200", "csrc"),
                           os.path.join(ROOT, "tests", "quotient_hostcheck.cpp"), "-o", exe])
    out = subprocess.run([exe], input=_quotient_hostcheck_input(), capture_output=True, text=True)
    assert out.returncode == 0 and " 0 failed" in out.stdout, out.stdout + out.stderr
    assert int(out.stdout.split()[0]) > 800  # one and two points per thread
