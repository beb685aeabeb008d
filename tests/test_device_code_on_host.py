"""The product's device arithmetic SOURCE (csrc/field.cuh, csrc/poseidon2.cuh), compiled for the host by tests/host_emul, must agree with
big-integer arithmetic and with the Poseidon2 known-answer vector.  This exercises the very headers the GPU kernels
include -- Montgomery reduction variant, Shoup diagonal multiplies, round constants, linear layers -- without a GPU.  It is a test
harness, not a CPU path: the library itself still fails loudly without a device (test_abi.py)."""
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2013265921
KAT = ("2ed3e23d 12921fb0 0e659e79 61d81dc9 32bae33b 62486ae3 1e681b60 24b91325 2a2ef5b9 50e8593e 5bc818ec 10691997 "
       "35a14520 2ba6a3c5 279d47ec 55014e81 5953a67f 2f403111 6b8828ff 1801301f 2749207a 3dc9cf21 3c985ba2 57a99864")


def _build(tmp, *defines):
    exe = os.path.join(str(tmp), "device_on_host" + "".join(d.replace("=", "_") for d in defines))
    cmd = ["g++", "-O2", "-std=c++17"] + ["-D" + d for d in defines] + ["-o", exe, os.path.join(ROOT, "tests", "host_emul", "device_on_host.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    return _build(tmp_path_factory.mktemp("host_emul"))


def run(exe, *args):
    return subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, check=True).stdout.split()


def test_poseidon2_kat_from_the_device_source(harness):
    assert " ".join(run(harness, "kat")) == KAT


@pytest.mark.parametrize("defines", [("B200_REDC_V=0",), ("B200_REDC_V=1",), ("B200_REDC_V=2",), ("B200_P2_SHOUP=0",), ("B200_P2_MERGED=1",),
                                     ("B200_P2_Z=63",), ("B200_P2_V=15",), ("B200_ADD_V=1",), ("B200_P2_UNROLL_INT=3", "B200_P2_UNROLL_EXT=2"),
                                     ("B200_P2_LAZY=1",), ("B200_P2_LAZY=2",), ("B200_P2_LAZY=4",), ("B200_P2_LAZY=7",), ("B200_P2_LAZY=3",),
                                     ("B200_REDC_V=3",), ("B200_REDC_V=3", "B200_P2_LAZY=7"), ("B200_P2_NMACC=1",), ("B200_P2_NMACC=2", "B200_P2_LAZY=1"),
                                     ("B200_P2_NMACC=8", "B200_P2_LAZY=1"), ("B200_P2_NMACC=23",), ("B200_P2_NMACC=24", "B200_P2_LAZY=1"),
                                     ("B200_P2_LAZY=3", "B200_P2_SHOUP=0"), ("B200_P2_LAZY=0", "B200_P2_ZALL=0"), ("B200_P2_LAZY=0",), ("B200_P2_ZALL=0",),
                                     ("B200_P2_LAZY=1", "B200_P2_NMACC=6", "B200_P2_ZALL=1")])
def test_every_build_variant_passes_the_kat(tmp_path, harness, defines):
    """the measured-and-rejected formulations and the prepared lazy-reduction variants kept behind macros stay
    correct: the KAT, and 3 x 2002 chained permutations of random / all-zero / all-(p-1) states against the default build"""
    exe = _build(tmp_path, *defines)
    assert " ".join(run(exe, "kat")) == KAT
    assert run(exe, "perms", 12345, 2000) == run(harness, "perms", 12345, 2000)


def test_field_ops_against_big_integers(harness):
    rng = random.Random(0xB200)
    cases = [(0, 0), (0, 1), (1, 0), (P - 1, P - 1), (P - 1, 1), (1, P - 1), (2, 30), (123456789, 987654321)]
    cases += [(rng.randrange(P), rng.randrange(P)) for _ in range(40)]
    for a, b in cases:
        got = [int(x) for x in run(harness, "ops", a, b)]
        want = [(a + b) % P, (a - b) % P, (-a) % P, (2 * a) % P, (a * b) % P, (a * a) % P, pow(a, b, P)]
        assert got == want, (a, b)


def test_extension_field_product(harness):
    """ExtElem multiply over X^4 = -11 (risc0-zkp field/baby_bear.rs)"""
    rng = random.Random(4)
    for _ in range(20):
        a = [rng.randrange(P) for _ in range(4)]
        b = [rng.randrange(P) for _ in range(4)]
        prod = [0] * 7
        for i in range(4):
            for j in range(4):
                prod[i + j] += a[i] * b[j]
        want = [(prod[k] - 11 * (prod[k + 4] if k + 4 < 7 else 0)) % P for k in range(4)]
        assert [int(x) for x in run(harness, "fp4", *a, *b)] == want
