"""The per-element NTT tables of round 2 (csrc/tables.cpp fill_full_table: inter-pass twiddles and zk_shift factors in the data layout
of a two-pass transform) against big-integer arithmetic, on the host.  The harness links the product's tables.cpp without a device; the
GPU tests check the same tables through the transforms that read them (tests/test_gpu_kernels.py)."""
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2013265921
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def rou(k, inverse=False):
    w = pow(137, 1 << (27 - k), P)          # ROU_FWD[27] = 137
    return pow(w, P - 2, P) if inverse else w


def bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    from boundless_b200 import build as _build
    inc = os.path.join(_build.CSRC, "constants.inc")
    if not os.path.exists(inc):
        subprocess.check_call([os.sys.executable, os.path.join(_build.HERE, "tools", "gen_constants.py"), inc])
    exe = os.path.join(str(tmp_path_factory.mktemp("tables")), "tables_on_host")
    cmd = ["g++", "-O2", "-std=c++17", "-I", _build.CSRC, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"), "-o", exe,
           os.path.join(ROOT, "tests", "host_emul", "tables_on_host.cpp"), "-x", "c++", os.path.join(_build.CSRC, "tables.cpp"),
           "-L" + os.path.join(CUDA, "lib64"), "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def entries(exe, kind, lg_m, lg_rows, idx):
    out = subprocess.run([exe, str(kind), str(lg_m), str(lg_rows)] + [str(i) for i in idx], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == 1 << lg_m
    return [int(x) for x in out[1:]]


def expected(kind, lg_m, lg_rows, i):
    lg_cols = lg_m - lg_rows
    rho, pos = i >> lg_cols, i & ((1 << lg_cols) - 1)
    if kind == 2:           # zk_shift: slot (rho, pos) of the bit-reversed coefficients holds degree bitrev(rho) + 2^lg_rows * bitrev(pos)
        return pow(3, bitrev(rho, lg_rows) + (bitrev(pos, lg_cols) << lg_rows), P)
    e = (pos * bitrev(rho, lg_rows)) % (1 << lg_m)
    if kind == 1:
        return pow(rou(lg_m), e, P)
    # inverse: the 1/2^m of the transform is folded in (1/2^10 for the sizes that take the three-pass route)
    scale = pow(1 << (10 if lg_m > 24 else lg_m), P - 2, P)
    return pow(rou(lg_m, True), e, P) * scale % P


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("lg_m,lg_rows", [(1, 0), (4, 2), (10, 0), (12, 4), (15, 5), (18, 10), (20, 10), (22, 10)])
def test_full_table_entries(harness, kind, lg_m, lg_rows):
    n = 1 << lg_m
    rng = random.Random(lg_m * 100 + lg_rows * 3 + kind)
    idx = list(range(n)) if n <= 4096 else sorted({0, 1, n - 1, n // 2, (1 << (lg_m - lg_rows)) + 1} | {rng.randrange(n) for _ in range(300)})
    got = entries(harness, kind, lg_m, lg_rows, idx)
    assert got == [expected(kind, lg_m, lg_rows, i) for i in idx]
