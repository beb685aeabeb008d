"""The risc0-sys / sppark symbol-compatible exports (include/b200_risc0_sys_compat.h) against the CPU oracle, bit-exact.

These are the names risc0-zkp's CUDA Hal binds (sppark_batch_iNTT, sppark_batch_NTT, sppark_batch_expand, sppark_batch_zk_shift,
sppark_poseidon2_rows, sppark_poseidon2_fold, supra_poly_divide); errors come back as sppark::Error by value."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 2013265921


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint32)


def ptr(t):
    return C.c_void_p(t.data_ptr())


def ok(err):
    from boundless_b200 import lib
    lib.check_sppark(err)


def rand(rng, shape, oracle):
    return oracle.to_mont(rng.integers(0, P, shape, dtype=np.int64))


def test_init(gpu, b200lib):
    ok(b200lib.sppark_init())
    ok(b200lib.sppark_init())


@pytest.mark.parametrize("lg_n,count", [(4, 3), (10, 16), (13, 5), (16, 4), (20, 2)])
def test_intt_shift_ntt(gpu, b200lib, oracle, lg_n, count):
    """Hal::batch_interpolate_ntt -> zk_shift -> batch_evaluate_ntt; the calls are synchronous (no torch sync needed)."""
    torch = gpu
    a = rand(np.random.default_rng(lg_n + count), count << lg_n, oracle)
    d = dev(torch, a)
    ok(b200lib.sppark_batch_iNTT(ptr(d), lg_n, count))
    ref = oracle.batch_intt(a, lg_n, count)
    assert np.array_equal(host(d), ref)
    ok(b200lib.sppark_batch_zk_shift(ptr(d), lg_n, count))
    ref = oracle.batch_zk_shift(ref, lg_n, count)
    assert np.array_equal(host(d), ref)
    ok(b200lib.sppark_batch_NTT(ptr(d), lg_n, count))
    assert np.array_equal(host(d), oracle.batch_ntt(ref, lg_n, count))


@pytest.mark.parametrize("lg_n,lg_b,count", [(3, 2, 2), (10, 2, 16), (12, 1, 3), (14, 3, 2), (18, 2, 4), (20, 2, 2)])
def test_expand_then_ntt_is_the_low_degree_extension(gpu, b200lib, oracle, lg_n, lg_b, count):
    """Hal::batch_expand_into_evaluate_ntt = sppark_batch_expand + sppark_batch_NTT == the oracle's expand-into-evaluate
    (and == the fused b200_batch_expand_ntt)."""
    torch = gpu
    a = rand(np.random.default_rng(100 * lg_n + lg_b), count << lg_n, oracle)
    d_in = dev(torch, a)
    d_out = torch.full((count << (lg_n + lg_b),), -1, dtype=torch.int32, device="cuda")
    ok(b200lib.sppark_batch_expand(ptr(d_out), ptr(d_in), lg_n, lg_b, count))
    spread = host(d_out).reshape(count << lg_n, 1 << lg_b)
    assert np.array_equal(spread[:, 0], a) and not spread[:, 1:].any()
    ok(b200lib.sppark_batch_NTT(ptr(d_out), lg_n + lg_b, count))
    ref = oracle.batch_expand_ntt(a, lg_n, count, lg_b)
    assert np.array_equal(host(d_out), ref)
    d_fused = torch.empty_like(d_out)
    assert b200lib.b200_batch_expand_ntt(ptr(d_fused), ptr(d_in), lg_n, lg_b, count, None) is None
    torch.cuda.synchronize()
    assert np.array_equal(host(d_fused), ref)


def test_expand_rejects_oversize(gpu, b200lib):
    from boundless_b200 import lib
    err = b200lib.sppark_batch_expand(None, None, 25, 2, 1)          # 2^27 > 2^26, the largest transform (po2 24 x blow-up 4)
    with pytest.raises(lib.B200Error, match="2\\^26"):
        lib.check_sppark(err)


@pytest.mark.parametrize("rows,cols", [(1, 1), (37, 16), (256, 17), (4096, 33), (1 << 16, 48)])
def test_poseidon2_rows_and_fold(gpu, b200lib, oracle, rows, cols):
    torch = gpu
    m = rand(np.random.default_rng(rows + cols), rows * cols, oracle)
    d_m = dev(torch, m)
    d_leaf = torch.empty(rows * 8, dtype=torch.int32, device="cuda")
    ok(b200lib.sppark_poseidon2_rows(ptr(d_leaf), ptr(d_m), rows, cols))
    leaves = oracle.hash_rows(m, rows, cols)
    assert np.array_equal(host(d_leaf), leaves.reshape(-1))
    n_out = rows // 2
    if n_out:
        d_par = torch.empty(n_out * 8, dtype=torch.int32, device="cuda")
        ok(b200lib.sppark_poseidon2_fold(ptr(d_par), ptr(d_leaf), n_out))
        lv = leaves.reshape(rows, 8)
        ref = np.stack([oracle.hash_pair(lv[2 * i], lv[2 * i + 1]) for i in range(min(n_out, 64))])
        assert np.array_equal(host(d_par).reshape(n_out, 8)[:ref.shape[0]], ref)


@pytest.mark.parametrize("size", [1, 9, 2049, 1 << 16, (1 << 18) + 5])
def test_supra_poly_divide_host_scalars(gpu, b200lib, oracle, size):
    """polynomial on the device, pow / remainder in HOST memory (the original's calling convention)."""
    torch = gpu
    rng = np.random.default_rng(size)
    p = rand(rng, (size, 4), oracle)
    z = np.ascontiguousarray(rand(rng, 4, oracle))
    rem = np.zeros(4, np.uint32)
    d_p = dev(torch, p)
    ok(b200lib.supra_poly_divide(ptr(d_p), size, rem.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p)))
    q, r = oracle.poly_divide(p, z)
    assert np.array_equal(rem, r)
    assert np.array_equal(host(d_p).reshape(-1), q.reshape(-1))
    # quotient * (x - z) + remainder == p, checked at a random point through the oracle's Horner evaluation
    ok(b200lib.supra_poly_divide(ptr(d_p), size, rem.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p)))   # scratch reuse
