"""Bit-exact parity of the small HAL operations (K8, K9, prefix_products, element-wise helpers, commit_group) against the CPU
oracle (oracle/halops.c), through the C ABI with raw device pointers."""
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


def ck(err):
    assert err is None, err


def rand(rng, shape, oracle):
    return oracle.to_mont(rng.integers(0, P, shape, dtype=np.int64))


@pytest.mark.parametrize("input_size,count,n_combos", [(1, 1, 1), (7, 5, 3), (33, 1000, 4), (272, 4096, 3), (300, 70001, 5),
                                                      (3000, 257, 2)])
def test_mix_poly_coeffs(gpu, b200lib, oracle, input_size, count, n_combos):
    torch = gpu
    rng = np.random.default_rng(input_size * 7 + count)
    inp = rand(rng, input_size * count, oracle)
    combos = rng.integers(0, n_combos, input_size).astype(np.uint32)
    out0 = rand(rng, (n_combos * count, 4), oracle)
    ms, mx = rand(rng, 4, oracle), rand(rng, 4, oracle)
    d_out, d_in, d_cb, d_ms, d_mx = dev(torch, out0), dev(torch, inp), dev(torch, combos), dev(torch, ms), dev(torch, mx)
    ck(b200lib.b200_mix_poly_coeffs(ptr(d_out), ptr(d_ms), ptr(d_mx), ptr(d_in), ptr(d_cb), input_size, count, n_combos, None))
    torch.cuda.synchronize()
    ref = oracle.mix_poly_coeffs(out0, ms, mx, inp, combos, input_size, count)
    assert np.array_equal(host(d_out).reshape(-1), ref.reshape(-1))


def test_mix_poly_coeffs_rejects_oversize(gpu, b200lib):
    torch = gpu
    z = torch.zeros(64, dtype=torch.int32, device="cuda")
    err = b200lib.b200_mix_poly_coeffs(ptr(z), ptr(z), ptr(z), ptr(z), ptr(z), 20000, 1, 1, None)
    assert err is not None and b"input_size" in err


@pytest.mark.parametrize("count,to_add", [(1, 1), (9, 4), (5000, 3), (1 << 16, 3)])
def test_eltwise_sum_extelem(gpu, b200lib, oracle, count, to_add):
    torch = gpu
    rng = np.random.default_rng(count + to_add)
    inp = rand(rng, (to_add * count, 4), oracle)
    d_in = dev(torch, inp)
    d_out = torch.zeros(4 * count, dtype=torch.int32, device="cuda")
    ck(b200lib.b200_eltwise_sum_extelem(ptr(d_out), ptr(d_in), count, to_add, None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d_out), oracle.eltwise_sum_extelem(inp, count, to_add))


def test_eltwise_add_copy_zeroize(gpu, b200lib, oracle):
    torch = gpu
    rng = np.random.default_rng(5)
    n = 100003
    a, b = rand(rng, n, oracle), rand(rng, n, oracle)
    a[:3] = oracle.to_mont(np.array([P - 1, 0, P - 1])); b[:3] = oracle.to_mont(np.array([P - 1, 0, 1]))
    d_a, d_b = dev(torch, a), dev(torch, b)
    d_o = torch.zeros(n, dtype=torch.int32, device="cuda")
    ck(b200lib.b200_eltwise_add_elem(ptr(d_o), ptr(d_a), ptr(d_b), n, None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d_o), oracle.eltwise_add_elem(a, b))
    ck(b200lib.b200_eltwise_copy_elem(ptr(d_o), ptr(d_a), n, None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d_o), a)
    z = a.copy(); z[::7] = 0xFFFFFFFF
    d_z = dev(torch, z)
    ck(b200lib.b200_eltwise_zeroize_elem(ptr(d_z), n, None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d_z), oracle.eltwise_zeroize_elem(z))


@pytest.mark.parametrize("size", [1, 2, 7, 8, 9, 2047, 2048, 2049, 5000, 1 << 16, (1 << 18) + 13])
def test_poly_divide(gpu, b200lib, oracle, size):
    torch = gpu
    rng = np.random.default_rng(size)
    p = rand(rng, (size, 4), oracle)
    z = rand(rng, 4, oracle)
    d_p, d_z = dev(torch, p), dev(torch, z)
    d_r = torch.zeros(4, dtype=torch.int32, device="cuda")
    d_s = torch.zeros(b200lib.b200_poly_divide_scratch_words(size), dtype=torch.int32, device="cuda")
    ck(b200lib.b200_poly_divide(ptr(d_p), size, ptr(d_r), ptr(d_z), ptr(d_s), None))
    torch.cuda.synchronize()
    q, r = oracle.poly_divide(p, z)
    assert np.array_equal(host(d_r), r)
    assert np.array_equal(host(d_p).reshape(-1), q.reshape(-1))


@pytest.mark.parametrize("count", [1, 2, 8, 9, 2047, 2048, 2049, 70000, (1 << 20) + 5])
def test_prefix_products(gpu, b200lib, oracle, count):
    torch = gpu
    rng = np.random.default_rng(count)
    x = rand(rng, (count, 4), oracle)
    if count > 5:
        x[3] = oracle.to_mont(np.array([1, 0, 0, 0]))
    d_x = dev(torch, x)
    d_s = torch.zeros(b200lib.b200_prefix_products_scratch_words(count), dtype=torch.int32, device="cuda")
    ck(b200lib.b200_prefix_products(ptr(d_x), count, ptr(d_s), None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d_x).reshape(-1), oracle.prefix_products(x).reshape(-1))


def test_gather_sample_and_scatter(gpu, b200lib, oracle):
    torch = gpu
    rng = np.random.default_rng(8)
    size, stride = 272, 4096
    src = rand(rng, size * stride, oracle)
    d_src = dev(torch, src)
    d_dst = torch.zeros(size, dtype=torch.int32, device="cuda")
    for idx in (0, 1234, stride - 1):
        ck(b200lib.b200_gather_sample(ptr(d_dst), ptr(d_src), idx, size, stride, None))
        torch.cuda.synchronize()
        assert np.array_equal(host(d_dst), oracle.gather_sample(src, idx, size, stride))
    n_rows, per = 1000, 5
    index = (np.arange(n_rows + 1, dtype=np.uint32) * per + 3).astype(np.uint32)
    n_values = int(index[-1]) + 4
    offsets = rng.permutation(n_values * 2)[:n_values].astype(np.uint32)
    values = rand(rng, n_values, oracle)
    into = rand(rng, n_values * 2, oracle)
    d_into, d_index, d_off, d_val = dev(torch, into), dev(torch, index), dev(torch, offsets), dev(torch, values)
    ck(b200lib.b200_scatter(ptr(d_into), ptr(d_index), index.size, ptr(d_off), ptr(d_val), n_values, None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d_into), oracle.scatter(into, index, offsets, values))


@pytest.mark.parametrize("lg_rows,cols,top", [(5, 3, 32), (6, 5, 8), (12, 16, 32), (14, 64, 32)])
def test_merkle_open(gpu, b200lib, oracle, lg_rows, cols, top):
    torch = gpu
    rng = np.random.default_rng(lg_rows * 31 + cols)
    rows = 1 << lg_rows
    m = rand(rng, rows * cols, oracle)
    d_m = dev(torch, m)
    d_nodes = torch.zeros(2 * rows * 8, dtype=torch.int32, device="cuda")
    ck(b200lib.b200_merkle_tree(ptr(d_nodes), ptr(d_m), lg_rows, cols, None))
    nodes = oracle.merkle_build(m, rows, cols)
    words = b200lib.b200_merkle_open_words(lg_rows, cols, top)
    d_out = torch.zeros(max(words, 1), dtype=torch.int32, device="cuda")
    for idx in (0, 1, rows // 3, rows - 1):
        ck(b200lib.b200_merkle_open(ptr(d_out), ptr(d_nodes), ptr(d_m), lg_rows, cols, top, idx, None))
        torch.cuda.synchronize()
        ref = oracle.merkle_open(nodes, m, rows, cols, top, idx)
        assert ref.size == words
        assert np.array_equal(host(d_out)[:words], ref)
    assert b200lib.b200_merkle_open(ptr(d_out), ptr(d_nodes), ptr(d_m), lg_rows, cols, top, rows, None) is not None
    assert b200lib.b200_merkle_open(ptr(d_out), ptr(d_nodes), ptr(d_m), lg_rows, cols, 3, 0, None) is not None


@pytest.mark.parametrize("lg_n,count", [(6, 3), (10, 16), (13, 5), (16, 4)])
def test_commit_group(gpu, b200lib, oracle, lg_n, count):
    torch = gpu
    rng = np.random.default_rng(lg_n + count)
    a = rand(rng, count << lg_n, oracle)
    d_c = dev(torch, a)
    d_e = torch.zeros(count << (lg_n + 2), dtype=torch.int32, device="cuda")
    d_n = torch.zeros(2 * (1 << (lg_n + 2)) * 8, dtype=torch.int32, device="cuda")
    ck(b200lib.b200_commit_group(ptr(d_c), ptr(d_e), ptr(d_n), lg_n, count, None))
    torch.cuda.synchronize()
    co, ev, nodes = oracle.commit_group(a, lg_n, count)
    assert np.array_equal(host(d_c), co)
    assert np.array_equal(host(d_e), ev)
    assert np.array_equal(host(d_n)[8:], nodes[8:])        # node 0 is unused


def test_commit_group_headline_size(gpu, b200lib, oracle):
    """commit_group at N = 2^20: iNTT + zk_shift, expand + NTT to 2^22, 2^22 leaves, the whole tree -- every
    word of coefficients, evaluations and nodes against the oracle."""
    torch = gpu
    lg_n, count = 20, 16
    g = torch.Generator(device="cuda"); g.manual_seed(2020)
    d_c = torch.randint(0, 2013265921, (count << lg_n,), dtype=torch.int32, device="cuda", generator=g)
    a = host(d_c).copy()
    d_e = torch.zeros(count << (lg_n + 2), dtype=torch.int32, device="cuda")
    d_n = torch.zeros(2 * (1 << (lg_n + 2)) * 8, dtype=torch.int32, device="cuda")
    ck(b200lib.b200_commit_group(ptr(d_c), ptr(d_e), ptr(d_n), lg_n, count, None))
    torch.cuda.synchronize()
    co, ev, nodes = oracle.commit_group(a, lg_n, count)
    assert np.array_equal(host(d_c), co)
    assert np.array_equal(host(d_e), ev)
    assert np.array_equal(host(d_n)[8:], nodes[8:])


def test_full_size_deep_pipeline_properties(gpu, b200lib, oracle):
    """BASELINE size (N = 2^20): mix 272 columns into 3 combos, divide each by (x - z), sum: size-independent checks --
    the remainder equals the combo evaluated by b200_batch_evaluate_any-style Horner on a strided sample is too slow on the
    CPU, so use linearity: dividing (A + B) equals dividing A plus dividing B, bit for bit."""
    torch = gpu
    rng = np.random.default_rng(99)
    n = 1 << 20
    a, b = rand(rng, (n, 4), oracle), rand(rng, (n, 4), oracle)
    z = rand(rng, 4, oracle)
    d_z = dev(torch, z)
    d_a, d_b = dev(torch, a), dev(torch, b)
    d_s = torch.zeros(8 * n, dtype=torch.int32, device="cuda")
    ck(b200lib.b200_eltwise_add_elem(ptr(d_s), ptr(d_a), ptr(d_b), 4 * n, None))
    d_sum = d_s[:4 * n].clone()
    scratch = torch.zeros(b200lib.b200_poly_divide_scratch_words(n), dtype=torch.int32, device="cuda")
    rem = torch.zeros(12, dtype=torch.int32, device="cuda")
    for k, t in enumerate((d_a, d_b, d_sum)):
        ck(b200lib.b200_poly_divide(ptr(t), n, C.c_void_p(rem.data_ptr() + 16 * k), ptr(d_z), ptr(scratch), None))
    d_q = torch.zeros(4 * n, dtype=torch.int32, device="cuda")
    ck(b200lib.b200_eltwise_add_elem(ptr(d_q), ptr(d_a), ptr(d_b), 4 * n, None))
    torch.cuda.synchronize()
    assert torch.equal(d_q, d_sum)
    r = host(rem).reshape(3, 4)
    assert np.array_equal(oracle.eltwise_add_elem(r[0], r[1]), r[2])
    # and the quotient's top coefficient is zero while the next one is the original leading coefficient
    q = host(d_a).reshape(n, 4)
    assert not q[n - 1].any() and np.array_equal(q[n - 2], a[n - 1])
