"""Bit-exact parity of every C-ABI kernel entry point against the CPU oracle (SURVEY.md 7 acceptance (iii)).

All calls go through libb200zkp.so's extern "C" surface with raw device pointers (torch only owns the memory).
Integer work: the bar is bit-exact equality.
"""
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


def rand_elems(rng, n, oracle):
    return oracle.to_mont(rng.integers(0, P, n, dtype=np.int64))


EDGE = ["zero", "pm1", "equal"]


def edge_input(kind, n, oracle):
    if kind == "zero":
        return np.zeros(n, np.uint32)
    if kind == "pm1":
        return oracle.to_mont(np.full(n, P - 1, np.int64))
    return oracle.to_mont(np.full(n, 123456789, np.int64))


@pytest.mark.parametrize("lg_n,count", [(1, 3), (4, 5), (10, 1), (10, 16), (11, 3), (12, 7), (13, 3), (14, 16), (15, 2),
                                        (12, 272), (16, 16), (18, 3), (19, 2), (20, 2), (21, 2)])
def test_batch_intt_and_shift(gpu, b200lib, oracle, lg_n, count):
    torch = gpu
    rng = np.random.default_rng(lg_n * 100 + count)
    a = rand_elems(rng, count << lg_n, oracle)
    d = dev(torch, a)
    ck(b200lib.b200_batch_intt(ptr(d), lg_n, count, None))
    torch.cuda.synchronize()
    ref = oracle.batch_intt(a, lg_n, count)
    assert np.array_equal(host(d), ref)
    ck(b200lib.b200_batch_zk_shift(ptr(d), lg_n, count, None))
    torch.cuda.synchronize()
    shifted = oracle.batch_zk_shift(ref, lg_n, count)
    assert np.array_equal(host(d), shifted)
    # the fused K1+K2 entry point gives the same coefficients in one pass
    d2 = dev(torch, a)
    ck(b200lib.b200_batch_intt_zk_shift(ptr(d2), lg_n, count, None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d2), shifted)


@pytest.mark.parametrize("lg_n,count", [(1, 2), (5, 3), (10, 4), (12, 5), (13, 2), (14, 3), (16, 4), (17, 2), (19, 2), (20, 2), (21, 1), (22, 1), (23, 1), (24, 1), (25, 1), (26, 1)])
def test_batch_ntt_roundtrip(gpu, b200lib, oracle, lg_n, count):
    torch = gpu
    rng = np.random.default_rng(7 + lg_n)
    a = rand_elems(rng, count << lg_n, oracle)
    d = dev(torch, a)
    ck(b200lib.b200_batch_ntt(ptr(d), lg_n, count, None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d), oracle.batch_ntt(a, lg_n, count))
    # size-independent property: iNTT(NTT(x)) == x, for every size including the largest
    ck(b200lib.b200_batch_intt(ptr(d), lg_n, count, None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d), a)
    # bit_reverse twice is the identity, once matches the oracle's permutation
    ck(b200lib.b200_batch_bit_reverse(ptr(d), lg_n, count, None))
    torch.cuda.synchronize()
    ref = np.concatenate([oracle.bit_reverse(a[c << lg_n:(c + 1) << lg_n], lg_n) for c in range(count)])
    assert np.array_equal(host(d), ref)


@pytest.mark.parametrize("lg_n,count", [(3, 2), (8, 3), (10, 16), (11, 4), (12, 5), (13, 3), (14, 4), (12, 272), (16, 8), (18, 4), (19, 2), (20, 3), (21, 2)])
def test_batch_expand_ntt(gpu, b200lib, oracle, lg_n, count):
    torch = gpu
    rng = np.random.default_rng(31 + lg_n)
    a = rand_elems(rng, count << lg_n, oracle)
    d_in = dev(torch, a)
    d_out = torch.zeros(count << (lg_n + 2), dtype=torch.int32, device="cuda")
    ck(b200lib.b200_batch_expand_ntt(ptr(d_out), ptr(d_in), lg_n, 2, count, None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d_out), oracle.batch_expand_ntt(a, lg_n, count, 2))


@pytest.mark.parametrize("kind", EDGE)
@pytest.mark.parametrize("lg_n", [10, 14, 20])
def test_ntt_edge_inputs(gpu, b200lib, oracle, kind, lg_n):
    torch = gpu
    count = 3
    a = edge_input(kind, count << lg_n, oracle)
    d = dev(torch, a)
    ck(b200lib.b200_batch_intt(ptr(d), lg_n, count, None))
    torch.cuda.synchronize()
    ref = oracle.batch_intt(a, lg_n, count)
    assert np.array_equal(host(d), ref)
    d_out = torch.zeros(count << (lg_n + 2), dtype=torch.int32, device="cuda")
    ck(b200lib.b200_batch_expand_ntt(ptr(d_out), ptr(d), lg_n, 2, count, None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d_out), oracle.batch_expand_ntt(ref, lg_n, count, 2))


def test_intt_ntt_identity_max_size(gpu, b200lib, oracle):
    """2^24 elements (BASELINE config 5 upper end): round trip only, the oracle is not run at this size."""
    torch = gpu
    lg_n = 24
    rng = np.random.default_rng(5)
    a = rand_elems(rng, 1 << lg_n, oracle)
    d = dev(torch, a)
    ck(b200lib.b200_batch_intt(ptr(d), lg_n, 1, None))
    ck(b200lib.b200_batch_ntt(ptr(d), lg_n, 1, None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d), a)


@pytest.mark.parametrize("rows,cols", [(1, 1), (33, 16), (256, 17), (1024, 0), (1000, 5), (4096, 32), (1 << 14, 208), (1 << 12, 64)])
def test_poseidon2_rows(gpu, b200lib, oracle, rows, cols):
    torch = gpu
    rng = np.random.default_rng(rows + cols)
    m = rand_elems(rng, max(rows * cols, 1), oracle)[: rows * cols]
    d_m = dev(torch, m if m.size else np.zeros(1, np.uint32))
    d_out = torch.zeros(rows * 8, dtype=torch.int32, device="cuda")
    ck(b200lib.b200_poseidon2_rows(ptr(d_out), ptr(d_m), rows, cols, None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d_out), oracle.hash_rows(m, rows, cols))


def test_poseidon2_kat_through_rows(gpu, b200lib, oracle):
    """The permutation KAT through the product path: a 16-column row [0..16) absorbs into a zero state."""
    torch = gpu
    row = oracle.to_mont(np.arange(16, dtype=np.int64))
    d_m = dev(torch, row)
    d_out = torch.zeros(8, dtype=torch.int32, device="cuda")
    ck(b200lib.b200_poseidon2_rows(ptr(d_out), ptr(d_m), 1, 16, None))
    torch.cuda.synchronize()
    st = np.zeros(24, np.uint32); st[:16] = row
    assert np.array_equal(host(d_out), oracle.p2_mix(st)[:8])


@pytest.mark.parametrize("n_out", [1, 2, 31, 1024, 5000])
def test_poseidon2_fold(gpu, b200lib, oracle, n_out):
    torch = gpu
    rng = np.random.default_rng(n_out)
    inp = rand_elems(rng, n_out * 16, oracle)
    d_in = dev(torch, inp)
    d_out = torch.zeros(n_out * 8, dtype=torch.int32, device="cuda")
    ck(b200lib.b200_poseidon2_fold(ptr(d_out), ptr(d_in), n_out, None))
    torch.cuda.synchronize()
    ref = np.concatenate([oracle.hash_pair(inp[16 * i:16 * i + 8], inp[16 * i + 8:16 * i + 16]) for i in range(n_out)])
    assert np.array_equal(host(d_out), ref)


@pytest.mark.parametrize("lg_rows,cols", [(1, 3), (5, 16), (10, 20), (12, 64), (14, 16)])
def test_merkle_tree(gpu, b200lib, oracle, lg_rows, cols):
    torch = gpu
    rows = 1 << lg_rows
    rng = np.random.default_rng(lg_rows)
    m = rand_elems(rng, rows * cols, oracle)
    d_m = dev(torch, m)
    d_nodes = torch.zeros(2 * rows * 8, dtype=torch.int32, device="cuda")
    ck(b200lib.b200_merkle_tree(ptr(d_nodes), ptr(d_m), lg_rows, cols, None))
    torch.cuda.synchronize()
    ref = oracle.merkle_build(m, rows, cols)
    assert np.array_equal(host(d_nodes)[8:], ref[8:])      # node 0 is unused


@pytest.mark.parametrize("lg_size", [4, 8, 12, 16, 20])
def test_fri_fold(gpu, b200lib, oracle, lg_size):
    torch = gpu
    size = 1 << lg_size
    rng = np.random.default_rng(lg_size)
    inp = rand_elems(rng, 4 * size, oracle)
    mix = rand_elems(rng, 4, oracle)
    d_in, d_mix = dev(torch, inp), dev(torch, mix)
    d_out = torch.zeros(4 * (size // 16), dtype=torch.int32, device="cuda")
    ck(b200lib.b200_fri_fold(ptr(d_out), ptr(d_in), size, ptr(d_mix), None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d_out), oracle.fri_fold(inp, size, mix))


@pytest.mark.parametrize("lg_n,count", [(3, 2), (10, 5), (12, 3), (13, 4), (16, 6)])
def test_batch_evaluate_any(gpu, b200lib, oracle, lg_n, count):
    torch = gpu
    rng = np.random.default_rng(lg_n)
    co = rand_elems(rng, count << lg_n, oracle)
    x = rand_elems(rng, 4, oracle)
    d_co, d_x = dev(torch, co), dev(torch, x)
    d_out = torch.zeros(count * 4, dtype=torch.int32, device="cuda")
    d_scr = torch.zeros(b200lib.b200_evaluate_scratch_words(lg_n, count), dtype=torch.int32, device="cuda")
    ck(b200lib.b200_batch_evaluate_any(ptr(d_out), ptr(d_co), lg_n, count, ptr(d_x), ptr(d_scr), None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d_out).reshape(count, 4), oracle.batch_evaluate_any(co, lg_n, count, x))


def test_empty_and_out_of_range_inputs(gpu, b200lib, oracle):
    """Empty batches are no-ops, out-of-range sizes are reported as errors (never a crash, never a CPU fallback)."""
    torch = gpu
    d = torch.zeros(64, dtype=torch.int32, device="cuda")
    for fn in (b200lib.b200_batch_intt, b200lib.b200_batch_ntt, b200lib.b200_batch_zk_shift, b200lib.b200_batch_bit_reverse):
        ck(fn(ptr(d), 4, 0, None))                       # count == 0
    ck(b200lib.b200_batch_expand_ntt(ptr(d), ptr(d), 4, 2, 0, None))
    ck(b200lib.b200_poseidon2_rows(ptr(d), ptr(d), 0, 16, None))        # rows == 0
    ck(b200lib.b200_poseidon2_fold(ptr(d), ptr(d), 0, None))
    ck(b200lib.b200_fri_fold(ptr(d), ptr(d), 8, ptr(d), None))          # fewer than 16 coefficients: nothing to fold
    torch.cuda.synchronize()
    assert int(d.abs().sum()) == 0
    assert b200lib.b200_batch_intt(ptr(d), 27, 1, None) is not None     # > 2^26
    assert b200lib.b200_batch_expand_ntt(ptr(d), ptr(d), 25, 2, 1, None) is not None
    assert b200lib.b200_merkle_tree(ptr(d), ptr(d), 27, 1, None) is not None
    # size-1 transforms are the identity
    one = dev(torch, oracle.to_mont(np.array([5, 7, 11])))
    ck(b200lib.b200_batch_intt(ptr(one), 0, 3, None)); ck(b200lib.b200_batch_ntt(ptr(one), 0, 3, None))
    torch.cuda.synchronize()
    assert np.array_equal(host(one), oracle.to_mont(np.array([5, 7, 11])))


def test_stream_argument(gpu, b200lib, oracle):
    """Entry points honour the caller's stream (the agent passes its own): run on a side stream."""
    torch = gpu
    s = torch.cuda.Stream()
    a = rand_elems(np.random.default_rng(1), 4 << 12, oracle)
    d = dev(torch, a)
    torch.cuda.synchronize()
    ck(b200lib.b200_batch_intt(ptr(d), 12, 4, C.c_void_p(s.cuda_stream)))
    s.synchronize()
    assert np.array_equal(host(d), oracle.batch_intt(a, 12, 4))


# ---- BASELINE-size cases (VERDICT r01 "Next round" item 1): the shapes the headline runs, bit-exact against the oracle ----------------
def dev_rand(torch, n, seed):
    """n canonical Montgomery words generated on the device (any value below p is a valid element); host copy for the oracle"""
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    t = torch.randint(0, P, (n,), dtype=torch.int32, device="cuda", generator=g)
    return t, t.cpu().numpy().view(np.uint32)


def test_poseidon2_rows_full_data_group(gpu, b200lib, oracle):
    """K4 at the headline shape: 2^22 rows x 208 columns (the data group of a 2^20-row segment, 13 permutations per row)."""
    torch = gpu
    rows, cols = 1 << 22, 208
    d_m, m = dev_rand(torch, rows * cols, 2208)
    d_out = torch.zeros(rows * 8, dtype=torch.int32, device="cuda")
    ck(b200lib.b200_poseidon2_rows(ptr(d_out), ptr(d_m), rows, cols, None))
    torch.cuda.synchronize()
    got = host(d_out)
    del d_m, d_out
    assert np.array_equal(got, oracle.hash_rows(m, rows, cols))


def test_merkle_tree_full_size(gpu, b200lib, oracle):
    """K4 + K5 at 2^22 leaves (every layer kernel of the headline tree: wide folds, the one-CTA top, the warp-form last six layers)."""
    torch = gpu
    lg_rows, cols = 22, 16
    rows = 1 << lg_rows
    d_m, m = dev_rand(torch, rows * cols, 2216)
    d_nodes = torch.zeros(2 * rows * 8, dtype=torch.int32, device="cuda")
    ck(b200lib.b200_merkle_tree(ptr(d_nodes), ptr(d_m), lg_rows, cols, None))
    torch.cuda.synchronize()
    ref = oracle.merkle_build(m, rows, cols)
    assert np.array_equal(host(d_nodes)[8:], ref[8:])


@pytest.mark.parametrize("lg_n,count", [(23, 2), (24, 1), (21, 3)])
def test_batch_expand_ntt_three_pass_sizes(gpu, b200lib, oracle, lg_n, count):
    """K3 for po2 23 / 24 segments: 2^23 -> 2^25 and 2^24 -> 2^26 evaluations (three-pass route: a complete expand + NTT per row of the
    outer split, the inter-pass twiddle as its own kernel, the strided pass), and (21 -> 23) as the largest two-pass neighbour."""
    torch = gpu
    d_in, a = dev_rand(torch, count << lg_n, 3000 + lg_n)
    d_out = torch.zeros(count << (lg_n + 2), dtype=torch.int32, device="cuda")
    ck(b200lib.b200_batch_expand_ntt(ptr(d_out), ptr(d_in), lg_n, 2, count, None))
    torch.cuda.synchronize()
    got = host(d_out)
    del d_in, d_out
    assert np.array_equal(got, oracle.batch_expand_ntt(a, lg_n, count, 2))


@pytest.mark.parametrize("count", [16, 272])
def test_batch_expand_ntt_full_size(gpu, b200lib, oracle, count):
    """K3 at the headline shape: 2^20 coefficients -> 2^22 evaluations, 16 columns (one launch group) and all 272 columns of a segment."""
    torch = gpu
    lg_n = 20
    d_in, a = dev_rand(torch, count << lg_n, 2000 + count)
    d_out = torch.zeros(count << (lg_n + 2), dtype=torch.int32, device="cuda")
    ck(b200lib.b200_batch_expand_ntt(ptr(d_out), ptr(d_in), lg_n, 2, count, None))
    torch.cuda.synchronize()
    got = host(d_out)
    del d_in, d_out
    assert np.array_equal(got, oracle.batch_expand_ntt(a, lg_n, count, 2))


@pytest.mark.parametrize("lg_n,count", [(20, 272), (22, 4), (23, 2), (24, 2), (25, 2), (26, 1)])
def test_batch_intt_large_against_the_oracle(gpu, b200lib, oracle, lg_n, count):
    """K1 (+K2 fused) against the oracle at the headline shape (272 x 2^20) and at lg_n 22-26, not only as a round trip: 2^22 is the
    check-polynomial iNTT of a 2^20 segment; 2^25 and 2^26 (the evaluation domains of po2 23 / 24, upstream MAX_CYCLES_PO2 = 24) take
    the three-pass route (outer strided pass + a complete inner transform per row)."""
    torch = gpu
    d, a = dev_rand(torch, count << lg_n, 100 * lg_n + count)
    d2 = d.clone()
    ck(b200lib.b200_batch_intt(ptr(d), lg_n, count, None))
    ck(b200lib.b200_batch_intt_zk_shift(ptr(d2), lg_n, count, None))
    torch.cuda.synchronize()
    ref = oracle.batch_intt(a, lg_n, count)
    assert np.array_equal(host(d), ref)
    assert np.array_equal(host(d2), oracle.batch_zk_shift(ref, lg_n, count))
    ck(b200lib.b200_batch_ntt(ptr(d), lg_n, count, None))
    torch.cuda.synchronize()
    assert np.array_equal(host(d), a)


# Every alternative route of the NTT launchers (read from the environment per launch) must give the oracle's words too: the radix-16 and
# cp.async-staged strided passes, the radix-16 inverse pass B, the inter-pass twiddle in pass A's epilogue, the two-table twiddle /
# zk_shift decomposition instead of the per-element tables, and a table-size cap below the transform size.
NTT_ROUTES = [
    {"B200_NTT_FULL": "0"},
    {"B200_NTT_FULL_MAX_LG": "12"},
    {"B200_NTT_INVB_R32": "0"},
    {"B200_NTT_INVB_R32": "0", "B200_NTT_FULL": "0"},
    {"B200_NTT_TW_IN_B": "0"},
    {"B200_NTT_R32_DIRECT": "0"},
    {"B200_NTT_R32_DIRECT": "0", "B200_NTT_TW_IN_B": "0", "B200_NTT_R32_MINB": "3"},
    {"B200_NTT_R32": "0"},
    {"B200_NTT_FUSED": "0"},
    {"B200_NTT_INVB_R32_MINB": "4", "B200_NTT_R32D_MINB": "5", "B200_NTT_FWD1_MINB": "5", "B200_NTT_INVB_R32_WAVES": "1"},
]


@pytest.mark.parametrize("route", NTT_ROUTES, ids=lambda r: ",".join("%s=%s" % kv for kv in r.items()))
def test_ntt_alternative_routes_bit_exact(gpu, b200lib, oracle, monkeypatch, route):
    torch = gpu
    for k, v in route.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(2024)
    for lg_n, count in [(20, 2), (18, 3), (15, 4)]:
        a = rand_elems(rng, count << lg_n, oracle)
        ref = oracle.batch_intt(a, lg_n, count)
        d = dev(torch, a)
        ck(b200lib.b200_batch_intt(ptr(d), lg_n, count, None))
        torch.cuda.synchronize()
        assert np.array_equal(host(d), ref), (route, "intt", lg_n)
        d = dev(torch, a)
        ck(b200lib.b200_batch_intt_zk_shift(ptr(d), lg_n, count, None))
        torch.cuda.synchronize()
        assert np.array_equal(host(d), oracle.batch_zk_shift(ref, lg_n, count)), (route, "intt+shift", lg_n)
        if lg_n <= 18 or count <= 2:
            d_out = torch.zeros(count << (lg_n + 2), dtype=torch.int32, device="cuda")
            ck(b200lib.b200_batch_expand_ntt(ptr(d_out), ptr(dev(torch, a)), lg_n, 2, count, None))
            torch.cuda.synchronize()
            assert np.array_equal(host(d_out), oracle.batch_expand_ntt(a, lg_n, count, 2)), (route, "expand", lg_n)
