"""The oracle's small HAL operations (oracle/halops.c) against independent pure-Python field arithmetic and the
mathematical identities each operation must satisfy (no reference vectors exist for them: SURVEY.md 8c)."""
import numpy as np

P = 2013265921
NBETA = P - 11


def fp4_mul(a, b):
    return [(a[0] * b[0] + NBETA * (a[1] * b[3] + a[2] * b[2] + a[3] * b[1])) % P,
            (a[0] * b[1] + a[1] * b[0] + NBETA * (a[2] * b[3] + a[3] * b[2])) % P,
            (a[0] * b[2] + a[1] * b[1] + a[2] * b[0] + NBETA * a[3] * b[3]) % P,
            (a[0] * b[3] + a[1] * b[2] + a[2] * b[1] + a[3] * b[0]) % P]


def fp4_add(a, b):
    return [(x + y) % P for x, y in zip(a, b)]


def ints(a):
    return [[int(v) for v in row] for row in a]


def test_mix_poly_coeffs_definition(oracle):
    rng = np.random.default_rng(11)
    input_size, count, n_combos = 7, 5, 3
    inp = rng.integers(0, P, (input_size, count), dtype=np.int64)
    combos = np.array([0, 2, 1, 0, 2, 2, 1], np.uint32)
    out0 = rng.integers(0, P, (n_combos * count, 4), dtype=np.int64)
    mix_start = [int(v) for v in rng.integers(0, P, 4)]
    mix = [int(v) for v in rng.integers(0, P, 4)]
    got = oracle.from_mont(oracle.mix_poly_coeffs(oracle.to_mont(out0), oracle.to_mont(np.array(mix_start)), oracle.to_mont(np.array(mix)),
                                                  oracle.to_mont(inp.reshape(-1)), combos, input_size, count)).reshape(-1, 4)
    exp = ints(out0)
    for idx in range(count):
        cur = mix_start
        for i in range(input_size):
            o = int(combos[i]) * count + idx
            exp[o] = fp4_add(exp[o], [c * int(inp[i][idx]) % P for c in cur])
            cur = fp4_mul(cur, mix)
    assert ints(got) == exp


def test_eltwise_ops(oracle):
    rng = np.random.default_rng(12)
    count, to_add = 9, 4
    inp = rng.integers(0, P, (to_add * count, 4), dtype=np.int64)
    got = oracle.from_mont(oracle.eltwise_sum_extelem(oracle.to_mont(inp), count, to_add)).reshape(4, count)
    for idx in range(count):
        tot = [0, 0, 0, 0]
        for i in range(to_add):
            tot = fp4_add(tot, [int(v) for v in inp[i * count + idx]])
        assert [int(got[j][idx]) for j in range(4)] == tot
    a = rng.integers(0, P, 33, dtype=np.int64); b = rng.integers(0, P, 33, dtype=np.int64)
    a[0], b[0] = P - 1, P - 1
    assert np.array_equal(oracle.from_mont(oracle.eltwise_add_elem(oracle.to_mont(a), oracle.to_mont(b))), ((a + b) % P).astype(np.uint32))
    z = np.array([5, 0xFFFFFFFF, 0, P - 1, 0xFFFFFFFF], np.uint32)
    assert list(oracle.eltwise_zeroize_elem(z)) == [5, 0, 0, P - 1, 0]


def test_poly_divide_is_division_by_x_minus_z(oracle):
    """q(x) * (x - z) + r == p(x) coefficient-wise, and r == p(z)."""
    rng = np.random.default_rng(13)
    for size in (1, 2, 17, 64):
        p = rng.integers(0, P, (size, 4), dtype=np.int64)
        z = [int(v) for v in rng.integers(0, P, 4)]
        q, r = oracle.poly_divide(oracle.to_mont(p), oracle.to_mont(np.array(z)))
        q = ints(oracle.from_mont(q).reshape(size, 4)); r = [int(v) for v in oracle.from_mont(r)]
        assert q[size - 1] == [0, 0, 0, 0]            # degree drops by one
        negz = [(P - v) % P for v in z]
        for d in range(size):
            lhs = fp4_mul(q[d], negz)
            if d > 0:
                lhs = fp4_add(lhs, q[d - 1])
            if d == 0:
                lhs = fp4_add(lhs, r)
            assert lhs == [int(v) for v in p[d]]
        acc = [0, 0, 0, 0]
        for d in reversed(range(size)):
            acc = fp4_add(fp4_mul(acc, z), [int(v) for v in p[d]])
        assert acc == r


def test_prefix_products_and_gathers(oracle):
    rng = np.random.default_rng(14)
    x = rng.integers(0, P, (20, 4), dtype=np.int64)
    got = ints(oracle.from_mont(oracle.prefix_products(oracle.to_mont(x))).reshape(20, 4))
    cur = [1, 0, 0, 0]
    for i in range(20):
        cur = fp4_mul(cur, [int(v) for v in x[i]])
        assert got[i] == cur
    src = rng.integers(0, P, 6 * 8, dtype=np.int64).astype(np.uint32)
    assert list(oracle.gather_sample(src, 3, 6, 8)) == [int(src[g * 8 + 3]) for g in range(6)]
    into = np.zeros(16, np.uint32)
    index = np.array([1, 3, 3, 6], np.uint32)
    offsets = np.array([9, 0, 5, 7, 2, 11, 13], np.uint32)
    values = np.arange(100, 107, dtype=np.uint32)
    out = oracle.scatter(into, index, offsets, values)
    exp = np.zeros(16, np.uint32)
    for k in range(1, 6):
        exp[offsets[k]] = values[k]
    assert np.array_equal(out, exp)


def test_merkle_open_verifies_against_the_root(oracle):
    rng = np.random.default_rng(15)
    rows, cols, top = 64, 5, 8
    m = oracle.to_mont(rng.integers(0, P, rows * cols, dtype=np.int64))
    nodes = oracle.merkle_build(m, rows, cols).reshape(-1, 8)
    for idx in (0, 1, 37, 63):
        op = oracle.merkle_open(nodes.reshape(-1), m, rows, cols, top, idx)
        assert op.size == cols + 8 * 3                      # 64 -> 8: three sibling digests
        assert np.array_equal(op[:cols], m.reshape(cols, rows)[:, idx])
        cur = oracle.hash_elems(op[:cols])
        node = idx + rows
        for lvl in range(3):
            sib = op[cols + 8 * lvl: cols + 8 * lvl + 8]
            cur = oracle.hash_pair(sib, cur) if node & 1 else oracle.hash_pair(cur, sib)
            node >>= 1
        assert np.array_equal(cur, nodes[node]) and top <= node < 2 * top


def test_commit_group_is_the_composition(oracle):
    rng = np.random.default_rng(16)
    n, count = 6, 3
    a = oracle.to_mont(rng.integers(0, P, count << n, dtype=np.int64))
    co, ev, nodes = oracle.commit_group(a, n, count)
    ref_co = oracle.batch_zk_shift(oracle.batch_intt(a, n, count), n, count)
    assert np.array_equal(co, ref_co)
    ref_ev = oracle.batch_expand_ntt(ref_co, n, count, 2)
    assert np.array_equal(ev, ref_ev)
    assert np.array_equal(nodes, oracle.merkle_build(ref_ev, 1 << (n + 2), count))
