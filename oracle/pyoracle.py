"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from boundless_b200/ (the product path).
PARITY UNPINNED at seal level (see oracle/bb.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

P = 2013265921
QUERIES = 50


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("poseidon2.c", "ntt.c", "stark.c", "halops.c", "bb.h", "oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


class Circuit(C.Structure):
    _fields_ = [("po2", C.c_uint32), ("w_code", C.c_uint32), ("w_data", C.c_uint32),
                ("w_accum", C.c_uint32), ("kind", C.c_uint32)]


class Fp4(C.Structure):
    _fields_ = [("c", C.c_uint32 * 4)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u32p = C.POINTER(C.c_uint32)
        L.oracle_p2_rc_canon.restype = u32p
        L.oracle_p2_diag_canon.restype = u32p
        L.oracle_seal_words.restype = C.c_size_t
        L.oracle_seal_words.argtypes = [C.POINTER(Circuit)]
        L.oracle_prove.restype = C.c_int
        L.oracle_prove.argtypes = [C.POINTER(Circuit), C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_verify.restype = C.c_int
        L.oracle_verify.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_selftest.restype = C.c_int
        L.oracle_rou_fwd.restype = C.c_uint32
        L.oracle_rou_rev.restype = C.c_uint32
        L.oracle_rou_fwd.argtypes = [C.c_uint]
        L.oracle_rou_rev.argtypes = [C.c_uint]
        for name, args in {
            "oracle_p2_mix": [C.c_void_p],
            "oracle_p2_hash_elems": [C.c_void_p, C.c_void_p, C.c_size_t],
            "oracle_p2_hash_pair": [C.c_void_p, C.c_void_p, C.c_void_p],
            "oracle_p2_hash_rows": [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t],
            "oracle_interpolate_ntt": [C.c_void_p, C.c_uint],
            "oracle_evaluate_ntt": [C.c_void_p, C.c_uint, C.c_uint],
            "oracle_bit_reverse": [C.c_void_p, C.c_uint],
            "oracle_batch_interpolate_ntt": [C.c_void_p, C.c_uint, C.c_size_t],
            "oracle_batch_zk_shift": [C.c_void_p, C.c_uint, C.c_size_t],
            "oracle_batch_evaluate_ntt": [C.c_void_p, C.c_uint, C.c_size_t],
            "oracle_batch_expand_into_evaluate_ntt": [C.c_void_p, C.c_void_p, C.c_uint, C.c_size_t, C.c_uint],
            "oracle_merkle_build": [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t],
            "oracle_fri_fold": [C.c_void_p, C.c_void_p, C.c_size_t, Fp4],
            "oracle_batch_evaluate_any": [C.c_void_p, C.c_void_p, C.c_uint, C.c_size_t, Fp4],
            "oracle_gen_trace": [C.c_void_p, C.c_uint64, C.c_uint, C.c_size_t],
            "oracle_segment_digest": [C.c_void_p, C.c_uint64],
            "oracle_seal_digest": [C.c_void_p, C.c_void_p, C.c_size_t],
            "oracle_mix_poly_coeffs": [C.c_void_p, Fp4, Fp4, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t],
            "oracle_eltwise_sum_extelem": [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t],
            "oracle_eltwise_add_elem": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t],
            "oracle_eltwise_copy_elem": [C.c_void_p, C.c_void_p, C.c_size_t],
            "oracle_eltwise_zeroize_elem": [C.c_void_p, C.c_size_t],
            "oracle_prefix_products": [C.c_void_p, C.c_size_t],
            "oracle_gather_sample": [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t],
            "oracle_scatter": [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p],
            "oracle_commit_group": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_size_t],
            "oracle_rng_init": [C.c_void_p],
            "oracle_rng_mix": [C.c_void_p, C.c_void_p],
        }.items():
            getattr(L, name).argtypes = args
            getattr(L, name).restype = None
        L.oracle_poly_divide.argtypes = [C.c_void_p, C.c_size_t, Fp4]
        L.oracle_poly_divide.restype = Fp4
        L.oracle_merkle_open.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]
        L.oracle_merkle_open.restype = C.c_size_t
        L.oracle_rng_elem.argtypes = [C.c_void_p]
        L.oracle_rng_elem.restype = C.c_uint32
        L.oracle_p2_init()
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a


# ---- field helpers (numpy, uint64 intermediates) ----
R = (1 << 32) % P
R2 = (R * R) % P
RINV = pow(R, -1, P)


def to_mont(x):
    return ((np.asarray(x, dtype=np.uint64) % P) * np.uint64(R) % np.uint64(P)).astype(np.uint32)


def from_mont(x):
    return (np.asarray(x, dtype=np.uint64) * np.uint64(RINV) % np.uint64(P)).astype(np.uint32)


def fp4(vals):
    f = Fp4()
    for i in range(4):
        f.c[i] = int(vals[i])
    return f


# ---- wrappers ----
def selftest():
    return lib().oracle_selftest()


def p2_mix(cells):
    a = _u32(cells).copy()
    lib().oracle_p2_mix(_p(a))
    return a


def hash_elems(elems):
    a = _u32(elems)
    out = np.zeros(8, np.uint32)
    lib().oracle_p2_hash_elems(_p(out), _p(a), a.size)
    return out


def hash_pair(a, b):
    out = np.zeros(8, np.uint32)
    a = _u32(a); b = _u32(b)
    lib().oracle_p2_hash_pair(_p(out), _p(a), _p(b))
    return out


def hash_rows(matrix, rows, cols):
    m = _u32(matrix)
    assert m.size == rows * cols
    out = np.zeros(rows * 8, np.uint32)
    lib().oracle_p2_hash_rows(_p(out), _p(m), rows, cols)
    return out


def merkle_build(matrix, rows, cols):
    m = _u32(matrix)
    nodes = np.zeros(2 * rows * 8, np.uint32)
    lib().oracle_merkle_build(_p(nodes), _p(m), rows, cols)
    return nodes


def batch_intt(io, n, count):
    a = _u32(io).copy()
    lib().oracle_batch_interpolate_ntt(_p(a), n, count)
    return a


def batch_zk_shift(io, n, count):
    a = _u32(io).copy()
    lib().oracle_batch_zk_shift(_p(a), n, count)
    return a


def batch_ntt(io, n, count):
    a = _u32(io).copy()
    lib().oracle_batch_evaluate_ntt(_p(a), n, count)
    return a


def batch_expand_ntt(inp, n_in, count, e=2):
    a = _u32(inp)
    out = np.zeros(count << (n_in + e), np.uint32)
    lib().oracle_batch_expand_into_evaluate_ntt(_p(out), _p(a), n_in, count, e)
    return out


def bit_reverse(io, n):
    a = _u32(io).copy()
    lib().oracle_bit_reverse(_p(a), n)
    return a


def fri_fold(inp, in_size, mix):
    a = _u32(inp)
    out = np.zeros(4 * (in_size // 16), np.uint32)
    lib().oracle_fri_fold(_p(out), _p(a), in_size, fp4(mix))
    return out


def batch_evaluate_any(coeffs, n, count, x):
    a = _u32(coeffs)
    out = np.zeros(count * 4, np.uint32)
    lib().oracle_batch_evaluate_any(_p(out), _p(a), n, count, fp4(x))
    return out.reshape(count, 4)


def mix_poly_coeffs(out, mix_start, mix, inp, combos, input_size, count):
    """out: (n_combos*count, 4) accumulated into (copy returned)."""
    o = _u32(out).copy(); a = _u32(inp); c = _u32(combos)
    lib().oracle_mix_poly_coeffs(_p(o), fp4(mix_start), fp4(mix), _p(a), _p(c), input_size, count)
    return o


def eltwise_sum_extelem(inp, count, to_add):
    a = _u32(inp)
    out = np.zeros(4 * count, np.uint32)
    lib().oracle_eltwise_sum_extelem(_p(out), _p(a), count, to_add)
    return out


def eltwise_add_elem(a, b):
    a = _u32(a); b = _u32(b)
    out = np.zeros(a.size, np.uint32)
    lib().oracle_eltwise_add_elem(_p(out), _p(a), _p(b), a.size)
    return out


def eltwise_zeroize_elem(io):
    a = _u32(io).copy()
    lib().oracle_eltwise_zeroize_elem(_p(a), a.size)
    return a


def poly_divide(poly, z):
    """poly: (size, 4) natural-order Fp4 coefficients.  Returns (quotient array, remainder[4])."""
    a = _u32(poly).copy()
    r = lib().oracle_poly_divide(_p(a), a.size // 4, fp4(z))
    return a, np.array(list(r.c), np.uint32)


def prefix_products(io):
    a = _u32(io).copy()
    lib().oracle_prefix_products(_p(a), a.size // 4)
    return a


def gather_sample(src, idx, size, stride):
    a = _u32(src)
    out = np.zeros(size, np.uint32)
    lib().oracle_gather_sample(_p(out), _p(a), idx, size, stride)
    return out


def scatter(into, index, offsets, values):
    o = _u32(into).copy(); i = _u32(index); f = _u32(offsets); v = _u32(values)
    lib().oracle_scatter(_p(o), _p(i), i.size, _p(f), _p(v))
    return o


def merkle_open(nodes, matrix, rows, cols, top_size, idx):
    n = _u32(nodes); m = _u32(matrix)
    out = np.zeros(cols + 8 * 32, np.uint32)
    w = lib().oracle_merkle_open(_p(out), _p(n), _p(m), rows, cols, top_size, idx)
    return out[:w].copy()


def commit_group(cols_evals, n, count):
    """Returns (coeffs, evals, nodes) of PolyGroup::new over `count` columns of 2^n evaluations."""
    co = _u32(cols_evals).copy()
    ev = np.zeros(count << (n + 2), np.uint32)
    nodes = np.zeros(2 * (1 << (n + 2)) * 8, np.uint32)
    lib().oracle_commit_group(_p(co), _p(ev), _p(nodes), n, count)
    return co, ev, nodes


def gen_trace(seed, po2, cols):
    out = np.zeros(cols << po2, np.uint32)
    lib().oracle_gen_trace(_p(out), seed, po2, cols)
    return out


def segment_digest(seed):
    out = np.zeros(8, np.uint32)
    lib().oracle_segment_digest(_p(out), seed)
    return out


def seal_digest(seal):
    s = _u32(seal)
    out = np.zeros(8, np.uint32)
    lib().oracle_seal_digest(_p(out), _p(s), s.size)
    return out


def seal_words(po2, w_code=16, w_data=208, w_accum=32, kind=0):
    c = Circuit(po2, w_code, w_data, w_accum, kind)
    return lib().oracle_seal_words(C.byref(c))


def prove(po2, seed, w_code=16, w_data=208, w_accum=32, kind=0, input_digest=None, trace=None):
    c = Circuit(po2, w_code, w_data, w_accum, kind)
    if input_digest is None:
        input_digest = segment_digest(seed)
    d = _u32(input_digest)
    seal = np.zeros(lib().oracle_seal_words(C.byref(c)), np.uint32)
    tr = None
    if trace is not None:
        tr = _u32(trace)
    rc = lib().oracle_prove(C.byref(c), seed, _p(d), _p(tr) if tr is not None else None, _p(seal))
    if rc != 0:
        raise RuntimeError("oracle_prove failed rc=%d" % rc)
    return seal


def verify(seal):
    s = _u32(seal)
    return lib().oracle_verify(_p(s), s.size)


class Rng:
    def __init__(self):
        self.buf = np.zeros(25, np.uint32)
        lib().oracle_rng_init(_p(self.buf))

    def mix(self, digest):
        d = _u32(digest)
        lib().oracle_rng_mix(_p(self.buf), _p(d))

    def elem(self):
        return lib().oracle_rng_elem(_p(self.buf))

    def ext(self):
        return [self.elem() for _ in range(4)]
