"""Latency of the warp-form Poseidon2 (P2Warp): a Merkle tree over 2^12 rows x 16 columns is twelve dependent narrow layers, and
k_hash_elems over 4096 elements is a chain of 256 permutations in one warp.  usage: python tools/time_p2w.py"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boundless_b200 import lib
L = lib.require_gpu(0)
P = 2013265921
p = lambda t: C.c_void_p(t.data_ptr())
def timeit(fn, reps=20):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
out = {}
for lg in (6, 12, 16):
    m = torch.randint(0, P, (16 << lg,), dtype=torch.int32, device="cuda")
    nodes = torch.empty(16 << lg, dtype=torch.int32, device="cuda")
    out["merkle_tree_2^%d_x16_us" % lg] = round(timeit(lambda: L.b200_merkle_tree(p(nodes), p(m), lg, 16, None)), 2)
print(json.dumps(out))
