"""Launch each hot kernel a few times at the BASELINE shapes (16 columns x 2^20 rows) for ncu captures.
usage: python tools/prof_kernels.py [what...]   what in {ntt, rows, fold}"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boundless_b200 import lib

L = lib.require_gpu(0)
what = sys.argv[1:] or ["ntt", "rows", "fold"]
n, cnt = 20, 16
P = 2013265921
p = lambda t: C.c_void_p(t.data_ptr())
if "ntt" in what:
    a = torch.randint(0, P, (cnt << n,), dtype=torch.int32, device="cuda")
    o = torch.empty(cnt << (n + 2), dtype=torch.int32, device="cuda")
    for _ in range(2):
        assert L.b200_batch_intt(p(a), n, cnt, None) is None
        assert L.b200_batch_zk_shift(p(a), n, cnt, None) is None
        assert L.b200_batch_expand_ntt(p(o), p(a), n, 2, cnt, None) is None
    torch.cuda.synchronize()
if "rows" in what:
    rows, cols = 1 << 22, 32
    m = torch.randint(0, P, (rows * cols,), dtype=torch.int32, device="cuda")
    d = torch.empty(rows * 8, dtype=torch.int32, device="cuda")
    for _ in range(2):
        assert L.b200_poseidon2_rows(p(d), p(m), rows, cols, None) is None
    torch.cuda.synchronize()
if "rows208" in what:      # the bench's roofline kernel: K4 on the data group (2^22 x 208)
    rows, cols = 1 << 22, 208
    m = torch.randint(0, P, (rows * cols,), dtype=torch.int32, device="cuda")
    d = torch.empty(rows * 8, dtype=torch.int32, device="cuda")
    for _ in range(2):
        assert L.b200_poseidon2_rows(p(d), p(m), rows, cols, None) is None
    torch.cuda.synchronize()
if "fold" in what:
    nodes = torch.randint(0, P, (2 * (1 << 22) * 8,), dtype=torch.int32, device="cuda")
    assert L.b200_poseidon2_fold(p(nodes), p(nodes[(1 << 22) * 8:]), 1 << 21, None) is None
    torch.cuda.synchronize()
print("done")
