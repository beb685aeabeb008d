"""Launch each hot kernel a few times at the BASELINE shapes (16 columns x 2^20 rows) for ncu captures.
usage: python tools/prof_kernels.py [what...]   what in {ntt, rows, rows208, fold, tree, hal}
hal = K6 fri_fold, K7 batch_evaluate_any, K8 mix_poly_coeffs / eltwise_sum_extelem / poly_divide at the segment's sizes."""
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
if "ntt2" in what:     # round 2 final arrangement: fused iNTT + zk_shift (2 kernels), expand + NTT (2 kernels)
    a = torch.randint(0, P, (cnt << n,), dtype=torch.int32, device="cuda")
    o = torch.empty(cnt << (n + 2), dtype=torch.int32, device="cuda")
    for _ in range(3):
        assert L.b200_batch_intt_zk_shift(p(a), n, cnt, None) is None
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
if "tree" in what:         # a whole 2^20-leaf Merkle tree (wide folds, the one-CTA top, the warp-form last layers) + transcript kernels via a small proof
    lg_rows, cols = 20, 16
    m = torch.randint(0, P, ((1 << lg_rows) * cols,), dtype=torch.int32, device="cuda")
    nodes = torch.empty(2 * (1 << lg_rows) * 8, dtype=torch.int32, device="cuda")
    assert L.b200_merkle_tree(p(nodes), p(m), lg_rows, cols, None) is None
    torch.cuda.synchronize()
    from boundless_b200 import ProverOpts, Segment, VerifierContext, get_prover_server
    srv = get_prover_server(ProverOpts(segment_po2=12, recursion_po2=11, slots=1))
    srv.prove_segment(VerifierContext(), Segment(index=0, po2=12))
    srv.close()
if "hal" in what:
    N = 1 << 20
    rnd = lambda k: torch.randint(0, P, (k,), dtype=torch.int32, device="cuda")
    # K6: first FRI round, 4 planes x 2^20 -> 4 planes x 2^16
    fin, fout, mix = rnd(4 * N), rnd(4 * (N // 16)), rnd(4)
    assert L.b200_fri_fold(p(fout), p(fin), N, p(mix), None) is None
    # K7: 64 coefficient columns x 2^20 at one Fp4 point
    cols = 64
    co, x, ev = rnd(cols * N), rnd(4), rnd(cols * 4)
    scr = torch.empty(L.b200_evaluate_scratch_words(20, cols), dtype=torch.int32, device="cuda")
    assert L.b200_batch_evaluate_any(p(ev), p(co), 20, cols, p(x), p(scr), None) is None
    # K8a: 64 columns mixed into 3 combos of 2^20 ExtElems
    comb = torch.tensor([i % 3 for i in range(cols)], dtype=torch.int32, device="cuda")
    acc = torch.zeros(3 * N * 4, dtype=torch.int32, device="cuda")
    assert L.b200_mix_poly_coeffs(p(acc), p(mix), p(x), p(co), p(comb), cols, N, 3, None) is None
    # K8b: sum of the 3 combos -> 4 planes
    planes = torch.empty(4 * N, dtype=torch.int32, device="cuda")
    assert L.b200_eltwise_sum_extelem(p(planes), p(acc), N, 3, None) is None
    # K8c: synthetic division of a 2^20-coefficient ExtElem polynomial
    rem = torch.empty(4, dtype=torch.int32, device="cuda")
    dscr = torch.empty(L.b200_poly_divide_scratch_words(N), dtype=torch.int32, device="cuda")
    assert L.b200_poly_divide(p(acc), N, p(rem), p(x), p(dscr), None) is None
    torch.cuda.synchronize()
print("done")
