"""A/B the NTT kernel families at the segment shapes (env switches are read per launch): K1 iNTT and K3 expand+NTT of
16 x 2^20, plus the 2^22 iNTT of the check group.  usage: python tools/time_ntt2.py [ENV=VAL,ENV=VAL ...]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boundless_b200 import lib
L = lib.require_gpu(0)
n, cnt, P = 20, 16, 2013265921
a0 = torch.randint(0, P, (cnt << n,), dtype=torch.int32, device="cuda")
a = a0.clone()
o = torch.empty(cnt << (n + 2), dtype=torch.int32, device="cuda")
big0 = torch.randint(0, P, (4 << 22,), dtype=torch.int32, device="cuda")
big = big0.clone()
p = lambda t: C.c_void_p(t.data_ptr())
def timeit(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
configs = sys.argv[1:] or ["B200_NTT_R32=0", "B200_NTT_R32=1,B200_NTT_R32_MINB=2", "B200_NTT_R32=1,B200_NTT_R32_MINB=3"]
ref = {}
for cfg in configs:
    for kv in cfg.split(","):
        k, v = kv.split("="); os.environ[k] = v
    te = timeit(lambda: L.b200_batch_expand_ntt(p(o), p(a0), n, 2, cnt, None))
    ti = timeit(lambda: L.b200_batch_intt_zk_shift(p(a), n, cnt, None))
    tb = timeit(lambda: L.b200_batch_intt(p(big), 22, 4, None))
    # bit-exact agreement between configurations (the first one is the reference)
    a.copy_(a0); L.b200_batch_intt_zk_shift(p(a), n, cnt, None); big.copy_(big0); L.b200_batch_intt(p(big), 22, 4, None)
    L.b200_batch_expand_ntt(p(o), p(a0), n, 2, cnt, None); torch.cuda.synchronize()
    sig = (int(a.to(torch.int64).sum()), int(o.to(torch.int64).sum()), int(big.to(torch.int64).sum()))
    if not ref: ref["a"], ref["o"], ref["b"] = a.clone(), o.clone(), big.clone()
    same = bool(torch.equal(a, ref["a"]) and torch.equal(o, ref["o"]) and torch.equal(big, ref["b"]))
    print("%-44s expand+ntt %.4f ms (%.0f GB/s)  intt+shift %.4f ms (%.0f GB/s)  intt 4x2^22 %.4f ms  same=%s" % (
        cfg, te, 20.0 * cnt * (1 << n) / te * 1e-6, ti, 8.0 * cnt * (1 << n) / ti * 1e-6, tb, same), flush=True)
