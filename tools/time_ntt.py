"""Time K1/K3 at the BASELINE shape (16 columns x 2^20) for a few launch configurations (env overrides in ntt.cu)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boundless_b200 import lib
L = lib.require_gpu(0)
n, cnt, P = 20, 16, 2013265921
a = torch.randint(0, P, (cnt << n,), dtype=torch.int32, device="cuda")
o = torch.empty(cnt << (n + 2), dtype=torch.int32, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for lgtw in (3, 4, 5):
    for th in (256, 512, 1024):
        os.environ["B200_NTT_LGTW"] = str(lgtw); os.environ["B200_NTT_THREADS"] = str(th)
        try:
            te = timeit(lambda: L.b200_batch_expand_ntt(p(o), p(a), n, 2, cnt, None))
            ti = timeit(lambda: L.b200_batch_intt(p(a), n, cnt, None))
            print("lgTW=%d threads=%4d  expand+ntt %.3f ms  intt %.3f ms" % (lgtw, th, te, ti), flush=True)
        except Exception as ex:
            print("lgTW=%d threads=%d failed %s" % (lgtw, th, ex))
