"""Can a multiplier-bound kernel (K4 leaf hashing) and a latency-bound one (K3 expand+NTT) overlap on one GPU?  Times each alone and both
at once on two streams (optionally with the NTT stream at high priority).  If the concurrent time is close to the sum, the hardware's
CTA scheduler serialises them (K4's CTAs hold every register) and only a designed co-residency could recover the NTT's idle pipe cycles.
    python tools/overlap_probe.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boundless_b200 import lib
L = lib.require_gpu(0)
P = 2013265921
rows, cols = 1 << 22, 208
m = torch.randint(0, P, (rows * cols,), dtype=torch.int32, device="cuda")
d = torch.empty(rows * 8, dtype=torch.int32, device="cuda")
n, cnt = 20, 16
a = torch.randint(0, P, (cnt << n,), dtype=torch.int32, device="cuda")
o = torch.empty(cnt << (n + 2), dtype=torch.int32, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
REPS_NTT = 60
def k4(s): assert L.b200_poseidon2_rows(p(d), p(m), rows, cols, C.c_void_p(s.cuda_stream)) is None
def k3(s):
    for _ in range(REPS_NTT):
        assert L.b200_batch_expand_ntt(p(o), p(a), n, 2, cnt, C.c_void_p(s.cuda_stream)) is None
def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for prio in (0, -1):
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(priority=prio)
    k4(s1); k3(s2); torch.cuda.synchronize()
    t4 = timed(lambda: k4(s1)); t3 = timed(lambda: k3(s2))
    both = timed(lambda: (k3(s2), k4(s1)))
    both2 = timed(lambda: (k4(s1), k3(s2)))
    print("ntt stream priority %d: K4 alone %.2f ms, %d x K3 alone %.2f ms, sum %.2f | concurrent (NTT enqueued first) %.2f ms, (K4 first) %.2f ms"
          % (prio, t4, REPS_NTT, t3, t4 + t3, both, both2), flush=True)
