"""BASELINE config 5: BabyBear NTT size sweep 2^18..2^24 elements per polynomial, forward (bit-reversed coefficients -> evaluations),
inverse, and the prover's expand x4 + NTT, batched; achieved algorithmic GB/s against the measured HBM peak and Gmulmod/s against the
measured integer roofline.  Writes one JSON line per size to stdout.   python tools/ntt_sweep.py > profiles/ntt_sweep_r01.jsonl"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from boundless_b200 import lib
L = lib.require_gpu(0)
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]; SRC = "measured"
except Exception:
    PEAK, SRC = 6650.0, "fallback"
INT_PEAK = 3508.0
P = 2013265921
p = lambda t: C.c_void_p(t.data_ptr())
def timeit(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for lg in range(18, 25):
    for count in (1, 16, 256):
        if (count << lg) * 4 > (8 << 30):        # keep the sweep under 8 GiB per buffer (and > L2 for count >= 16)
            continue
        a = torch.randint(0, P, (count << lg,), dtype=torch.int32, device="cuda")
        n = 1 << lg
        t_inv = timeit(lambda: L.b200_batch_intt(p(a), lg, count, None))
        t_fwd = timeit(lambda: L.b200_batch_ntt(p(a), lg, count, None))
        rec = {"lg_n": lg, "count": count, "bytes_moved_algorithmic": 8 * n * count,
               "intt_ms": t_inv, "intt_gbs": 8.0 * n * count / t_inv * 1e-6, "intt_frac_hbm": 8.0 * n * count / t_inv * 1e-6 / PEAK,
               "ntt_ms": t_fwd, "ntt_gbs": 8.0 * n * count / t_fwd * 1e-6, "ntt_frac_hbm": 8.0 * n * count / t_fwd * 1e-6 / PEAK,
               "butterfly_gmulmod_s_inv": 0.5 * lg * n * count / t_inv * 1e-6, "int_roofline_gmulmod_s": INT_PEAK, "hbm_peak_gbs": PEAK, "peak_source": SRC}
        if lg <= 22 and (count << (lg + 2)) * 4 <= (8 << 30):
            o = torch.empty(count << (lg + 2), dtype=torch.int32, device="cuda")
            t_e = timeit(lambda: L.b200_batch_expand_ntt(p(o), p(a), lg, 2, count, None))
            rec.update({"expand_ntt_ms": t_e, "expand_ntt_gbs": 20.0 * n * count / t_e * 1e-6, "expand_ntt_frac_hbm": 20.0 * n * count / t_e * 1e-6 / PEAK})
            del o
        print(json.dumps(rec), flush=True)
        del a
