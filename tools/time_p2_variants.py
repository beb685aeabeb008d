"""Time build variants of the Poseidon2 leaf-hash kernel (K4) and the fold kernel (K5) at the headline shapes: 2^22 rows x 208 columns
and 2^21 parents.  One subprocess per (library, launch shape): B200_P2_CFG is read once per process.
    python tools/time_p2_variants.py [cfgs, default "0,1"]        (after tools/build_p2_variants.sh)"""
import ctypes as C, glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2 and sys.argv[1] == "child":
    import torch
    L = C.CDLL(sys.argv[2])
    for f in (L.p2_rows, L.p2_fold):
        f.restype = C.c_char_p
    L.p2_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    L.p2_fold.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    rows, cols, P = 1 << 22, 208, 2013265921
    m = torch.randint(0, P, (rows * cols,), dtype=torch.int32, device="cuda")
    d = torch.empty(rows * 8, dtype=torch.int32, device="cuda")
    def t(fn, reps=3):
        fn(); fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); [fn() for _ in range(reps)]; e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ms = t(lambda: L.p2_rows(d.data_ptr(), m.data_ptr(), rows, cols, None))
    n_out = 1 << 21
    ms_f = t(lambda: L.p2_fold(d.data_ptr(), m.data_ptr(), n_out, None), 5)
    chk = int(d[:4096].to(torch.int64).sum())
    print("%-16s cfg %s: rows %.3f ms %.3f Gperm/s | fold %.3f ms %.3f Gperm/s | chk %d" % (
        os.path.basename(sys.argv[2])[6:-3], os.environ.get("B200_P2_CFG", "0"), ms, rows * 13 / ms * 1e-6, ms_f, n_out / ms_f * 1e-6, chk), flush=True)
else:
    cfgs = (sys.argv[1] if len(sys.argv) > 1 else "0,1").split(",")
    for so in sorted(glob.glob(os.path.join(ROOT, "build", "p2var", "libp2_*.so"))):
        for cfg in cfgs:
            subprocess.run([sys.executable, __file__, "child", so], env=dict(os.environ, B200_P2_CFG=cfg))
