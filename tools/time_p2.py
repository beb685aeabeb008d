"""A/B the launch shape of the Poseidon2 leaf kernel (env B200_P2_CFG is read once per process -> one subprocess per config)."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from boundless_b200 import lib
    L = lib.require_gpu(0)
    rows, cols, P = 1 << 22, 32, 2013265921
    m = torch.randint(0, P, (rows * cols,), dtype=torch.int32, device="cuda")
    d = torch.empty(rows * 8, dtype=torch.int32, device="cuda")
    f = lambda: L.b200_poseidon2_rows(C.c_void_p(d.data_ptr()), C.c_void_p(m.data_ptr()), rows, cols, None)
    f(); f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [f() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("cfg %s: %.3f ms  %.3f Gperm/s" % (os.environ.get("B200_P2_CFG", "0"), ms, rows * 2 / ms * 1e-6), flush=True)
else:
    for cfg in "01234":
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, B200_P2_CFG=cfg))
