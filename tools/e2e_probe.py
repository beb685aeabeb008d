"""Where does the end-to-end arm lose against the device-resident arm?  Raw pinned H2D rate, and per-step wall times of the e2e loop."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundless_b200 import ProverOpts, Segment, get_prover_server, lib as b200lib
L = b200lib.require_gpu(0)
slots = int(sys.argv[1]) if len(sys.argv) > 1 else 4
srv = get_prover_server(ProverOpts(segment_po2=20, slots=slots))
c = srv.seg_circuit
tw = (c.w_code + c.w_data) << c.po2
p = C.c_void_p(); b200lib.check(L.b200_host_alloc(C.byref(p), tw * 4))
arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(tw,))
seed = 0xB2000000 + 500
b200lib.check(L.b200_witgen_to_host(srv.h, 0, C.byref(c), seed, p))
d = torch.empty(tw, dtype=torch.int32, device="cuda")
cudart = torch.cuda.cudart()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    torch.cuda.cudart().cudaMemcpyAsync(d.data_ptr(), p.value, tw * 4, 1, torch.cuda.current_stream().cuda_stream) if hasattr(cudart, "cudaMemcpyAsync") else d.copy_(torch.from_numpy(arr.view(np.int32)), non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("raw H2D %.1f MB in %.2f ms = %.1f GB/s" % (tw * 4 / 1e6, dt * 1e3, tw * 4 / dt / 1e9), flush=True)
def run(n, trace, stagger_ms=0.0):
    inflight, marks = [], []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        slot = i % slots
        if 0 < i < slots and stagger_ms:
            time.sleep(stagger_ms * 1e-3)
        if len(inflight) == slots:
            srv.wait(inflight.pop(0)); marks.append(time.perf_counter() - t0)
        srv.submit_segment(slot, Segment(index=i, po2=20, trace=trace, seed=seed if trace is not None else None))
        inflight.append(slot)
    for s in inflight:
        srv.wait(s); marks.append(time.perf_counter() - t0)
    return marks
for stagger in (0.0, 14.0, 20.0):
    for trace, name in ((None, "device-resident"), (arr, "host trace")):
        run(4, trace)
        for n in (12, 24):
            m = run(n, trace, stagger)
            print("stagger %4.1f ms  %-16s n=%2d completion (ms):" % (stagger, name, n), " ".join("%.0f" % (x * 1e3) for x in m[:12]), " => %.2f segments/s" % (n / m[-1]), flush=True)
srv.close()
