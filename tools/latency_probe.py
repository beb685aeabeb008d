"""Single-proof latencies on an otherwise idle GPU (one slot): 2^20 segment, lift and join at 2^18, device verify_integrity of each, and
the composite tasks.  These are what bound the join tree (BASELINE config 4), where proofs depend on each other.
    python tools/latency_probe.py            prints one JSON line;  B200_FOLD_WARP_MAX=0 reproduces round 1's tree tops."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boundless_b200 import ProverOpts, Segment, VerifierContext, get_prover_server
from boundless_b200.prover_server import KIND_JOIN

srv = get_prover_server(ProverOpts(segment_po2=20, recursion_po2=18, slots=1))
ctx = VerifierContext()
def wall(fn, reps=3):
    fn(); best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3, r
out = {"fold_warp_max": os.environ.get("B200_FOLD_WARP_MAX", "default")}
out["segment_ms"], seg = wall(lambda: srv.prove_segment(ctx, Segment(index=1)))
out["segment_device_ms"] = srv.last_ms(0)
out["lift_ms"], l0 = wall(lambda: srv.lift(seg))
out["lift_device_ms"] = srv.last_ms(0)
l1 = srv.lift(srv.prove_segment(ctx, Segment(index=2)))
out["join_ms"], j = wall(lambda: srv.join(l0, l1))
out["verify_segment_ms"], _ = wall(lambda: srv.verify_integrity(seg))
out["verify_join_ms"], _ = wall(lambda: srv.verify_integrity(j))
words = srv.seal_words(srv._rec_circuit(KIND_JOIN))
bufs = [torch.zeros(words, dtype=torch.int32, device="cuda") for _ in range(3)]
def prove_lift(i, b, verify):
    srv.submit_prove_lift(0, Segment(index=i), d_out=b.data_ptr(), verify=verify, host_seals=False)
    r = srv.wait_task(0)[1]; r.owner = b; return r
out["prove_lift_verified_ms"], a = wall(lambda: prove_lift(1, bufs[0], True))
out["prove_lift_unverified_ms"], _ = wall(lambda: prove_lift(1, bufs[0], False))
b = prove_lift(2, bufs[1], True)
def join_dev(verify):
    srv.submit_recursion_dev(0, KIND_JOIN, a, b, d_out=bufs[2].data_ptr(), verify=verify, host_seal=False)
    return srv.wait_task(0)
out["join_dev_verified_ms"], _ = wall(lambda: join_dev(True))
out["join_dev_unverified_ms"], _ = wall(lambda: join_dev(False))
srv.close()
print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in out.items()}))
