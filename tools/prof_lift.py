"""One lift at 2^18 (after a small segment) for an ncu launch list: which kernels make up the latency of a recursion proof."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boundless_b200 import ProverOpts, Segment, VerifierContext, get_prover_server
srv = get_prover_server(ProverOpts(segment_po2=12, recursion_po2=18, slots=1))
seg = srv.prove_segment(VerifierContext(), Segment(index=0, po2=12))
srv.lift(seg)
srv.close()
print("done")
