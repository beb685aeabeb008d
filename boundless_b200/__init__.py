"""boundless_b200: B200-native (sm_100a) kernels + prover pipeline for the Boundless/bento segment-proving path.

The product is libb200zkp.so (C ABI in include/b200zkp.h).  This package is the thin host-side mirror of the
reference's operator interface; nothing here (or in the library) falls back to the CPU.
"""
from .lib import B200Error, Circuit, load, require_gpu  # noqa: F401
from .planner import Planner, PlannerErr  # noqa: F401
from .prover_server import (ProverOpts, ProverServer, Segment, SegmentReceipt, SuccinctReceipt, VerificationError,  # noqa: F401
                            VerifierContext, get_prover_server)
