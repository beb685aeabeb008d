"""Host-side mirror of the reference's operator boundary for the segment-proving path.

The bento GPU agent holds `Rc<dyn risc0_zkvm::ProverServer>` obtained from
`get_prover_server(&ProverOpts::default())` (/root/reference/prover/crates/workflow/src/lib.rs:276-284) and calls
`prove_segment` (tasks/prove.rs:44-52), `lift` (:96-104), `join` (tasks/join.rs:52-56), `resolve`
(tasks/resolve.rs:84-88) and `union` (tasks/union.rs:43-47) on it.  This module keeps those names and argument
meanings over the C ABI of libb200zkp.so; errors surface as exceptions the way the reference surfaces
`anyhow::Error` (the agent turns any Err into a task retry, workflow/src/lib.rs:639-677).

Synthetic circuit: the rv32im witgen / eval_check are out of scope, so a `Segment` is
(index, po2, seed[, host trace]) and receipts carry the seal words plus the claim metadata needed by lift/join.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import lib as _lib
from .lib import B200Error, Circuit

KIND_SEGMENT, KIND_LIFT, KIND_JOIN, KIND_RESOLVE, KIND_UNION = range(5)
# proof-of-verifiable-work variants of the recursion programs (tasks/prove.rs:70-78 lift_povw, tasks/join_povw.rs:55 join_povw,
# tasks/resolve_povw.rs:57 unwrap_povw): receipts whose claim is WorkClaim<ReceiptClaim>; on the synthetic path they are further
# values of the circuit header's `kind` word, so a PoVW receipt can never be passed off as a plain one (verify_integrity binds the kind)
KIND_LIFT_POVW, KIND_JOIN_POVW, KIND_UNWRAP_POVW = 5, 6, 7
POVW_KINDS = (KIND_LIFT_POVW, KIND_JOIN_POVW)
KIND_KECCAK = 8        # the keccak coprocessor's proof (tasks/keccak.rs:71-75 prove_keccak); its receipts feed the union tree
KECCAK_STATE_BYTES = 200   # [u64; 25]
SEGMENT_WIDTHS = (16, 208, 32)       # code / data / accum
RECURSION_WIDTHS = (16, 128, 16)     # placeholder widths of the recursion circuit (po2 = 18)
RECURSION_PO2 = 18
SEED_BASE = 0xB2000000               # segment i uses seed SEED_BASE + i


@dataclass
class ProverOpts:
    """Stand-in for risc0_zkvm::ProverOpts: the shapes the prover is provisioned for."""
    segment_po2: int = 20                      # agent default --segment-po2 (workflow/src/lib.rs:83-84)
    segment_widths: tuple = SEGMENT_WIDTHS
    recursion_po2: int = RECURSION_PO2
    recursion_widths: tuple = RECURSION_WIDTHS
    slots: int = 2                             # proofs in flight per GPU
    device: int = 0


@dataclass
class Segment:
    """What the executor emits per 2^po2 cycles (tasks/executor.rs:721-757), synthetic form."""
    index: int
    po2: int = 20
    seed: Optional[int] = None
    trace: Optional[np.ndarray] = None         # optional host witness (w_code + w_data) x 2^po2, Montgomery u32
    assumptions: list = field(default_factory=list)   # claim digests this segment's guest output depends on (tasks/resolve.rs)

    def __post_init__(self):
        if self.seed is None:
            self.seed = SEED_BASE + self.index


@dataclass
class SegmentReceipt:
    seal: np.ndarray
    index: int
    po2: int
    assumptions: list = field(default_factory=list)

    def get_seal_bytes(self):
        return self.seal.tobytes()


@dataclass
class SuccinctReceipt:
    seal: np.ndarray
    kind: int
    claim: tuple = field(default_factory=tuple)   # (first_segment, last_segment) covered by this receipt
    assumptions: list = field(default_factory=list)   # unresolved assumption claim digests (hex), tasks/resolve.rs:47-60

    def claim_digest(self) -> str:
        """Stand-in for `receipt.claim.digest()` (tasks/resolve.rs:81): identifies this receipt as somebody's assumption."""
        import hashlib
        return hashlib.sha256(np.ascontiguousarray(self.seal, dtype="<u4").tobytes()).hexdigest()


class VerifierContext:
    """Placeholder for risc0_zkvm::VerifierContext (prove.rs:44): carries nothing on the synthetic path."""


class VerificationError(B200Error):
    """`verify_integrity` failed: the reference's `Err` from `verify_integrity_with_context` (tasks/prove.rs:56-58)."""

    def __init__(self, code, what=""):
        super().__init__("%sseal does not verify (check %d failed)" % (what + ": " if what else "", code))
        self.code = code
        self.what = what


@dataclass
class DeviceReceipt:
    """A succinct receipt whose seal lives in DEVICE memory (caller-owned buffer of `words` u32 at `ptr`): what moves between prove, lift
    and join on one GPU, or over NVLink between GPUs, without a host bounce.  `owner` keeps the backing allocation alive."""
    ptr: int
    words: int
    kind: int
    claim: tuple = field(default_factory=tuple)
    assumptions: list = field(default_factory=list)
    owner: object = None
    seal: Optional[np.ndarray] = None          # host copy, when one was asked for


class _Pinned:
    def __init__(self, L, words):
        self.L = L
        p = C.c_void_p()
        _lib.check(L.b200_host_alloc(C.byref(p), max(words, 1) * 4))
        self.ptr = p
        self.array = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(max(words, 1),))

    def free(self):
        if self.ptr:
            self.L.b200_host_free(self.ptr)
            self.ptr = None


class ProverServer:
    """One per GPU process, like the agent's `Rc<dyn ProverServer>` (not thread-safe, one task loop)."""

    def __init__(self, opts: ProverOpts):
        self.opts = opts
        self.L = _lib.require_gpu(opts.device)
        sw, rw = opts.segment_widths, opts.recursion_widths
        self.seg_circuit = Circuit(opts.segment_po2, sw[0], sw[1], sw[2], KIND_SEGMENT)
        # arena is sized for the larger of the two shapes
        maxc = Circuit(max(opts.segment_po2, opts.recursion_po2), max(sw[0], rw[0]), max(sw[1], rw[1]), max(sw[2], rw[2]), 0)
        h = C.c_void_p()
        _lib.check(self.L.b200_prover_create(C.byref(h), opts.device, C.byref(maxc), opts.slots))
        self.h = h
        max_words = max(self.seal_words(self.seg_circuit), self.seal_words(self._rec_circuit(KIND_LIFT)))
        self._seal_bufs = [_Pinned(self.L, max_words) for _ in range(opts.slots)]
        self._seal_bufs2 = [_Pinned(self.L, max_words) for _ in range(opts.slots)]      # second output of the composite tasks
        self._verdicts3 = [(C.c_int * 3)() for _ in range(opts.slots)]
        self._pending = [None] * opts.slots
        self._prefetched = [None] * opts.slots
        self._verdict = [C.c_int(0) for _ in range(opts.slots)]

    # -- helpers -------------------------------------------------------------------------------------------
    def _rec_circuit(self, kind):
        rw = self.opts.recursion_widths
        return Circuit(self.opts.recursion_po2, rw[0], rw[1], rw[2], kind)

    def _buf(self, slot):
        # an out-of-range slot is reported by the library ("slot out of range"), not by Python indexing
        return self._seal_bufs[slot].ptr if 0 <= slot < len(self._seal_bufs) else self._seal_bufs[0].ptr

    def seal_words(self, circuit):
        return self.L.b200_seal_words(C.byref(circuit))

    def device_bytes(self):
        return self.L.b200_prover_device_bytes(self.h)

    def close(self):
        if getattr(self, "h", None):
            for s in range(self.opts.slots):
                if self._pending[s] is not None:
                    self.L.b200_prover_wait(self.h, s)
            self.L.b200_prover_destroy(self.h)
            self.h = None
            for b in self._seal_bufs + self._seal_bufs2:
                b.free()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- asynchronous form (one proof in flight per slot) -------------------------------------------------
    def submit_segment(self, slot, segment: Segment):
        c = Circuit(segment.po2, *self.opts.segment_widths, KIND_SEGMENT)
        tr = None
        if segment.trace is not None:
            tr = np.ascontiguousarray(segment.trace, dtype=np.uint32)
            need = (c.w_code + c.w_data) << c.po2
            if tr.size != need:
                raise B200Error("segment trace has %d words, expected %d" % (tr.size, need))
        _lib.check(self.L.b200_prove_segment_async(self.h, slot, C.byref(c), segment.seed,
                                                    tr.ctypes.data_as(C.c_void_p) if tr is not None else None,
                                                    self._buf(slot)))
        self._pending[slot] = ("segment", c, segment, tr)

    def prefetch_segment(self, slot, segment: Segment):
        """Start the host -> device copy of `segment.trace` for the NEXT proof of `slot` while the slot is still proving; the
        following submit_segment(slot, segment) with the same trace buffer then starts without a copy on its own stream."""
        if segment.trace is None:
            return
        c = Circuit(segment.po2, *self.opts.segment_widths, KIND_SEGMENT)
        tr = np.ascontiguousarray(segment.trace, dtype=np.uint32)
        need = (c.w_code + c.w_data) << c.po2
        if tr.size != need:
            raise B200Error("segment trace has %d words, expected %d" % (tr.size, need))
        _lib.check(self.L.b200_prefetch_trace_async(self.h, slot, C.byref(c), tr.ctypes.data_as(C.c_void_p)))
        self._prefetched[slot] = tr          # keep the buffer alive until it is consumed

    def submit_recursion(self, slot, kind, a, b=None):
        c = self._rec_circuit(kind)
        sa = np.ascontiguousarray(a.seal, dtype=np.uint32)
        sb = np.ascontiguousarray(b.seal, dtype=np.uint32) if b is not None else None
        _lib.check(self.L.b200_recursion_async(self.h, slot, C.byref(c), sa.ctypes.data_as(C.c_void_p), sa.size,
                                               sb.ctypes.data_as(C.c_void_p) if sb is not None else None,
                                               sb.size if sb is not None else 0, self._buf(slot)))
        self._pending[slot] = ("recursion", c, (kind, a, b), (sa, sb))

    def wait(self, slot):
        pend = self._pending[slot]
        if pend is None:
            raise B200Error("slot %d has no proof in flight" % slot)
        _lib.check(self.L.b200_prover_wait(self.h, slot))
        self._pending[slot] = None
        what, c, arg, _keep = pend
        seal = self._seal_bufs[slot].array[: self.seal_words(c)].copy()
        if what == "segment":
            return SegmentReceipt(seal, arg.index, arg.po2, list(arg.assumptions))
        kind, a, b = arg
        lo = a.claim[0] if isinstance(a, SuccinctReceipt) else a.index
        if kind in (KIND_RESOLVE,):
            # the conditional receipt keeps its claim; the resolved assumption leaves its list
            gone = b.claim_digest()
            return SuccinctReceipt(seal, kind, tuple(a.claim), [x for x in a.assumptions if x != gone])
        last = b if b is not None else a
        hi = last.claim[1] if isinstance(last, SuccinctReceipt) else last.index
        asm = list(a.assumptions) + (list(b.assumptions) if b is not None and kind in (KIND_JOIN, KIND_JOIN_POVW) else [])
        return SuccinctReceipt(seal, kind, (lo, hi), asm if kind != KIND_UNION else [])

    # -- the agent's task bodies as single enqueues (device-resident receipts) ---------------------------------------------------
    def submit_prove_lift(self, slot, segment: Segment, d_out: int = 0, verify: bool = True, host_seals: bool = True):
        """tasks::prove::prover (tasks/prove.rs:44-108) as ONE enqueue: prove_segment -> verify_integrity -> lift -> verify_integrity, no host
        round trip in between.  `d_out` (optional): device buffer that receives the lifted seal.  wait_task(slot) returns the receipts."""
        c = Circuit(segment.po2, *self.opts.segment_widths, KIND_SEGMENT)
        lc = self._rec_circuit(KIND_LIFT)
        tr = None
        if segment.trace is not None:
            tr = np.ascontiguousarray(segment.trace, dtype=np.uint32)
            if tr.size != (c.w_code + c.w_data) << c.po2:
                raise B200Error("segment trace has %d words, expected %d" % (tr.size, (c.w_code + c.w_data) << c.po2))
        v = self._verdicts3[slot] if 0 <= slot < len(self._verdicts3) else self._verdicts3[0]
        _lib.check(self.L.b200_prove_lift_async(
            self.h, slot, C.byref(c), segment.seed, tr.ctypes.data_as(C.c_void_p) if tr is not None else None, C.byref(lc),
            self._buf(slot) if host_seals else None, self._seal_bufs2[slot].ptr if host_seals and 0 <= slot < len(self._seal_bufs2) else None,
            C.c_void_p(d_out) if d_out else None, C.cast(v, C.POINTER(C.c_int)) if verify else None))
        self._pending[slot] = ("prove_lift", (c, lc), (segment, d_out, verify, host_seals), tr)

    def submit_recursion_dev(self, slot, kind, a: "DeviceReceipt", b: Optional["DeviceReceipt"] = None, d_out: int = 0,
                             verify: bool = True, host_seal: bool = True):
        """tasks::join::join (tasks/join.rs:41-79; union / resolve alike) as ONE enqueue over device-resident receipts: verify_integrity of
        left and right against the circuits their metadata names, the recursion proof, verify_integrity of the result."""
        c = self._rec_circuit(kind)
        ca = self._rec_circuit(a.kind)
        cb = self._rec_circuit(b.kind) if b is not None else None
        v = self._verdicts3[slot] if 0 <= slot < len(self._verdicts3) else self._verdicts3[0]
        _lib.check(self.L.b200_recursion_verified_async(
            self.h, slot, C.byref(c), C.c_void_p(a.ptr), C.byref(ca), C.c_void_p(b.ptr) if b is not None else None,
            C.byref(cb) if cb is not None else None, self._buf(slot) if host_seal else None, C.c_void_p(d_out) if d_out else None,
            C.cast(v, C.POINTER(C.c_int)) if verify else None))
        self._pending[slot] = ("recursion_dev", c, (kind, a, b, d_out, verify, host_seal), None)

    def query(self, slot) -> bool:
        """True once everything enqueued on the slot has completed (never blocks); wait()/wait_task() then returns at once."""
        r = self.L.b200_prover_query(self.h, slot)
        if r < 0:
            _lib.check(self.L.b200_last_error())
        return r == 1

    def wait_task(self, slot):
        """Completion of submit_prove_lift -> (SegmentReceipt | None, DeviceReceipt) or submit_recursion_dev -> DeviceReceipt.  Raises
        VerificationError naming the step whose verify_integrity failed (the agent turns that into a task retry / failure)."""
        pend = self._pending[slot]
        if pend is None or pend[0] not in ("prove_lift", "recursion_dev"):
            raise B200Error("slot %d has no composite task in flight" % slot)
        _lib.check(self.L.b200_prover_wait(self.h, slot))
        self._pending[slot] = None
        v = self._verdicts3[slot]
        if pend[0] == "prove_lift":
            (c, lc), (segment, d_out, verify, host_seals) = pend[1], pend[2]
            if verify:
                for code, what in ((v[0], "segment receipt"), (v[1], "lift receipt")):
                    if code != 0:
                        raise VerificationError(int(code), what)
            seg_r = lift_seal = None
            if host_seals:
                seg_r = SegmentReceipt(self._seal_bufs[slot].array[: self.seal_words(c)].copy(), segment.index, segment.po2,
                                       list(segment.assumptions))
                lift_seal = self._seal_bufs2[slot].array[: self.seal_words(lc)].copy()
            return seg_r, DeviceReceipt(d_out, self.seal_words(lc), KIND_LIFT, (segment.index, segment.index), list(segment.assumptions),
                                        None, lift_seal)
        c, (kind, a, b, d_out, verify, host_seal) = pend[1], pend[2]
        if verify:
            for code, what in ((v[0], "left receipt"), (v[1], "right receipt"), (v[2], "joined receipt")):
                if code != 0:
                    raise VerificationError(int(code), what)
        seal = self._seal_bufs[slot].array[: self.seal_words(c)].copy() if host_seal else None
        if kind == KIND_RESOLVE:
            gone = b.claim_digest() if hasattr(b, "claim_digest") else None
            return DeviceReceipt(d_out, self.seal_words(c), kind, tuple(a.claim), [x for x in a.assumptions if x != gone], None, seal)
        last = b if b is not None else a
        asm = list(a.assumptions) + (list(b.assumptions) if b is not None and kind == KIND_JOIN else [])
        return DeviceReceipt(d_out, self.seal_words(c), kind, (a.claim[0], last.claim[1]), asm if kind != KIND_UNION else [], None, seal)

    # -- verify_integrity on the device (tasks/prove.rs:56-58, :106-108; tasks/join.rs:77-79) ---------------------------
    def expected_circuit(self, receipt):
        """The circuit a receipt of this type MUST have been proved with on this server: segment receipts carry their po2 (bounded by
        the provisioned one), succinct receipts are recursion-shaped with the kind their metadata names.  verify_integrity checks the
        seal's header against it, so a small seal of another kind cannot pass as a lift or join receipt."""
        if isinstance(receipt, SegmentReceipt):
            return Circuit(receipt.po2, *self.opts.segment_widths, KIND_SEGMENT)
        if receipt.kind == KIND_KECCAK:        # keccak proofs come in the size their request asked for: the seal length tells which
            rw = self.opts.recursion_widths
            for po2 in range(9, max(self.opts.segment_po2, self.opts.recursion_po2) + 1):
                c = Circuit(po2, rw[0], rw[1], rw[2], KIND_KECCAK)
                if self.seal_words(c) == int(np.asarray(receipt.seal).size):
                    return c
        return self._rec_circuit(receipt.kind)

    def submit_verify(self, slot, receipt=None, expect: Optional[Circuit] = None):
        """Enqueue the verification of `receipt` (slot must be idle) or, with receipt=None, of the seal the slot is producing
        (may follow submit_segment / submit_recursion directly: same stream, no host round trip).  wait_verify() gives the verdict."""
        if receipt is None:
            if expect is None:
                _lib.check(self.L.b200_verify_async(self.h, slot, None, 0, C.byref(self._verdict[slot])))
            else:
                _lib.check(self.L.b200_verify_circuit_async(self.h, slot, C.byref(expect), None, 0, 0, C.byref(self._verdict[slot])))
            return None
        c = expect or self.expected_circuit(receipt)
        if isinstance(receipt, DeviceReceipt) and receipt.seal is None:
            _lib.check(self.L.b200_verify_circuit_async(self.h, slot, C.byref(c), C.c_void_p(receipt.ptr), receipt.words, 1,
                                                        C.byref(self._verdict[slot])))
            return receipt
        seal = np.ascontiguousarray(receipt.seal, dtype=np.uint32)
        _lib.check(self.L.b200_verify_circuit_async(self.h, slot, C.byref(c), seal.ctypes.data_as(C.c_void_p), seal.size, 0,
                                                    C.byref(self._verdict[slot])))
        return seal

    def wait_verify(self, slot):
        _lib.check(self.L.b200_prover_wait(self.h, slot))
        code = int(self._verdict[slot].value)
        if code != 0:
            raise VerificationError(code)

    def verify_integrity(self, receipt, slot=0):
        """receipt.verify_integrity_with_context(&ctx): raises VerificationError unless the seal is valid."""
        if self._pending[slot] is not None:
            raise B200Error("slot %d busy" % slot)
        keep = self.submit_verify(slot, receipt)
        self.wait_verify(slot)
        del keep

    def verify_seal(self, seal, slot=0):
        """Header-trusting form (b200_verify_async): checks the seal as a proof of whatever circuit its own header names.  For tooling and
        the verifier-parity tests; the agent path uses verify_integrity, which binds the circuit the receipt is SUPPOSED to have."""
        if self._pending[slot] is not None:
            raise B200Error("slot %d busy" % slot)
        seal = np.ascontiguousarray(seal, dtype=np.uint32)
        _lib.check(self.L.b200_verify_async(self.h, slot, seal.ctypes.data_as(C.c_void_p), seal.size, C.byref(self._verdict[slot])))
        self.wait_verify(slot)

    def last_ms(self, slot):
        return float(self.L.b200_prover_last_ms(self.h, slot))

    def prove_and_lift_many(self, segments):
        """prove_segment + lift for a list of segments with every slot kept busy (tasks/prove.rs:44-104, pipelined):
        proofs rotate over the slots; a finished segment proof is lifted on the slot it ran on."""
        n, slots = len(segments), self.opts.slots
        out = [None] * n
        pending = []          # (slot, stage, index)
        free = list(range(slots))
        nxt = 0
        while nxt < n or pending:
            while free and nxt < n:
                s = free.pop(0)
                self.submit_segment(s, segments[nxt]); pending.append((s, "seg", nxt)); nxt += 1
            s, stage, i = pending.pop(0)
            r = self.wait(s)
            if stage == "seg":
                self.submit_recursion(s, KIND_LIFT, r); pending.append((s, "lift", i))
            else:
                out[i] = r; free.append(s)
        return out

    # -- the reference's method names (synchronous, slot 0) -------------------------------------------------
    def prove_segment(self, ctx: VerifierContext, segment: Segment) -> SegmentReceipt:
        self.submit_segment(0, segment)
        return self.wait(0)

    def lift(self, receipt: SegmentReceipt) -> SuccinctReceipt:
        self.submit_recursion(0, KIND_LIFT, receipt)
        return self.wait(0)

    def join(self, a: SuccinctReceipt, b: SuccinctReceipt) -> SuccinctReceipt:
        self.submit_recursion(0, KIND_JOIN, a, b)
        return self.wait(0)

    def resolve(self, conditional: SuccinctReceipt, assumption: SuccinctReceipt) -> SuccinctReceipt:
        self.submit_recursion(0, KIND_RESOLVE, conditional, assumption)
        return self.wait(0)

    def union(self, a: SuccinctReceipt, b: SuccinctReceipt) -> SuccinctReceipt:
        self.submit_recursion(0, KIND_UNION, a, b)
        return self.wait(0)

    def prove_keccak(self, claim_digest: str, po2: int, control_root: str, input_states: bytes) -> SuccinctReceipt:
        """prover.prove_keccak(&ProveKeccakRequest{claim_digest, po2, control_root, input}) (tasks/keccak.rs:55-75), synthetic form: a
        recursion-shaped proof of kind KIND_KECCAK with 2^po2 rows whose input digest is the digest of the keccak states (taken as
        16-bit words, each a field element); the receipt is a SuccinctReceipt<Unknown>: no segment claim, no assumptions."""
        if len(input_states) == 0 or len(input_states) % KECCAK_STATE_BYTES:
            raise B200Error("keccak input must be a non-empty multiple of %d bytes" % KECCAK_STATE_BYTES)
        rw = self.opts.recursion_widths
        c = Circuit(int(po2), rw[0], rw[1], rw[2], KIND_KECCAK)
        words = np.frombuffer(bytes(input_states), dtype="<u2").astype(np.uint32)
        _lib.check(self.L.b200_recursion_async(self.h, 0, C.byref(c), words.ctypes.data_as(C.c_void_p), words.size, None, 0, self._buf(0)))
        _lib.check(self.L.b200_prover_wait(self.h, 0))
        seal = self._seal_bufs[0].array[: self.seal_words(c)].copy()
        r = SuccinctReceipt(seal, KIND_KECCAK, (0, 0), [])
        r.keccak = (claim_digest, int(po2), control_root)
        return r

    # proof-of-verifiable-work variants (POVW_LOG_ID set: workflow/src/lib.rs:209-212)
    def lift_povw(self, receipt: SegmentReceipt) -> SuccinctReceipt:
        """prover.lift_povw(&segment_receipt) -> SuccinctReceipt<WorkClaim<ReceiptClaim>> (tasks/prove.rs:70-78)"""
        self.submit_recursion(0, KIND_LIFT_POVW, receipt)
        return self.wait(0)

    def join_povw(self, a: SuccinctReceipt, b: SuccinctReceipt) -> SuccinctReceipt:
        """prover.join_povw(&left, &right) over two WorkClaim receipts (tasks/join_povw.rs:55)"""
        for r in (a, b):
            if r.kind not in POVW_KINDS:
                raise B200Error("join_povw needs proof-of-verifiable-work receipts (kind %d given)" % r.kind)
        self.submit_recursion(0, KIND_JOIN_POVW, a, b)
        return self.wait(0)

    def unwrap_povw(self, receipt: SuccinctReceipt) -> SuccinctReceipt:
        """prover.unwrap_povw(&povw_receipt) -> SuccinctReceipt<ReceiptClaim>: drops the work claim (tasks/resolve_povw.rs:57)"""
        if receipt.kind not in POVW_KINDS:
            raise B200Error("unwrap_povw needs a proof-of-verifiable-work receipt (kind %d given)" % receipt.kind)
        self.submit_recursion(0, KIND_UNWRAP_POVW, receipt)
        return self.wait(0)


def get_prover_server(opts: Optional[ProverOpts] = None) -> ProverServer:
    """Mirror of risc0_zkvm::get_prover_server(&ProverOpts) (workflow/src/lib.rs:276-284)."""
    return ProverServer(opts or ProverOpts())
