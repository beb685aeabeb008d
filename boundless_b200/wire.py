"""Wire formats either side of the proving path: the task definitions the agent receives and the blobs it reads / writes.

Reference: `workflow_common` (/root/reference/prover/crates/workflow-common/src/lib.rs:13-188) -- the `TaskType` enum travels as
serde_json (externally tagged: {"Prove": {"index": 3}}, unit variant "Finalize"; parsed at workflow/src/lib.rs:688-689); segments
and receipts travel as bincode 1.x blobs (`serialize_obj` / `deserialize_obj`, workflow/src/tasks/mod.rs:57-66: little-endian
fixed-width integers, u64 length prefixes, u8 Option tags, u32 enum variant indices).

The risc0 `Segment` / `SuccinctReceipt` layouts themselves belong to un-vendored crates and the circuit here is synthetic, so the
blob *payloads* are this package's own structs; the *encoding rules* are bincode's, so a Rust struct with the same field list
round-trips them.
"""
import json
import struct
from dataclasses import MISSING, dataclass, field
from typing import List, Optional

import numpy as np

# worker stream identifiers (workflow-common/src/lib.rs:13-32)
AUX_WORK_TYPE = "aux"
EXEC_WORK_TYPE = "exec"
PROVE_WORK_TYPE = "prove"
COPROC_WORK_TYPE = "coproc"
JOIN_WORK_TYPE = "join"
SNARK_WORK_TYPE = "snark"
KECCAK_RECEIPT_PATH = "keccak_receipts"
# hot-store key prefixes (workflow/src/tasks/mod.rs:22-35)
RECUR_RECEIPT_PATH = "recursion_receipts"
RESOLVED_RECEIPT_PATH = "resolved_receipt"
SEGMENTS_PATH = "segments"
RECEIPT_PATH = "receipts"
COPROC_CB_PATH = "coproc"
# shared-storage layout of the final receipt (workflow-common/src/storage.rs; tasks/finalize.rs:74)
RECEIPT_BUCKET_DIR = "receipts"
STARK_BUCKET_DIR = "stark"

COMPRESS_TYPES = ("None", "Groth16", "Blake3Groth16")


class WireError(ValueError):
    pass


# ---- task definitions (serde_json) ------------------------------------------------------------------------------------
@dataclass
class ExecutorReq:
    image: str
    input: str
    user_id: str
    assumptions: List[str] = field(default_factory=list)
    execute_only: bool = False
    compress: str = "None"
    exec_limit: Optional[int] = None


@dataclass
class ProveReq:
    index: int


@dataclass
class JoinReq:
    idx: int
    left: int
    right: int


@dataclass
class UnionReq:
    idx: int
    left: int
    right: int


@dataclass
class ResolveReq:
    max_idx: int
    union_max_idx: Optional[int] = None


@dataclass
class SnarkReq:
    receipt: str
    compress_type: str = "None"


@dataclass
class KeccakReq:
    claim_digest: str          # risc0 `Digest` in a human-readable format (JSON) is a 64-character hex string
    po2: int
    control_root: str


@dataclass
class Finalize:
    pass


_VARIANTS = {"Executor": ExecutorReq, "Prove": ProveReq, "Join": JoinReq, "Resolve": ResolveReq, "Snark": SnarkReq,
             "Keccak": KeccakReq, "Union": UnionReq}
_JOB_TYPE_STR = {ExecutorReq: "executor", ProveReq: "prove-lift", JoinReq: "join", ResolveReq: "resolve", Finalize: "finalize",
                 SnarkReq: "snark", KeccakReq: "keccak", UnionReq: "union"}


def task_type_to_value(task):
    """serde_json::to_value(TaskType::X(req)) as a Python object (dict, or the string "Finalize")."""
    if isinstance(task, Finalize):
        return "Finalize"
    for name, cls in _VARIANTS.items():
        if isinstance(task, cls):
            return {name: dict(task.__dict__)}
    raise WireError("not a TaskType: %r" % (task,))


def task_type_from_value(value):
    """serde_json::from_value::<TaskType>(task_def) (workflow/src/lib.rs:688-689); WireError = "Invalid task_def"."""
    if isinstance(value, str):
        if value == "Finalize":
            return Finalize()
        raise WireError("unknown variant `%s`" % value)
    if not isinstance(value, dict) or len(value) != 1:
        raise WireError("expected an externally tagged enum")
    (name, body), = value.items()
    cls = _VARIANTS.get(name)
    if cls is None:
        raise WireError("unknown variant `%s`" % name)
    if not isinstance(body, dict):
        raise WireError("invalid type for variant `%s`" % name)
    # serde semantics of the reference structs (no `deny_unknown_fields`, no `#[serde(default)]`): unknown fields are IGNORED (a newer
    # control plane may add some), every field that is not an Option is REQUIRED, a missing Option is None
    fields = cls.__dataclass_fields__
    optional = {"exec_limit", "union_max_idx"}
    missing = [f for f in fields if f not in body and f not in optional]
    if missing:
        raise WireError("missing field `%s`" % missing[0])
    body = {k: v for k, v in body.items() if k in fields}
    for k in ("index", "idx", "left", "right", "max_idx", "po2"):
        if k in body and (not isinstance(body[k], int) or isinstance(body[k], bool) or body[k] < 0):
            raise WireError("invalid value for `%s`" % k)
    for k in ("compress", "compress_type"):
        if k in body and body[k] not in COMPRESS_TYPES:
            raise WireError("unknown variant `%s`, expected one of `None`, `Groth16`, `Blake3Groth16`" % (body[k],))
    for k in ("claim_digest", "control_root"):
        if k in body:
            v = body[k]
            if not isinstance(v, str) or len(v) != 64 or any(c not in "0123456789abcdefABCDEF" for c in v):
                raise WireError("invalid digest for `%s`" % k)
    return cls(**body)


def task_type_to_json(task):
    return json.dumps(task_type_to_value(task), separators=(",", ":"))


def task_type_from_json(text):
    try:
        return task_type_from_value(json.loads(text))
    except json.JSONDecodeError as e:
        raise WireError(str(e))


def to_job_type_str(task):
    """TaskType::to_job_type_str (workflow-common/src/lib.rs:171-184)."""
    return _JOB_TYPE_STR[type(task)]


# ---- bincode 1.x primitives ---------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self):
        self.parts = []

    def u8(self, v): self.parts.append(struct.pack("<B", v))
    def u32(self, v): self.parts.append(struct.pack("<I", v))
    def u64(self, v): self.parts.append(struct.pack("<Q", v))
    def boolean(self, v): self.u8(1 if v else 0)

    def string(self, s):
        b = s.encode("utf-8"); self.u64(len(b)); self.parts.append(b)

    def bytes_(self, b):
        self.u64(len(b)); self.parts.append(bytes(b))

    def vec_u32(self, a):
        a = np.ascontiguousarray(a, dtype="<u4"); self.u64(a.size); self.parts.append(a.tobytes())

    def done(self):
        return b"".join(self.parts)


class _Reader:
    def __init__(self, buf):
        self.buf = memoryview(bytes(buf)); self.pos = 0

    def _take(self, n):
        if self.pos + n > len(self.buf):
            raise WireError("io error: unexpected end of file")          # bincode's message for a short blob
        v = self.buf[self.pos:self.pos + n]; self.pos += n
        return v

    def u8(self): return struct.unpack("<B", self._take(1))[0]
    def u32(self): return struct.unpack("<I", self._take(4))[0]
    def u64(self): return struct.unpack("<Q", self._take(8))[0]

    def boolean(self):
        v = self.u8()
        if v > 1:
            raise WireError("invalid u8 while decoding bool, expected 0 or 1, found %d" % v)
        return bool(v)

    def option_tag(self):
        v = self.u8()
        if v > 1:
            raise WireError("invalid tag encoding for Option, found %d" % v)
        return bool(v)

    def string(self):
        n = self.u64()
        try:
            return bytes(self._take(n)).decode("utf-8")
        except UnicodeDecodeError:
            raise WireError("string is not valid utf8")

    def bytes_(self):
        return bytes(self._take(self.u64()))

    def vec_u32(self):
        n = self.u64()
        if n > (len(self.buf) - self.pos) // 4:
            raise WireError("io error: unexpected end of file")
        return np.frombuffer(self._take(4 * n), dtype="<u4").astype(np.uint32)

    def finish(self):
        # bincode::deserialize ignores trailing bytes; the agent never relies on that, so be strict
        if self.pos != len(self.buf):
            raise WireError("%d trailing bytes" % (len(self.buf) - self.pos))


# ---- blobs ------------------------------------------------------------------------------------------------------------
# struct Segment { index: u64, po2: u32, seed: u64, trace: Option<Vec<u32>>, assumptions: Vec<String> }
def serialize_segment(seg):
    w = _Writer()
    w.u64(seg.index); w.u32(seg.po2); w.u64(seg.seed)
    if seg.trace is None:
        w.u8(0)
    else:
        w.u8(1); w.vec_u32(seg.trace)
    asm = list(getattr(seg, "assumptions", ()) or ())
    w.u64(len(asm))
    for a in asm:
        w.string(a)
    return w.done()


def deserialize_segment(buf):
    from .prover_server import Segment
    r = _Reader(buf)
    index, po2, seed = r.u64(), r.u32(), r.u64()
    trace = r.vec_u32() if r.option_tag() else None
    asm = [r.string() for _ in range(r.u64())]
    r.finish()
    seg = Segment(index=index, po2=po2, seed=seed, trace=trace)
    seg.assumptions = asm
    return seg


# struct SuccinctReceipt { seal: Vec<u32>, kind: u32, claim: (u64, u64), assumptions: Vec<String> }
def serialize_succinct(rcpt):
    w = _Writer()
    w.vec_u32(rcpt.seal); w.u32(rcpt.kind)
    lo, hi = rcpt.claim if rcpt.claim else (0, 0)
    w.u64(lo); w.u64(hi)
    asm = list(getattr(rcpt, "assumptions", ()) or ())
    w.u64(len(asm))
    for a in asm:
        w.string(a)
    return w.done()


def deserialize_succinct(buf):
    from .prover_server import SuccinctReceipt
    r = _Reader(buf)
    seal = r.vec_u32(); kind = r.u32(); lo = r.u64(); hi = r.u64()
    asm = [r.string() for _ in range(r.u64())]
    r.finish()
    rc = SuccinctReceipt(seal, kind, (lo, hi))
    rc.assumptions = asm
    return rc


# struct Receipt { inner: InnerReceipt::Succinct(SuccinctReceipt) (variant index 1 as in risc0's enum order
# Composite/Succinct/Groth16/Fake), journal: Vec<u8> }
def serialize_rollup(succinct, journal: bytes):
    w = _Writer()
    w.u32(1)
    w.parts.append(serialize_succinct(succinct))
    w.bytes_(journal)
    return w.done()


def deserialize_rollup(buf):
    r = _Reader(buf)
    variant = r.u32()
    if variant != 1:
        raise WireError("rollup receipt is not Succinct (variant %d)" % variant)          # BENTO-FINALIZE-002
    # re-use the succinct decoder on the embedded struct: decode field by field
    seal = r.vec_u32(); kind = r.u32(); lo = r.u64(); hi = r.u64()
    asm = [r.string() for _ in range(r.u64())]
    journal = r.bytes_()
    r.finish()
    from .prover_server import SuccinctReceipt
    rc = SuccinctReceipt(seal, kind, (lo, hi))
    rc.assumptions = asm
    return rc, journal


def serialize_journal(journal: bytes):
    w = _Writer(); w.bytes_(journal); return w.done()


def deserialize_journal(buf):
    r = _Reader(buf); j = r.bytes_(); r.finish(); return j
