"""The agent's task functions for the proving path: fetch -> prove -> verify -> lift -> verify -> store, and the join / union /
resolve / finalize steps that reduce a job's segments to one receipt.

Reference (all under /root/reference/prover/crates/workflow/src): `tasks::prove::prover` (tasks/prove.rs:17-129),
`tasks::join::join` (tasks/join.rs:17-100), `tasks::union::union` (tasks/union.rs:17-80), `tasks::resolve::resolver`
(tasks/resolve.rs:21-190), `tasks::finalize::finalize` (tasks/finalize.rs:20-99), the dispatcher `Agent::process_work`
(lib.rs:686-799) and the claim / retry loop `Agent::poll_work` (lib.rs:611-677).  Key names, the order of the steps, which keys
are deleted after a task is marked done, the error contexts ("[BENTO-...]") and the retry / "retry max hit" / 1024-character
truncation rules are the reference's; the prover behind them is `boundless_b200.ProverServer` (device-side proofs AND
device-side verify_integrity), and storage / task database are interfaces (`HotStore`, `taskdb.MemoryTaskDb`) because the
reference's Redis / REST control plane is out of scope.

The executor here is the synthetic producer of SURVEY.md 8d: it emits `n_segments` seeded segments, planning the join tree
online with the reference Planner exactly as tasks/executor.rs:49-257 does (segment tasks have no prerequisites, a join is
created the moment two peaks merge, `resolve` and `finalize` hang off the root).
"""
import json
from dataclasses import dataclass
from typing import List, Optional

from . import wire
from .planner import CMD_FINALIZE, CMD_JOIN, CMD_KECCAK, CMD_SEGMENT, CMD_UNION, Planner
from .prover_server import Segment, SuccinctReceipt
from .taskdb import ReadyTask


class TaskError(Exception):
    """anyhow::Error with a context chain; str() renders like `{:#}` ("outer: inner")."""

    def __init__(self, msg, cause=None):
        self.msg, self.cause = msg, cause
        super().__init__(self.__str__())

    def __str__(self):
        return self.msg if self.cause is None else "%s: %s" % (self.msg, self.cause)


def _ctx(msg, fn, *a, **kw):
    """`fn(...).context(msg)`"""
    try:
        return fn(*a, **kw)
    except Exception as e:               # noqa: BLE001 -- anyhow wraps every error type
        raise TaskError(msg, e)


class KeyNotFound(KeyError):
    def __str__(self):
        return "key not found: %s" % (self.args[0],)


class MemoryHotStore:
    """The Redis hot store (`hot_get_bytes` / `hot_set_bytes` / `hot_delete`) plus the shared object store (`write_asset`)."""

    def __init__(self):
        self.kv = {}
        self.assets = {}
        self.gets = self.sets = self.deletes = 0

    def get_bytes(self, key: str) -> bytes:
        self.gets += 1
        if key not in self.kv:
            raise KeyNotFound(key)
        return self.kv[key]

    def set_bytes(self, key: str, value: bytes) -> None:
        self.sets += 1
        self.kv[key] = bytes(value)

    def delete(self, key: str) -> None:
        self.deletes += 1
        self.kv.pop(key, None)

    def write_asset(self, key: str, value: bytes) -> None:
        self.assets[key] = bytes(value)


@dataclass
class AgentArgs:
    """The retry / timeout knobs of the agent CLI (workflow/src/lib.rs:83-200 defaults)."""
    task_stream: str = wire.PROVE_WORK_TYPE
    segment_po2: int = 20
    prove_retries: int = 3
    prove_timeout: int = 30
    join_retries: int = 3
    join_timeout: int = 10
    resolve_retries: int = 3
    resolve_timeout: int = 10
    finalize_retries: int = 0
    finalize_timeout: int = 10
    # executor.rs:517-545: with JOIN_STREAM / UNION_STREAM set in the environment, join (and resolve) / union tasks go to the customer's
    # "join" stream instead of the prove stream, so dedicated workers can take them.  None = read the environment like the reference.
    join_stream: Optional[bool] = None
    union_stream: Optional[bool] = None
    povw_job_number: int = 0          # WorkClaim.work.nonce_min.job of this agent's jobs (resolve_povw.rs:253-262 metadata)


class Agent:
    """workflow::Agent: one per worker process; `prover` is set for GPU streams only (lib.rs:201, :276-284)."""

    def __init__(self, task_db, store, prover=None, args: Optional[AgentArgs] = None, povw: bool = False):
        self.task_db, self.store, self.prover = task_db, store, prover
        self.args = args or AgentArgs()
        self.povw = povw
        self.processed: List[str] = []          # "<job>|<task>" in completion order (for tests / tracing)
        self.errors: List[str] = []             # "<job>|<task>: <error>" of every failed attempt, retried or not

    # hot store helpers (lib.rs hot_get_bytes / hot_set_bytes / hot_delete)
    def hot_get_bytes(self, key): return self.store.get_bytes(key)
    def hot_set_bytes(self, key, value): self.store.set_bytes(key, value)
    def hot_delete(self, key): self.store.delete(key)
    def is_povw_enabled(self): return bool(self.povw)

    def _prover(self, tag):
        if self.prover is None:
            raise TaskError(tag)
        return self.prover

    def _verify(self, receipt, msg):
        _ctx(msg, self.prover.verify_integrity, receipt)


# ---- tasks/prove.rs ---------------------------------------------------------------------------------------------------------
def prover(agent: Agent, job_id: str, task_id: str, request: wire.ProveReq) -> List[str]:
    job_prefix = "job:%s" % job_id
    segment_key = "%s:%s:%d" % (job_prefix, wire.SEGMENTS_PATH, request.index)
    segment_vec = _ctx("segment data not found for segment key: %s" % segment_key, agent.hot_get_bytes, segment_key)
    segment = _ctx("Failed to deserialize segment data from redis", wire.deserialize_segment, segment_vec)
    p = agent._prover("[BENTO-PROVE-002] Missing prover from prove task")
    segment_receipt = p.prove_segment(None, segment)
    agent._verify(segment_receipt, "[BENTO-PROVE-004] Failed to verify segment receipt integrity")
    output_key = "%s:%s:%s" % (job_prefix, wire.RECUR_RECEIPT_PATH, task_id)
    if agent.is_povw_enabled():            # prove.rs:67-93
        lift_receipt = agent._prover("[BENTO-PROVE-005] Missing prover from resolve task").lift_povw(segment_receipt)
        agent._verify(lift_receipt, "Failed to verify lift receipt integrity")
        lift_asset = _ctx("Failed to serialize the POVW segment", wire.serialize_succinct, lift_receipt)
        _ctx("Failed to set POVW receipt key with expiry", agent.hot_set_bytes, output_key, lift_asset)
        return [segment_key]
    lift_receipt = agent._prover("[BENTO-PROVE-008] Missing prover from resolve task").lift(segment_receipt)
    agent._verify(lift_receipt, "[BENTO-PROVE-010] Failed to verify lift receipt integrity")
    lift_asset = _ctx("Failed to serialize the segment", wire.serialize_succinct, lift_receipt)
    _ctx("Failed to set receipt key with expiry", agent.hot_set_bytes, output_key, lift_asset)
    return [segment_key]


# ---- tasks/join.rs ------------------------------------------------------------------------------------------------------------
def join(agent: Agent, job_id: str, request: wire.JoinReq) -> List[str]:
    prefix = "job:%s:%s" % (job_id, wire.RECUR_RECEIPT_PATH)
    left_key, right_key = "%s:%d" % (prefix, request.left), "%s:%d" % (prefix, request.right)
    left = _ctx("failed to get receipt for key: %s" % left_key, agent.hot_get_bytes, left_key)
    right = _ctx("failed to get receipt for key: %s" % right_key, agent.hot_get_bytes, right_key)
    left = _ctx("[BENTO-JOIN-001] Failed to deserialize left receipt", wire.deserialize_succinct, left)
    right = _ctx("[BENTO-JOIN-002] Failed to deserialize right receipt", wire.deserialize_succinct, right)
    p = agent._prover("Missing prover from join task")
    agent._verify(left, "[BENTO-JOIN-003] Failed to verify left receipt integrity")
    agent._verify(right, "[BENTO-JOIN-004] Failed to verify right receipt integrity")
    joined = p.join(left, right)
    agent._verify(joined, "[BENTO-JOIN-006] Failed to verify join receipt integrity")
    blob = _ctx("Failed to serialize the joined receipt", wire.serialize_succinct, joined)
    _ctx("Failed to store joined receipt", agent.hot_set_bytes, "%s:%d" % (prefix, request.idx), blob)
    return [left_key, right_key]


# ---- tasks/join_povw.rs ---------------------------------------------------------------------------------------------------------
def join_povw(agent: Agent, job_id: str, request: wire.JoinReq) -> List[str]:
    prefix = "job:%s:%s" % (job_id, wire.RECUR_RECEIPT_PATH)
    left_key, right_key = "%s:%d" % (prefix, request.left), "%s:%d" % (prefix, request.right)
    left = _ctx("failed to get receipt for key: %s" % left_key, agent.hot_get_bytes, left_key)
    right = _ctx("failed to get receipt for key: %s" % right_key, agent.hot_get_bytes, right_key)
    left, right = wire.deserialize_succinct(left), wire.deserialize_succinct(right)
    if agent.prover is None:
        raise TaskError("No prover available for join task")
    agent._verify(left, "[BENTO-JOINPOVW-001] Failed to verify left receipt integrity")
    agent._verify(right, "[BENTO-JOINPOVW-002] Failed to verify right receipt integrity")
    joined = _ctx("POVW join method not available - POVW functionality requires RISC Zero POVW support", agent.prover.join_povw, left, right)
    agent._verify(joined, "[BENTO-JOINPOVW-003] Failed to verify joined POVW receipt integrity")
    blob = _ctx("[BENTO-JOINPOVW-004] Failed to serialize joined POVW receipt", wire.serialize_succinct, joined)
    _ctx("Failed to write joined POVW receipt to hot store", agent.hot_set_bytes, "%s:%d" % (prefix, request.idx), blob)
    return [left_key, right_key]


# ---- tasks/keccak.rs ------------------------------------------------------------------------------------------------------------
KECCAK_STATE_BYTES = 200            # size_of::<[u64; 25]>()


def keccak(agent: Agent, job_id: str, task_id: str, request: wire.KeccakReq) -> List[str]:
    """The keccak coprocessor's prove task: input states from the hot store (task-scoped key, legacy key as fall-back), prove_keccak,
    receipt to `job:{id}:keccak_receipts:{task}` where the union tree picks it up (tasks/keccak.rs:25-108)."""
    input_path = "job:%s:%s:%s:%s" % (job_id, wire.COPROC_CB_PATH, task_id, request.claim_digest)
    try:
        data = agent.hot_get_bytes(input_path)
    except Exception:                       # noqa: BLE001 -- [BENTO-KECCAK-013]: fall back to the legacy location
        data = agent.hot_get_bytes("job:%s:%s:%s" % (job_id, wire.COPROC_CB_PATH, request.claim_digest))
    if len(data) % KECCAK_STATE_BYTES:
        raise TaskError("[BENTO-KECCAK-001] Input length must be a multiple of KeccakState size")
    if not data:
        raise TaskError("[BENTO-KECCAK-002] Received empty keccak input with claim_digest: %s" % request.claim_digest)
    p = agent._prover("[BENTO-KECCAK-003] Missing prover from keccak prove task")
    receipt = _ctx("Failed to prove_keccak", p.prove_keccak, request.claim_digest, request.po2, request.control_root, data)
    blob = _ctx("[BENTO-KECCAK-005] Failed to serialize keccak receipt", wire.serialize_succinct, receipt)
    _ctx("Failed to write keccak receipt to hot store", agent.hot_set_bytes,
         "job:%s:%s:%s" % (job_id, wire.KECCAK_RECEIPT_PATH, task_id), blob)
    return [input_path]


# ---- tasks/union.rs -----------------------------------------------------------------------------------------------------------
def union(agent: Agent, job_id: str, request: wire.UnionReq) -> List[str]:
    prefix = "job:%s:%s" % (job_id, wire.KECCAK_RECEIPT_PATH)
    left_key, right_key = "%s:%d" % (prefix, request.left), "%s:%d" % (prefix, request.right)
    left = _ctx("failed to get receipt for key: %s" % left_key, agent.hot_get_bytes, left_key)
    right = _ctx("failed to get receipt for key: %s" % right_key, agent.hot_get_bytes, right_key)
    left = _ctx("[BENTO-UNION-001] Failed to deserialize left receipt", wire.deserialize_succinct, left)
    right = _ctx("[BENTO-UNION-002] Failed to deserialize right receipt", wire.deserialize_succinct, right)
    p = agent._prover("[BENTO-UNION-003] Missing prover from union prove task")
    unioned = _ctx("[BENTO-UNION-004] Failed to union on left/right receipt", p.union, left, right)
    agent._verify(unioned, "[BENTO-UNION-005] Failed to verify union receipt integrity")
    blob = _ctx("[BENTO-UNION-006] Failed to serialize union receipt", wire.serialize_succinct, unioned)
    _ctx("[BENTO-UNION-007] Failed to set hot-store key for union receipt", agent.hot_set_bytes, "%s:%d" % (prefix, request.idx), blob)
    return [left_key, right_key]


# ---- tasks/resolve.rs -----------------------------------------------------------------------------------------------------------
def resolver(agent: Agent, job_id: str, request: wire.ResolveReq):
    """Returns (assumption count or None, cleanup keys)."""
    job_prefix = "job:%s" % job_id
    receipts_key = "%s:%s" % (job_prefix, wire.RECEIPT_PATH)
    root_key = "%s:%s:%d" % (job_prefix, wire.RECUR_RECEIPT_PATH, request.max_idx)
    cleanup = [root_key]
    blob = _ctx("segment data not found for root receipt key: %s" % root_key, agent.hot_get_bytes, root_key)
    conditional = wire.deserialize_succinct(blob)
    assumptions_len = None
    if conditional.assumptions:
        assumptions = list(conditional.assumptions)
        assumptions_len = len(assumptions)
        union_claim = ""
        if request.union_max_idx is not None:
            ukey = "%s:%s:%d" % (job_prefix, wire.KECCAK_RECEIPT_PATH, request.union_max_idx)
            ublob = _ctx("Failed to get union receipt: %s" % ukey, agent.hot_get_bytes, ukey)
            union_receipt = _ctx("[BENTO-RESOLVE-004] Failed to deserialize to SuccinctReceipt<Unknown> type", wire.deserialize_succinct, ublob)
            union_claim = union_receipt.claim_digest()
            p = agent._prover("[BENTO-RESOLVE-005] Missing prover from resolve task")
            conditional = _ctx("Failed to resolve the union receipt", p.resolve, conditional, union_receipt)
        for claim in assumptions:
            if claim == union_claim:
                continue
            akey = "%s:%s" % (receipts_key, claim)
            cleanup.append(akey)
            ablob = _ctx("corroborating receipt not found: key %s" % akey, agent.hot_get_bytes, akey)
            arec = _ctx("[BENTO-RESOLVE-008] could not deserialize assumption receipt: %s" % akey, wire.deserialize_succinct, ablob)
            p = agent._prover("[BENTO-RESOLVE-009] Missing prover from resolve task")
            conditional = _ctx("Failed to resolve the conditional receipt", p.resolve, conditional, arec)
    out = _ctx("[BENTO-RESOLVE-011] Failed to serialize resolved receipt", wire.serialize_succinct, conditional)
    _ctx("Failed to set resolved receipt key with expiry", agent.hot_set_bytes, "%s:%s" % (job_prefix, wire.RESOLVED_RECEIPT_PATH), out)
    return assumptions_len, cleanup


# ---- tasks/resolve_povw.rs ------------------------------------------------------------------------------------------------------
WORK_RECEIPTS_BUCKET_DIR = "work_receipts"          # workflow-common/src/storage.rs:41


def resolve_povw(agent: Agent, job_id: str, request: wire.ResolveReq):
    """The PoVW root is first unwrapped to a plain ReceiptClaim receipt (prover.unwrap_povw), assumptions are resolved as in
    tasks/resolve.rs, and the ORIGINAL PoVW receipt is saved to the work-receipts bucket with its metadata (resolve_povw.rs:214-268)."""
    import os
    job_prefix = "job:%s" % job_id
    receipts_key = "%s:%s" % (job_prefix, wire.RECEIPT_PATH)
    root_key = "%s:%s:%d" % (job_prefix, wire.RECUR_RECEIPT_PATH, request.max_idx)
    cleanup = [root_key]
    blob = _ctx("segment data not found for root receipt key: %s" % root_key, agent.hot_get_bytes, root_key)
    povw_receipt = _ctx("Failed to deserialize as POVW receipt", wire.deserialize_succinct, blob)
    p = agent._prover("Missing prover for POVW resolve task")
    conditional = _ctx("POVW unwrap failed", p.unwrap_povw, povw_receipt)
    assumptions_len = None
    if conditional.assumptions:
        assumptions = list(conditional.assumptions)
        assumptions_len = len(assumptions)
        union_claim = ""
        if request.union_max_idx is not None:
            ukey = "%s:%s:%d" % (job_prefix, wire.KECCAK_RECEIPT_PATH, request.union_max_idx)
            ublob = _ctx("Failed to get union receipt: %s" % ukey, agent.hot_get_bytes, ukey)
            if not ublob:
                raise TaskError("Union receipt is empty for key: %s" % ukey)
            union_receipt = _ctx("Failed to deserialize union receipt (size: %d bytes) from key: %s" % (len(ublob), ukey),
                                 wire.deserialize_succinct, ublob)
            union_claim = union_receipt.claim_digest()
            conditional = _ctx("Failed to resolve the union receipt", agent._prover("Missing prover from resolve task").resolve,
                               conditional, union_receipt)
        for claim in assumptions:
            if claim == union_claim:
                continue
            akey = "%s:%s" % (receipts_key, claim)
            cleanup.append(akey)
            ablob = _ctx("corroborating receipt not found: key %s" % akey, agent.hot_get_bytes, akey)
            if not ablob:
                raise TaskError("Assumption receipt is empty for key: %s" % akey)
            arec = _ctx("Failed to deserialize assumption receipt (size: %d bytes) from key: %s" % (len(ablob), akey),
                        wire.deserialize_succinct, ablob)
            conditional = _ctx("Failed to resolve the conditional receipt", agent._prover("Missing prover from resolve task").resolve,
                               conditional, arec)
    out = _ctx("Failed to serialize resolved receipt", wire.serialize_succinct, conditional)
    _ctx("Failed to set resolved receipt key with expiry", agent.hot_set_bytes, "%s:%s" % (job_prefix, wire.RESOLVED_RECEIPT_PATH), out)
    _ctx("Failed to save resolved POVW receipt to work receipts bucket", agent.store.write_asset,
         "%s/%s.bincode" % (WORK_RECEIPTS_BUCKET_DIR, job_id), wire.serialize_succinct(povw_receipt))
    meta = {"job_id": job_id}
    if os.environ.get("POVW_LOG_ID") is not None:
        meta["povw_log_id"] = os.environ["POVW_LOG_ID"]
    elif isinstance(agent.povw, str):
        meta["povw_log_id"] = agent.povw
    meta["povw_job_number"] = str(agent.args.povw_job_number)
    _ctx("Failed to save POVW metadata to work receipts bucket", agent.store.write_asset,
         "%s/%s_metadata.json" % (WORK_RECEIPTS_BUCKET_DIR, job_id), json.dumps(meta).encode())
    return assumptions_len, cleanup


# ---- tasks/finalize.rs ----------------------------------------------------------------------------------------------------------
def finalize(agent: Agent, job_id: str) -> List[str]:
    job_prefix = "job:%s" % job_id
    root_key = "%s:%s" % (job_prefix, wire.RESOLVED_RECEIPT_PATH)
    blob = _ctx("failed to get the root receipt key: %s" % root_key, agent.hot_get_bytes, root_key)
    root = _ctx("could not deserialize the root receipt. Data length: %d bytes" % len(blob), wire.deserialize_succinct, blob)
    journal_key = "%s:journal" % job_prefix
    jblob = _ctx("Journal data not found for key ID: %s" % journal_key, agent.hot_get_bytes, journal_key)
    journal = _ctx("could not deserialize the journal. Data length: %d bytes" % len(jblob), wire.deserialize_journal, jblob)
    image_key = "%s:image_id" % job_prefix
    image_raw = _ctx("Image ID not found for key: %s" % image_key, agent.hot_get_bytes, image_key)
    image_id = _ctx("[BENTO-FINALIZE-003] Failed to decode image ID bytes as UTF-8", image_raw.decode, "utf-8")
    _ctx("Failed to convert imageId file to digest from_hex", bytes.fromhex, image_id)
    if len(image_id) != 64:
        raise TaskError("Failed to convert imageId file to digest from_hex")
    # rollup_receipt.verify(image_id): the succinct receipt must verify and must have no unresolved assumptions
    if root.assumptions:
        raise TaskError("[BENTO-FINALIZE-001] Receipt verification failed", "unresolved assumptions")
    # never skipped: a worker without a verifier must not upload a receipt nobody checked (the reference's finalize always verifies,
    # tasks/finalize.rs:69).  What is NOT bound on the synthetic path: image_id and the claim -- the real binding is the recursion
    # circuit's own constraint system (out of scope), see DESIGN.md "known gaps".
    agent._prover("[BENTO-FINALIZE-001] Receipt verification failed: this worker has no verifier")
    agent._verify(root, "[BENTO-FINALIZE-001] Receipt verification failed")
    key = "%s/%s/%s.bincode" % (wire.RECEIPT_BUCKET_DIR, wire.STARK_BUCKET_DIR, job_id)
    _ctx("Failed to upload final receipt to shared storage", agent.store.write_asset, key, wire.serialize_rollup(root, journal))
    return [root_key, journal_key, image_key]


# ---- tasks/executor.rs (synthetic producer) ---------------------------------------------------------------------------------------
@dataclass
class ExecutorResp:
    segments: int
    user_cycles: int
    total_cycles: int
    assumption_count: int
    povw_log_id: Optional[str] = None
    povw_job_number: Optional[str] = None


def _process_task(agent: Agent, streams: dict, job_id: str, tree_task, segment_index: Optional[int], assumptions: List[str]):
    """executor.rs:49-257 `process_task`: turn one Planner task into a taskdb row."""
    a, db = agent.args, agent.task_db
    name = str(tree_task.task_number)
    if tree_task.command == CMD_SEGMENT:
        if segment_index is None:
            raise TaskError("[BENTO-EXEC-004] INVALID STATE: segment task without segment index")
        db.create_task(job_id, name, streams["prove"], wire.task_type_to_value(wire.ProveReq(segment_index)), [], a.prove_retries,
                       a.prove_timeout)
    elif tree_task.command == CMD_JOIN:
        l, r = tree_task.depends_on
        db.create_task(job_id, name, streams["join"], wire.task_type_to_value(wire.JoinReq(tree_task.task_number, l, r)),
                       [str(l), str(r)], a.join_retries, a.join_timeout)
    elif tree_task.command == CMD_UNION:
        l, r = tree_task.keccak_depends_on
        db.create_task(job_id, name, streams["union"], wire.task_type_to_value(wire.UnionReq(tree_task.task_number, l, r)),
                       [str(l), str(r)], a.join_retries, a.join_timeout)
    elif tree_task.command == CMD_KECCAK:
        raise TaskError("[BENTO-EXEC-001] keccak_req returned None")          # no keccak coprocessor on the synthetic path
    elif tree_task.command == CMD_FINALIZE:
        keccak_count = 1 if tree_task.keccak_depends_on else 0
        assumption_count = len(assumptions) + keccak_count
        prereqs = [str(tree_task.depends_on[0])]
        union_max_idx = None
        if tree_task.keccak_depends_on:
            prereqs.append(str(tree_task.keccak_depends_on[0]))
            union_max_idx = tree_task.keccak_depends_on[0]
        db.create_task(job_id, "resolve", streams["join"],
                       wire.task_type_to_value(wire.ResolveReq(tree_task.depends_on[0], union_max_idx)), prereqs, a.resolve_retries,
                       a.resolve_timeout * assumption_count)
        db.create_task(job_id, "finalize", streams["aux"], wire.task_type_to_value(wire.Finalize()), ["resolve"], a.finalize_retries,
                       a.finalize_timeout)


def executor(agent: Agent, job_id: str, request: wire.ExecutorReq) -> ExecutorResp:
    """Synthetic stand-in for the RISC-V executor: `request.input` names a hot-store key holding
    {"segments": n, "po2": p, "seed_base": s}; segments are flushed to `job:{id}:segments:{i}` and planned online."""
    spec = json.loads(_ctx("input not found: %s" % request.input, agent.hot_get_bytes, request.input))
    n, po2 = int(spec["segments"]), int(spec.get("po2", agent.args.segment_po2))
    seed_base = int(spec.get("seed_base", 0xB2000000))
    if n <= 0:
        raise TaskError("[BENTO-EXEC-030] executor produced no segments")
    db = agent.task_db
    prove_stream = db.get_stream(request.user_id, wire.PROVE_WORK_TYPE)
    aux_stream = db.get_stream(request.user_id, wire.AUX_WORK_TYPE)
    if aux_stream is None:
        raise TaskError("Customer %s missing aux stream" % request.user_id)
    if prove_stream is None:
        raise TaskError("Customer %s missing gpu prove stream" % request.user_id)
    import os
    a = agent.args
    join_stream = union_stream = prove_stream
    if a.join_stream if a.join_stream is not None else "JOIN_STREAM" in os.environ:
        join_stream = db.get_stream(request.user_id, wire.JOIN_WORK_TYPE)
        if join_stream is None:
            raise TaskError("Customer %s missing gpu join stream" % request.user_id)
    if a.union_stream if a.union_stream is not None else "UNION_STREAM" in os.environ:
        union_stream = db.get_stream(request.user_id, wire.JOIN_WORK_TYPE)
        if union_stream is None:
            raise TaskError("Customer %s missing gpu union stream" % request.user_id)
    streams = {"prove": prove_stream, "join": join_stream, "union": union_stream, "aux": aux_stream}
    job_prefix = "job:%s" % job_id
    planner = Planner()
    for i in range(n):
        seg = Segment(index=i, po2=po2, seed=seed_base + i)
        if i == 0:
            seg.assumptions = list(request.assumptions)          # the guest's assumptions surface in its output
        agent.hot_set_bytes("%s:%s:%d" % (job_prefix, wire.SEGMENTS_PATH, i), wire.serialize_segment(seg))
        planner.enqueue_segment()
        while True:                                              # drain the plan as the executor does after every segment
            t = planner.next_task()
            if t is None:
                break
            _process_task(agent, streams, job_id, t, i if t.command == CMD_SEGMENT else None, request.assumptions)
    planner.finish()
    while True:
        t = planner.next_task()
        if t is None:
            break
        _process_task(agent, streams, job_id, t, None, request.assumptions)
    agent.hot_set_bytes("%s:journal" % job_prefix, wire.serialize_journal(json.dumps({"segments": n}).encode()))
    agent.hot_set_bytes("%s:image_id" % job_prefix, request.image.encode())
    cycles = n << po2
    return ExecutorResp(segments=n, user_cycles=cycles, total_cycles=cycles, assumption_count=len(request.assumptions))


# ---- Agent::process_work / poll_work -------------------------------------------------------------------------------------------
def process_work(agent: Agent, task: ReadyTask) -> None:
    try:
        task_type = wire.task_type_from_value(task.task_def)
    except wire.WireError as e:
        raise TaskError("Invalid task_def: %s:%s" % (task.job_id, task.task_id), e)
    cleanup: List[str] = []
    if isinstance(task_type, wire.ExecutorReq):
        res = _ctx("[BENTO-WF-113] Executor failed", executor, agent, task.job_id, task_type).__dict__
    elif isinstance(task_type, wire.ProveReq):
        cleanup = _ctx("[BENTO-WF-115] Prove failed", prover, agent, task.job_id, task.task_id, task_type)
        res = None
    elif isinstance(task_type, wire.JoinReq):
        if agent.is_povw_enabled():       # lib.rs:710-716
            cleanup = _ctx("[BENTO-WF-117] POVW join failed", join_povw, agent, task.job_id, task_type)
        else:
            cleanup = _ctx("[BENTO-WF-119] Join failed", join, agent, task.job_id, task_type)
        res = None
    elif isinstance(task_type, wire.ResolveReq):
        if agent.is_povw_enabled():       # lib.rs:726-734
            res, cleanup = _ctx("[BENTO-WF-121] POVW resolve failed", resolve_povw, agent, task.job_id, task_type)
        else:
            res, cleanup = _ctx("[BENTO-WF-123] Resolve failed", resolver, agent, task.job_id, task_type)
    elif isinstance(task_type, wire.Finalize):
        cleanup = _ctx("[BENTO-WF-125] Finalize failed", finalize, agent, task.job_id)
        res = None
    elif isinstance(task_type, wire.UnionReq):
        cleanup = _ctx("[BENTO-WF-131] Union failed", union, agent, task.job_id, task_type)
        res = None
    elif isinstance(task_type, wire.SnarkReq):
        raise TaskError("[BENTO-WF-127] Snark failed", "stark2snark is out of scope for this agent")
    else:
        cleanup = _ctx("[BENTO-WF-129] Keccak failed", keccak, agent, task.job_id, task.task_id, task_type)
        res = None
    agent.task_db.update_task_done(task.job_id, task.task_id, res)
    # best-effort cleanup only AFTER the task is marked done, so that a retry never finds its inputs missing (lib.rs:781-797)
    for key in cleanup:
        try:
            agent.hot_delete(key)
        except Exception:               # noqa: BLE001
            pass
    agent.processed.append("%s|%s" % (task.job_id, task.task_id))


def poll_work(agent: Agent, max_tasks: Optional[int] = None) -> int:
    """Claim and process until the stream is empty (the reference loops until SIGTERM); returns the number of tasks claimed.
    Failure handling is lib.rs:639-677: retry while retries remain, else fail the task (and with it the job) with the error
    string truncated to 1024 characters, prefixed "retry max hit: " when the retry budget is what ran out."""
    claimed = 0
    db = agent.task_db
    while max_tasks is None or claimed < max_tasks:
        task = db.request_work(agent.args.task_stream)
        if task is None:
            break
        claimed += 1
        try:
            process_work(agent, task)
            continue
        except Exception as err:          # noqa: BLE001
            err_str = str(err)
            agent.errors.append("%s|%s: %s" % (task.job_id, task.task_id, err_str))
        if task.max_retries > 0:
            current = db.get_task_retries_running(task.job_id, task.task_id)
            if current is not None and current + 1 > task.max_retries:
                err_str = err_str[:1024]
                final = "retry max hit" if not err_str else "retry max hit: %s" % err_str
                db.update_task_failed(task.job_id, task.task_id, final)
                continue
            db.update_task_retry(task.job_id, task.task_id)
        else:
            db.update_task_failed(task.job_id, task.task_id, err_str[:1024])
    return claimed


# ---- the same loop with several Prove claims in flight on one GPU ------------------------------------------------------------------
def _fail_or_retry(agent: Agent, task: ReadyTask, err_str: str) -> None:
    """lib.rs:639-677 (shared by poll_work_pipelined): retry while the budget lasts, else fail the task and with it the job."""
    db = agent.task_db
    agent.errors.append("%s|%s: %s" % (task.job_id, task.task_id, err_str))
    if task.max_retries > 0:
        current = db.get_task_retries_running(task.job_id, task.task_id)
        if current is not None and current + 1 > task.max_retries:
            err_str = err_str[:1024]
            db.update_task_failed(task.job_id, task.task_id, "retry max hit" if not err_str else "retry max hit: %s" % err_str)
            return
        db.update_task_retry(task.job_id, task.task_id)
    else:
        db.update_task_failed(task.job_id, task.task_id, err_str[:1024])


def poll_work_pipelined(agent: Agent, max_tasks: Optional[int] = None) -> int:
    """Agent::poll_work for a GPU worker whose prover has several proof slots: up to `slots` Prove tasks are claimed and in flight at
    once, each one the whole body of tasks::prove::prover (prove_segment -> verify_integrity -> lift -> verify_integrity, one enqueue:
    ProverServer.submit_prove_lift), and a task is marked done only after its lifted receipt is in the hot store -- the reference's
    ordering (tasks/prove.rs:113-117, lib.rs:781-797).  The reference reaches the same concurrency by running several agent processes;
    one process per GPU with several slots avoids re-loading tables and contexts.  Any other task type is processed by the plain
    synchronous path once the proofs in flight have drained.  Returns the number of tasks claimed."""
    p = agent._prover("[BENTO-PROVE-002] Missing prover from prove task")
    db = agent.task_db
    free = list(range(p.opts.slots))
    inflight = {}                          # slot -> (task, segment_key, output_key)
    claimed = 0

    def finish(slot):
        task, segment_key, output_key = inflight.pop(slot)
        free.append(slot)
        try:
            try:
                _seg, lift = p.wait_task(slot)
            except Exception as e:        # noqa: BLE001
                what = getattr(e, "what", "")
                tag = ("[BENTO-PROVE-004] Failed to verify segment receipt integrity" if what == "segment receipt" else
                       "[BENTO-PROVE-010] Failed to verify lift receipt integrity" if what == "lift receipt" else "[BENTO-PROVE-003] prove failed")
                raise TaskError("[BENTO-WF-115] Prove failed", TaskError(tag, e))
            receipt = SuccinctReceipt(lift.seal, lift.kind, tuple(lift.claim), list(lift.assumptions))
            blob = _ctx("Failed to serialize the segment", wire.serialize_succinct, receipt)
            _ctx("Failed to set receipt key with expiry", agent.hot_set_bytes, output_key, blob)
        except Exception as err:          # noqa: BLE001
            _fail_or_retry(agent, task, str(err))
            return
        db.update_task_done(task.job_id, task.task_id, None)
        try:
            agent.hot_delete(segment_key)
        except Exception:                 # noqa: BLE001
            pass
        agent.processed.append("%s|%s" % (task.job_id, task.task_id))

    while True:
        task = None
        if free and (max_tasks is None or claimed < max_tasks):
            task = db.request_work(agent.args.task_stream)
        if task is None:
            if not inflight:
                break
            done = [s for s in inflight if p.query(s)]
            for s in done or [next(iter(inflight))]:      # nothing finished yet: block on the oldest
                finish(s)
            continue
        claimed += 1
        try:
            task_type = wire.task_type_from_value(task.task_def)
        except wire.WireError as e:
            _fail_or_retry(agent, task, str(TaskError("Invalid task_def: %s:%s" % (task.job_id, task.task_id), e)))
            continue
        if not isinstance(task_type, wire.ProveReq) or agent.is_povw_enabled():
            for s in list(inflight):
                finish(s)
            try:
                process_work(agent, task)
            except Exception as err:      # noqa: BLE001
                _fail_or_retry(agent, task, str(err))
            continue
        try:
            segment_key = "job:%s:%s:%d" % (task.job_id, wire.SEGMENTS_PATH, task_type.index)
            segment_vec = _ctx("segment data not found for segment key: %s" % segment_key, agent.hot_get_bytes, segment_key)
            segment = _ctx("Failed to deserialize segment data from redis", wire.deserialize_segment, segment_vec)
            slot = free.pop(0)
            try:
                p.submit_prove_lift(slot, segment)
            except Exception:
                free.insert(0, slot)
                raise
            inflight[slot] = (task, segment_key, "job:%s:%s:%s" % (task.job_id, wire.RECUR_RECEIPT_PATH, task.task_id))
        except Exception as err:          # noqa: BLE001
            _fail_or_retry(agent, task, str(TaskError("[BENTO-WF-115] Prove failed", err)))
    return claimed
