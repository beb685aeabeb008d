"""Host logic either side of the proving path, on the CPU with a fake prover: wire formats (workflow-common), taskdb scheduling
semantics (redis_backend.rs Lua), the task functions (tasks/prove.rs, join.rs, union.rs, resolve.rs, finalize.rs) and the
agent's dispatch / retry loop (workflow/src/lib.rs:611-799).  The GPU twin with real proofs is tests/test_gpu_tasks.py."""
import hashlib
import json

import numpy as np
import pytest

from boundless_b200 import tasks, wire
from boundless_b200.prover_server import (KIND_JOIN, KIND_LIFT, KIND_RESOLVE, KIND_UNION, Segment, SegmentReceipt, SuccinctReceipt,
                                          VerificationError)
from boundless_b200.taskdb import INIT_TASK, MemoryTaskDb, TaskDbError

IMAGE = "ab" * 32


def _seal(tag, *parts):
    h = hashlib.sha256(repr((tag,) + parts).encode()).digest()
    return np.frombuffer(h * 4, dtype=np.uint32).copy()


class FakeProver:
    """Deterministic stand-in with the ProverServer method names; a seal whose first word is 0xBAD fails verification."""

    def __init__(self):
        self.calls = []
        self.fail_next = {}

    def _maybe_fail(self, what):
        if self.fail_next.get(what, 0) > 0:
            self.fail_next[what] -= 1
            raise RuntimeError("injected %s failure" % what)

    def prove_segment(self, ctx, segment):
        self._maybe_fail("prove_segment")
        self.calls.append(("prove_segment", segment.index))
        return SegmentReceipt(_seal("seg", segment.index, segment.seed, segment.po2), segment.index, segment.po2, list(segment.assumptions))

    def lift(self, r):
        self.calls.append(("lift", r.index))
        return SuccinctReceipt(_seal("lift", r.seal.tobytes()), KIND_LIFT, (r.index, r.index), list(r.assumptions))

    def join(self, a, b):
        self._maybe_fail("join")
        self.calls.append(("join", a.claim, b.claim))
        assert a.claim[1] + 1 == b.claim[0], "join of non-adjacent receipts"
        return SuccinctReceipt(_seal("join", a.seal.tobytes(), b.seal.tobytes()), KIND_JOIN, (a.claim[0], b.claim[1]),
                               list(a.assumptions) + list(b.assumptions))

    def union(self, a, b):
        self.calls.append(("union",))
        return SuccinctReceipt(_seal("union", a.seal.tobytes(), b.seal.tobytes()), KIND_UNION, (0, 0))

    def resolve(self, cond, asm):
        self.calls.append(("resolve", asm.claim_digest()))
        gone = asm.claim_digest()
        assert gone in cond.assumptions
        return SuccinctReceipt(_seal("resolve", cond.seal.tobytes(), asm.seal.tobytes()), KIND_RESOLVE, tuple(cond.claim),
                               [x for x in cond.assumptions if x != gone])

    def verify_integrity(self, receipt, slot=0):
        self.calls.append(("verify",))
        if int(receipt.seal[0]) == 0xBAD:
            raise VerificationError(120)


# ---- wire ------------------------------------------------------------------------------------------------------------------
def test_task_type_json_matches_serde():
    """Externally tagged enums, unit variant as a bare string (workflow-common/src/lib.rs:150-168)."""
    assert wire.task_type_to_json(wire.ProveReq(3)) == '{"Prove":{"index":3}}'
    assert wire.task_type_to_json(wire.JoinReq(5, 1, 2)) == '{"Join":{"idx":5,"left":1,"right":2}}'
    assert wire.task_type_to_json(wire.ResolveReq(100, 50)) == '{"Resolve":{"max_idx":100,"union_max_idx":50}}'
    assert wire.task_type_to_json(wire.ResolveReq(7)) == '{"Resolve":{"max_idx":7,"union_max_idx":null}}'
    assert wire.task_type_to_json(wire.Finalize()) == '"Finalize"'
    assert wire.task_type_to_json(wire.UnionReq(9, 3, 4)) == '{"Union":{"idx":9,"left":3,"right":4}}'
    for t in (wire.ProveReq(0), wire.JoinReq(2, 0, 1), wire.UnionReq(2, 0, 1), wire.ResolveReq(4, None), wire.Finalize(),
              wire.SnarkReq("abc", "Groth16"), wire.ExecutorReq("img", "inp", "user", ["a"], False, "None", 5)):
        assert wire.task_type_from_json(wire.task_type_to_json(t)) == t
    # the reference's own unit test (workflow-common/src/lib.rs:190-224)
    r = wire.ResolveReq(max_idx=100, union_max_idx=50)
    assert r.max_idx == 100 and r.union_max_idx == 50
    names = [wire.to_job_type_str(t) for t in (wire.ExecutorReq("i", "j", "u"), wire.ProveReq(1), wire.JoinReq(1, 2, 3), wire.ResolveReq(1),
                                               wire.Finalize(), wire.SnarkReq("r"), wire.KeccakReq("00" * 32, 16, "00" * 32), wire.UnionReq(1, 2, 3))]
    assert names == ["executor", "prove-lift", "join", "resolve", "finalize", "snark", "keccak", "union"]


_EXEC = '"image":"a","input":"b","user_id":"c","assumptions":[],"execute_only":false'


@pytest.mark.parametrize("bad", ['{"Prove":{}}', '{"Prove":{"index":-1}}', '{"Nope":{}}', '"Prove"', "[1]",
                                 '{"Prove":{"index":1},"Join":{}}', "{", '{"Executor":{%s,"compress":"Zip"}}' % _EXEC,
                                 '{"Executor":{"image":"a","input":"b","user_id":"c"}}',                  # non-Option fields are required
                                 '{"Snark":{"receipt":"r"}}', '{"Snark":{"receipt":"r","compress_type":"Zip"}}',
                                 '{"Keccak":{"claim_digest":[0,0],"po2":16,"control_root":"%s"}}' % ("00" * 32)])
def test_task_type_rejects_malformed(bad):
    with pytest.raises(wire.WireError):
        wire.task_type_from_json(bad)


def test_task_type_follows_serde_defaults():
    """serde ignores unknown fields (no deny_unknown_fields), requires every non-Option field, treats a missing Option as
    None; risc0 Digests are hex strings in JSON."""
    assert wire.task_type_from_json('{"Prove":{"index":1,"added_by_a_newer_control_plane":2}}') == wire.ProveReq(1)
    e = wire.task_type_from_json('{"Executor":{%s,"compress":"Groth16","future":true}}' % _EXEC)
    assert e == wire.ExecutorReq("a", "b", "c", [], False, "Groth16", None)
    assert wire.task_type_from_json('{"Resolve":{"max_idx":3}}') == wire.ResolveReq(3, None)
    k = wire.task_type_from_json('{"Keccak":{"claim_digest":"%s","po2":17,"control_root":"%s"}}' % ("ab" * 32, "cd" * 32))
    assert k == wire.KeccakReq("ab" * 32, 17, "cd" * 32)
    assert wire.task_type_from_json('{"Snark":{"receipt":"r","compress_type":"Blake3Groth16"}}') == wire.SnarkReq("r", "Blake3Groth16")


def test_bincode_blobs_roundtrip_and_layout():
    seg = Segment(index=3, po2=12, seed=0xB2000003)
    blob = wire.serialize_segment(seg)
    # bincode 1.x: u64 index, u32 po2, u64 seed, Option tag 0, u64 len 0
    assert blob == (3).to_bytes(8, "little") + (12).to_bytes(4, "little") + (0xB2000003).to_bytes(8, "little") + b"\x00" + bytes(8)
    back = wire.deserialize_segment(blob)
    assert (back.index, back.po2, back.seed, back.trace, back.assumptions) == (3, 12, 0xB2000003, None, [])
    seg2 = Segment(index=1, po2=9, seed=5, trace=np.arange(7, dtype=np.uint32), assumptions=["ff" * 32])
    b2 = wire.deserialize_segment(wire.serialize_segment(seg2))
    assert np.array_equal(b2.trace, seg2.trace) and b2.assumptions == ["ff" * 32]
    rc = SuccinctReceipt(np.arange(100, dtype=np.uint32), KIND_JOIN, (2, 5), ["aa", "bb"])
    b3 = wire.deserialize_succinct(wire.serialize_succinct(rc))
    assert np.array_equal(b3.seal, rc.seal) and (b3.kind, b3.claim, b3.assumptions) == (KIND_JOIN, (2, 5), ["aa", "bb"])
    roll = wire.serialize_rollup(rc, b"journal-bytes")
    assert roll[:4] == (1).to_bytes(4, "little")                      # InnerReceipt::Succinct
    r4, j4 = wire.deserialize_rollup(roll)
    assert np.array_equal(r4.seal, rc.seal) and j4 == b"journal-bytes"
    for broken in (blob[:-1], blob + b"\x00", b"", roll[:10]):
        with pytest.raises(wire.WireError):
            (wire.deserialize_rollup if broken is roll[:10] else wire.deserialize_segment)(broken)
    with pytest.raises(wire.WireError):
        wire.deserialize_rollup((0).to_bytes(4, "little") + roll[4:])


# ---- taskdb ----------------------------------------------------------------------------------------------------------------
def _db():
    db = MemoryTaskDb()
    prove = db.create_stream(wire.PROVE_WORK_TYPE, user_id="u")
    aux = db.create_stream(wire.AUX_WORK_TYPE, user_id="u")
    execs = db.create_stream(wire.EXEC_WORK_TYPE, user_id="u")
    return db, prove, aux, execs


def test_taskdb_ready_order_and_dependencies():
    db, prove, aux, execs = _db()
    job = db.create_job(execs, {"init": 1}, user_id="u")
    assert db.request_work(wire.EXEC_WORK_TYPE).task_id == INIT_TASK and db.request_work(wire.EXEC_WORK_TYPE) is None
    db.create_task(job, "0", prove, "a", [], 1, 10)
    db.create_task(job, "1", prove, "b", [], 1, 10)
    db.create_task(job, "2", prove, "join", ["0", "1"], 1, 10)          # created before segment 3 => claimed before it
    db.create_task(job, "3", prove, "c", [], 1, 10)
    assert db.task_state(job, "2") == "pending" and db.task_field(job, "2", "waiting_on") == 2
    with pytest.raises(TaskDbError, match="missing prerequisite task: 9"):
        db.create_task(job, "4", prove, "x", ["9"], 1, 10)
    with pytest.raises(TaskDbError, match="task already exists: 3"):
        db.create_task(job, "3", prove, "x", [], 1, 10)
    t0 = db.request_work(wire.PROVE_WORK_TYPE); t1 = db.request_work(wire.PROVE_WORK_TYPE)
    assert (t0.task_id, t1.task_id) == ("0", "1")
    db.update_task_done(job, "0"); assert db.task_field(job, "2", "waiting_on") == 1
    db.update_task_done(job, "1"); assert db.task_state(job, "2") == "ready"
    assert db.request_work(wire.PROVE_WORK_TYPE).task_id == "2"           # the join jumps ahead of the later segment
    assert db.request_work(wire.PROVE_WORK_TYPE).task_id == "3"
    assert not db.update_task_done(job, "0")                               # already done
    # a prerequisite that is already done does not block
    db.create_task(job, "5", prove, "late", ["0"], 0, 10)
    assert db.task_state(job, "5") == "ready"


def test_taskdb_priority_retry_timeout_and_failure():
    now = [1000.0]
    db = MemoryTaskDb(clock=lambda: now[0])
    prove = db.create_stream(wire.PROVE_WORK_TYPE, user_id="u")
    execs = db.create_stream(wire.EXEC_WORK_TYPE, user_id="u")
    low = db.create_job(execs, None, user_id="u", priority=2)
    high = db.create_job(execs, None, user_id="u", priority=0)
    db.create_task(low, "a", prove, 1, [], 1, 30)
    db.create_task(high, "b", prove, 2, [], 1, 30)
    assert db.request_work(wire.PROVE_WORK_TYPE).job_id == high            # priority 0 before priority 2, despite creation order
    t = db.request_work(wire.PROVE_WORK_TYPE)
    assert t.job_id == low and db.get_task_retries_running(low, "a") == 0
    assert db.requeue_tasks() == 0
    now[0] += 31                                                             # both claims expire
    assert db.requeue_tasks() == 2 and db.task_state(low, "a") == "ready" and db.task_field(low, "a", "retries") == 1
    t = db.request_work(wire.PROVE_WORK_TYPE); t = db.request_work(wire.PROVE_WORK_TYPE)
    assert not db.update_task_retry(low, "a")                                # retries 2 > max 1
    assert db.task_state(low, "a") == "failed" and db.job_state(low) == "failed" and db.job_error(low) == "retry max hit"
    # failing a task fails the job and cancels what has not started
    db.create_task(high, "c", prove, 3, ["b"], 1, 30)
    db.create_task(high, "d", prove, 4, [], 1, 30)
    assert db.update_task_failed(high, "b", "boom")
    assert db.task_state(high, "c") == "cancelled" and db.task_state(high, "d") == "cancelled"
    assert db.request_work(wire.PROVE_WORK_TYPE) is None


# ---- the job, end to end ------------------------------------------------------------------------------------------------------
def _run_job(n_segments, prover=None, assumptions=(), args=None):
    db, prove, aux, execs = _db()
    store = tasks.MemoryHotStore()
    prover = prover or FakeProver()
    store.set_bytes("input:1", json.dumps({"segments": n_segments, "po2": 10}).encode())
    req = wire.ExecutorReq(image=IMAGE, input="input:1", user_id="u", assumptions=list(assumptions))
    job = db.create_job(execs, wire.task_type_to_value(req), user_id="u")
    exec_agent = tasks.Agent(db, store, None, tasks.AgentArgs(task_stream=wire.EXEC_WORK_TYPE, segment_po2=10))
    gpu_agent = tasks.Agent(db, store, prover, args or tasks.AgentArgs(task_stream=wire.PROVE_WORK_TYPE))
    aux_agent = tasks.Agent(db, store, prover, tasks.AgentArgs(task_stream=wire.AUX_WORK_TYPE))
    return db, store, job, prover, exec_agent, gpu_agent, aux_agent


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 13])
def test_job_reduces_to_one_verified_receipt(n):
    db, store, job, prover, exec_agent, gpu_agent, aux_agent = _run_job(n)
    assert tasks.poll_work(exec_agent) == 1
    assert db.task_field(job, INIT_TASK, "output")["segments"] == n
    assert sum(1 for k in store.kv if ":segments:" in k) == n
    assert tasks.poll_work(aux_agent) == 0                                   # finalize is still pending
    claimed = tasks.poll_work(gpu_agent)
    assert claimed == n + (n - 1) + 1                                        # proves, joins, resolve
    assert tasks.poll_work(aux_agent) == 1 and db.job_state(job) == "done"
    # every intermediate blob was cleaned up after its consumer was marked done; the final receipt is in shared storage
    assert not [k for k in store.kv if k.startswith("job:")], sorted(store.kv)
    (key, blob), = store.assets.items()
    assert key == "receipts/stark/%s.bincode" % job
    root, journal = wire.deserialize_rollup(blob)
    assert root.claim == (0, n - 1) and json.loads(journal) == {"segments": n}
    # prove -> verify -> lift -> verify per segment; verify, verify, join, verify per join
    assert prover.calls.count(("verify",)) == 2 * n + 3 * (n - 1) + 1
    # depth-first reduction: a join runs as soon as both inputs exist, before later segments (planner/mod.rs:100-113)
    order = [c for c in prover.calls if c[0] in ("prove_segment", "join")]
    if n >= 3:
        assert order[:3] == [("prove_segment", 0), ("prove_segment", 1), ("join", (0, 0), (1, 1))]


def test_two_gpu_agents_share_the_queue():
    """compose.yml:113: one agent per GPU pulling from the same stream; any interleaving must give the same root."""
    db, store, job, prover, exec_agent, a0, aux_agent = _run_job(9)
    a1 = tasks.Agent(db, store, prover, tasks.AgentArgs(task_stream=wire.PROVE_WORK_TYPE))
    tasks.poll_work(exec_agent)
    turn = 0
    while tasks.poll_work((a0, a1)[turn % 2], max_tasks=1 + turn % 3):
        turn += 1
    tasks.poll_work(aux_agent)
    assert db.job_state(job) == "done" and a0.processed and a1.processed
    root, _ = wire.deserialize_rollup(next(iter(store.assets.values())))
    db2, store2, job2, _, e2, g2, x2 = _run_job(9)
    tasks.poll_work(e2); tasks.poll_work(g2); tasks.poll_work(x2)
    root2, _ = wire.deserialize_rollup(next(iter(store2.assets.values())))
    assert np.array_equal(root.seal, root2.seal)


def test_assumptions_are_resolved_before_finalize():
    fp = FakeProver()
    asm = SuccinctReceipt(_seal("assumption"), KIND_LIFT, (0, 0))
    claim = asm.claim_digest()
    db, store, job, prover, exec_agent, gpu_agent, aux_agent = _run_job(3, fp, assumptions=[claim])
    store.set_bytes("job:%s:receipts:%s" % (job, claim), wire.serialize_succinct(asm))
    tasks.poll_work(exec_agent); tasks.poll_work(gpu_agent)
    assert db.task_field(job, "resolve", "output") == 1 and ("resolve", claim) in fp.calls
    assert db.task_field(job, "resolve", "timeout_secs") == tasks.AgentArgs().resolve_timeout * 1
    tasks.poll_work(aux_agent)
    assert db.job_state(job) == "done"
    root, _ = wire.deserialize_rollup(next(iter(store.assets.values())))
    assert root.kind == KIND_RESOLVE and root.assumptions == []
    # a missing corroborating receipt fails the resolve task with the reference's context chain
    db, store, job, prover, exec_agent, gpu_agent, aux_agent = _run_job(2, FakeProver(), assumptions=[claim])
    tasks.poll_work(exec_agent); tasks.poll_work(gpu_agent)
    assert db.job_state(job) == "failed"
    assert db.job_error(job).startswith("retry max hit: [BENTO-WF-123] Resolve failed: corroborating receipt not found: key job:")


def test_failures_retry_then_fail_with_truncated_error():
    fp = FakeProver(); fp.fail_next["join"] = 1                             # one transient failure: retried, job completes
    db, store, job, prover, exec_agent, gpu_agent, aux_agent = _run_job(2, fp)
    tasks.poll_work(exec_agent); tasks.poll_work(gpu_agent); tasks.poll_work(aux_agent)
    assert db.job_state(job) == "done" and db.task_field(job, "2", "retries") == 1
    assert "job:%s:recursion_receipts:0" % job not in store.kv              # inputs survived the failed attempt, deleted after success
    fp = FakeProver(); fp.fail_next["prove_segment"] = 99                   # permanent failure
    db, store, job, prover, exec_agent, gpu_agent, aux_agent = _run_job(3, fp)
    tasks.poll_work(exec_agent); tasks.poll_work(gpu_agent)
    assert db.job_state(job) == "failed" and db.task_field(job, "0", "retries") == 3
    assert db.job_error(job) == "retry max hit: [BENTO-WF-115] Prove failed: injected prove_segment failure"
    assert db.task_state(job, "finalize") == "cancelled" and "job:%s:segments:0" % job in store.kv
    # max_retries == 0: fail at once, message truncated to 1024 characters
    fp = FakeProver(); fp.fail_next["prove_segment"] = 1
    db, store, job, prover, exec_agent, gpu_agent, aux_agent = _run_job(1, fp, args=tasks.AgentArgs(prove_retries=0))
    exec_agent.args.prove_retries = 0
    store.set_bytes("input:1", json.dumps({"segments": 1, "po2": 10}).encode())
    tasks.poll_work(exec_agent)
    fp._maybe_fail = lambda what: (_ for _ in ()).throw(RuntimeError("x" * 5000))
    tasks.poll_work(gpu_agent)
    assert db.job_state(job) == "failed" and len(db.job_error(job)) == 1024


def test_bad_receipts_and_blobs_are_caught():
    class Forger(FakeProver):
        def lift(self, r):
            out = super().lift(r); out.seal[0] = 0xBAD; return out
    db, store, job, prover, exec_agent, gpu_agent, aux_agent = _run_job(1, Forger())
    tasks.poll_work(exec_agent); tasks.poll_work(gpu_agent)
    assert "[BENTO-PROVE-010] Failed to verify lift receipt integrity: seal does not verify (check 120 failed)" in db.job_error(job)
    # corrupt stored receipt -> join reports which side failed to deserialize
    db, store, job, prover, exec_agent, gpu_agent, aux_agent = _run_job(2)
    tasks.poll_work(exec_agent); tasks.poll_work(gpu_agent, max_tasks=2)
    store.set_bytes("job:%s:recursion_receipts:1" % job, b"\x01\x02")
    tasks.poll_work(gpu_agent)
    assert "[BENTO-WF-119] Join failed: [BENTO-JOIN-002] Failed to deserialize right receipt" in db.job_error(job)
    # an agent without a prover cannot take GPU work; an unknown task_def is "Invalid task_def"
    db, store, job, prover, exec_agent, gpu_agent, aux_agent = _run_job(1)
    tasks.poll_work(exec_agent)
    gpu_agent.prover = None
    tasks.poll_work(gpu_agent)
    assert "[BENTO-PROVE-002] Missing prover from prove task" in db.job_error(job)
    db, prove, aux, execs = _db()
    job = db.create_job(execs, {"Bogus": {}}, user_id="u")
    tasks.poll_work(tasks.Agent(db, tasks.MemoryHotStore(), None, tasks.AgentArgs(task_stream=wire.EXEC_WORK_TYPE)))
    assert db.job_error(job).startswith("Invalid task_def: %s:init: unknown variant `Bogus`" % job)


# ---- several Prove claims in flight on one GPU (poll_work_pipelined) -----------------------------------------------------------------
class SlottedFakeProver(FakeProver):
    """FakeProver with the asynchronous composite-task surface of ProverServer: submit_prove_lift / query / wait_task over `slots`."""

    class _Opts:
        def __init__(self, slots): self.slots = slots

    def __init__(self, slots=3):
        super().__init__()
        self.opts = self._Opts(slots)
        self.slot_job, self.polls, self.max_inflight = {}, {}, 0

    def submit_prove_lift(self, slot, segment, d_out=0, verify=True, host_seals=True):
        assert slot not in self.slot_job, "slot reused while busy"
        self._maybe_fail("submit")
        self.slot_job[slot] = segment
        self.polls[slot] = 2 + segment.index % 3                       # completes a few polls later, out of order
        self.max_inflight = max(self.max_inflight, len(self.slot_job))

    def query(self, slot):
        self.polls[slot] -= 1
        return self.polls[slot] <= 0

    def wait_task(self, slot):
        from boundless_b200.prover_server import DeviceReceipt
        segment = self.slot_job.pop(slot)
        seg_r = self.prove_segment(None, segment)
        if int(seg_r.seal[0]) == 0xBAD or self.fail_next.get("verify_segment", 0) > 0:
            self.fail_next["verify_segment"] = self.fail_next.get("verify_segment", 1) - 1
            raise VerificationError(120, "segment receipt")
        lift = self.lift(seg_r)
        return seg_r, DeviceReceipt(0, lift.seal.size, lift.kind, lift.claim, list(lift.assumptions), None, lift.seal)


@pytest.mark.parametrize("n", [1, 4, 9])
def test_pipelined_agent_gives_the_same_job_result(n):
    """The pipelined loop claims in the same (priority, creation) order, keeps `slots` proofs in flight, stores each lifted receipt
    before marking its task done, and the job reduces to the same root as with the synchronous loop."""
    db, store, job, prover, exec_agent, gpu_agent, aux_agent = _run_job(n, prover=SlottedFakeProver(3))
    tasks.poll_work(exec_agent)
    claimed = tasks.poll_work_pipelined(gpu_agent)
    assert claimed == n + (n - 1) + 1 and gpu_agent.errors == []
    assert prover.max_inflight == min(3, n)
    tasks.poll_work(aux_agent)
    assert db.job_state(job) == "done"
    assert not [k for k in store.kv if k.startswith("job:")]
    root, _ = wire.deserialize_rollup(next(iter(store.assets.values())))
    db2, store2, job2, _, e2, g2, x2 = _run_job(n)
    tasks.poll_work(e2); tasks.poll_work(g2); tasks.poll_work(x2)
    root2, _ = wire.deserialize_rollup(next(iter(store2.assets.values())))
    assert np.array_equal(root.seal, root2.seal) and root.claim == root2.claim == (0, n - 1)


def test_pipelined_agent_retries_and_fails_like_the_reference():
    """A failed verification inside a pipelined Prove task goes through the same retry / "retry max hit" rules (lib.rs:639-677)."""
    p = SlottedFakeProver(2)
    db, store, job, prover, exec_agent, gpu_agent, aux_agent = _run_job(3, prover=p)
    tasks.poll_work(exec_agent)
    p.fail_next["verify_segment"] = 1                                  # one transient failure: retried, job still completes
    tasks.poll_work_pipelined(gpu_agent)
    assert len(gpu_agent.errors) == 1 and "[BENTO-WF-115] Prove failed: [BENTO-PROVE-004]" in gpu_agent.errors[0]
    tasks.poll_work(aux_agent)
    assert db.job_state(job) == "done"
    p2 = SlottedFakeProver(2)
    db, store, job, prover, exec_agent, gpu_agent, aux_agent = _run_job(2, prover=p2)
    tasks.poll_work(exec_agent)
    p2.fail_next["verify_segment"] = 100                               # permanent: the task exhausts its retries and fails the job
    tasks.poll_work_pipelined(gpu_agent)
    assert db.job_state(job) == "failed" and db.job_error(job).startswith("retry max hit: [BENTO-WF-115] Prove failed")


def test_join_stream_mode_routes_joins_to_the_join_stream():
    """executor.rs:517-525: with JOIN_STREAM set, join / resolve tasks are created on the customer's "join" stream; a worker on the
    prove stream then sees Prove tasks only, one on the join stream the rest."""
    db, prove, aux, execs = _db()
    db.create_stream(wire.JOIN_WORK_TYPE, user_id="u")
    store = tasks.MemoryHotStore()
    p = SlottedFakeProver(2)
    store.set_bytes("input:1", json.dumps({"segments": 4, "po2": 10}).encode())
    job = db.create_job(execs, wire.task_type_to_value(wire.ExecutorReq(image=IMAGE, input="input:1", user_id="u")), user_id="u")
    tasks.poll_work(tasks.Agent(db, store, None, tasks.AgentArgs(task_stream=wire.EXEC_WORK_TYPE, segment_po2=10, join_stream=True)))
    prove_agent = tasks.Agent(db, store, p, tasks.AgentArgs(task_stream=wire.PROVE_WORK_TYPE))
    assert tasks.poll_work_pipelined(prove_agent) == 4 and len(prove_agent.processed) == 4
    assert db.job_state(job) == "running"
    join_agent = tasks.Agent(db, store, p, tasks.AgentArgs(task_stream=wire.JOIN_WORK_TYPE))
    assert tasks.poll_work(join_agent) == 3 + 1                          # three joins and the resolve
    assert tasks.poll_work(tasks.Agent(db, store, p, tasks.AgentArgs(task_stream=wire.AUX_WORK_TYPE))) == 1
    assert db.job_state(job) == "done"
    # missing join stream is the reference's error
    db2, _, _, execs2 = _db()
    store2 = tasks.MemoryHotStore(); store2.set_bytes("input:1", json.dumps({"segments": 2, "po2": 10}).encode())
    job2 = db2.create_job(execs2, wire.task_type_to_value(wire.ExecutorReq(image=IMAGE, input="input:1", user_id="u")), user_id="u")
    tasks.poll_work(tasks.Agent(db2, store2, None, tasks.AgentArgs(task_stream=wire.EXEC_WORK_TYPE, segment_po2=10, join_stream=True)))
    assert db2.job_state(job2) == "failed" and "missing gpu join stream" in db2.job_error(job2)


# ---- proof-of-verifiable-work flow (POVW_LOG_ID set: lib.rs:209-212, :710-734) -----------------------------------------------------------
class PovwFakeProver(FakeProver):
    def lift_povw(self, r):
        self.calls.append(("lift_povw", r.index))
        return SuccinctReceipt(_seal("lift_povw", r.seal.tobytes()), 5, (r.index, r.index), list(r.assumptions))

    def join_povw(self, a, b):
        self._maybe_fail("join_povw")
        self.calls.append(("join_povw", a.claim, b.claim))
        assert a.kind in (5, 6) and b.kind in (5, 6) and a.claim[1] + 1 == b.claim[0]
        return SuccinctReceipt(_seal("join_povw", a.seal.tobytes(), b.seal.tobytes()), 6, (a.claim[0], b.claim[1]),
                               list(a.assumptions) + list(b.assumptions))

    def unwrap_povw(self, r):
        self.calls.append(("unwrap_povw", r.claim))
        assert r.kind in (5, 6)
        return SuccinctReceipt(_seal("unwrap", r.seal.tobytes()), 7, tuple(r.claim), list(r.assumptions))


def test_povw_job_uses_the_povw_programs_end_to_end():
    """With PoVW enabled the prove task lifts with lift_povw, joins are join_povw, resolve unwraps the root first and saves the PoVW
    receipt + metadata to the work-receipts bucket (prove.rs:67-93, join_povw.rs, resolve_povw.rs:214-268)."""
    db, prove, aux, execs = _db()
    store = tasks.MemoryHotStore()
    p = PovwFakeProver()
    n = 5
    store.set_bytes("input:1", json.dumps({"segments": n, "po2": 10}).encode())
    job = db.create_job(execs, wire.task_type_to_value(wire.ExecutorReq(image=IMAGE, input="input:1", user_id="u")), user_id="u")
    tasks.poll_work(tasks.Agent(db, store, None, tasks.AgentArgs(task_stream=wire.EXEC_WORK_TYPE, segment_po2=10)))
    gpu_agent = tasks.Agent(db, store, p, tasks.AgentArgs(task_stream=wire.PROVE_WORK_TYPE, povw_job_number=42), povw="0x" + "ab" * 20)
    assert tasks.poll_work(gpu_agent) == n + (n - 1) + 1 and gpu_agent.errors == []
    tasks.poll_work(tasks.Agent(db, store, p, tasks.AgentArgs(task_stream=wire.AUX_WORK_TYPE)))
    assert db.job_state(job) == "done"
    names = [c[0] for c in p.calls]
    assert names.count("lift_povw") == n and names.count("join_povw") == n - 1 and names.count("unwrap_povw") == 1
    assert "lift" not in names and "join" not in names
    work = wire.deserialize_succinct(store.assets["work_receipts/%s.bincode" % job])
    assert work.kind == 6 and work.claim == (0, n - 1)                    # the PoVW root itself, not the unwrapped one
    meta = json.loads(store.assets["work_receipts/%s_metadata.json" % job])
    assert meta == {"job_id": job, "povw_log_id": "0x" + "ab" * 20, "povw_job_number": "42"}
    root, _ = wire.deserialize_rollup(store.assets["receipts/stark/%s.bincode" % job])
    assert root.kind == 7 and root.claim == (0, n - 1)                    # finalize uploads the unwrapped receipt


def test_povw_join_failures_carry_the_reference_contexts():
    db, prove, aux, execs = _db()
    store = tasks.MemoryHotStore()
    p = PovwFakeProver()
    store.set_bytes("input:1", json.dumps({"segments": 2, "po2": 10}).encode())
    job = db.create_job(execs, wire.task_type_to_value(wire.ExecutorReq(image=IMAGE, input="input:1", user_id="u")), user_id="u")
    tasks.poll_work(tasks.Agent(db, store, None, tasks.AgentArgs(task_stream=wire.EXEC_WORK_TYPE, segment_po2=10, join_retries=0)))
    agent = tasks.Agent(db, store, p, tasks.AgentArgs(task_stream=wire.PROVE_WORK_TYPE), povw=True)
    p.fail_next["join_povw"] = 1
    tasks.poll_work(agent)
    assert db.job_state(job) == "failed"
    assert db.job_error(job).startswith("[BENTO-WF-117] POVW join failed: POVW join method not available")


# ---- keccak coprocessor task (tasks/keccak.rs) -----------------------------------------------------------------------------------
class KeccakFakeProver(FakeProver):
    def prove_keccak(self, claim_digest, po2, control_root, input_states):
        self._maybe_fail("prove_keccak")
        self.calls.append(("prove_keccak", claim_digest, po2, len(input_states)))
        return SuccinctReceipt(_seal("keccak", claim_digest.encode(), bytes(input_states)), 8, (0, 0), [])


def test_keccak_task_then_union():
    """Two Keccak tasks prove their input states, store receipts under keccak_receipts, and the Union task joins them
    (tasks/keccak.rs:25-108, tasks/union.rs); the error contexts are the reference's."""
    db, prove, aux, execs = _db()
    store = tasks.MemoryHotStore()
    p = KeccakFakeProver()
    coproc = db.create_stream(wire.COPROC_WORK_TYPE, user_id="u")
    job = db.create_job(execs, wire.task_type_to_value(wire.ExecutorReq(image=IMAGE, input="x", user_id="u")), user_id="u")
    db.update_task_done(job, INIT_TASK, None)
    digests = ["%064x" % (i + 1) for i in range(2)]
    for i, d in enumerate(digests):
        db.create_task(job, str(i), coproc, wire.task_type_to_value(wire.KeccakReq(d, 17, "ee" * 32)), [], 1, 10)
    store.set_bytes("job:%s:coproc:0:%s" % (job, digests[0]), bytes(range(200)) * 3)            # task-scoped key
    store.set_bytes("job:%s:coproc:%s" % (job, digests[1]), bytes(200))                          # legacy key ([BENTO-KECCAK-013])
    db.create_task(job, "2", prove, wire.task_type_to_value(wire.UnionReq(2, 0, 1)), ["0", "1"], 1, 10)
    agent = tasks.Agent(db, store, p, tasks.AgentArgs(task_stream=wire.COPROC_WORK_TYPE))
    assert tasks.poll_work(agent) == 2 and agent.errors == []
    assert [c[:3] for c in p.calls if c[0] == "prove_keccak"] == [("prove_keccak", digests[0], 17), ("prove_keccak", digests[1], 17)]
    r0 = wire.deserialize_succinct(store.get_bytes("job:%s:keccak_receipts:0" % job))
    assert r0.kind == 8 and "job:%s:coproc:0:%s" % (job, digests[0]) not in store.kv          # input cleaned up after done
    assert tasks.poll_work(tasks.Agent(db, store, p, tasks.AgentArgs(task_stream=wire.PROVE_WORK_TYPE))) == 1
    u = wire.deserialize_succinct(store.get_bytes("job:%s:keccak_receipts:2" % job))
    assert u.kind == KIND_UNION
    # malformed inputs
    for n, (data, tag) in enumerate([(bytes(199), "[BENTO-KECCAK-001] Input length must be a multiple of KeccakState size"),
                                     (b"", "[BENTO-KECCAK-002] Received empty keccak input with claim_digest: " + digests[0])]):
        tid = "k%d" % n
        db.create_task(job, tid, coproc, wire.task_type_to_value(wire.KeccakReq(digests[0], 17, "ee" * 32)), [], 0, 10)
        store.set_bytes("job:%s:coproc:%s:%s" % (job, tid, digests[0]), data)
        a = tasks.Agent(db, store, p, tasks.AgentArgs(task_stream=wire.COPROC_WORK_TYPE))
        tasks.poll_work(a)
        assert a.errors and a.errors[0].endswith("[BENTO-WF-129] Keccak failed: " + tag), a.errors
        db.jobs[job]["state"] = "running"          # let the next malformed case be claimed
