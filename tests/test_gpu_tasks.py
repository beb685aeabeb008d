"""The bento job flow with real device-side proofs and device-side verify_integrity: executor stand-in -> taskdb -> GPU agent
(tasks/prove.rs, join.rs, resolve.rs) -> aux agent (finalize.rs).  The root receipt must equal, word for word, the one the CPU
oracle builds for the same job, and must pass the oracle's verifier."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

IMAGE = "cd" * 32


def test_job_through_the_agent_loop(gpu, oracle):
    from boundless_b200 import ProverOpts, get_prover_server, tasks, wire
    from boundless_b200.prover_server import KIND_JOIN, KIND_LIFT, RECURSION_WIDTHS
    from boundless_b200.taskdb import MemoryTaskDb
    n, po2, rp = 3, 9, 11
    srv = get_prover_server(ProverOpts(segment_po2=12, recursion_po2=rp, slots=1))
    try:
        db = MemoryTaskDb()
        db.create_stream(wire.PROVE_WORK_TYPE, user_id="u"); db.create_stream(wire.AUX_WORK_TYPE, user_id="u")
        execs = db.create_stream(wire.EXEC_WORK_TYPE, user_id="u")
        store = tasks.MemoryHotStore()
        store.set_bytes("input:1", json.dumps({"segments": n, "po2": po2}).encode())
        job = db.create_job(execs, wire.task_type_to_value(wire.ExecutorReq(image=IMAGE, input="input:1", user_id="u")), user_id="u")
        launches0 = gpu.cuda.is_available() and srv.L.b200_kernel_launches()
        assert tasks.poll_work(tasks.Agent(db, store, None, tasks.AgentArgs(task_stream=wire.EXEC_WORK_TYPE))) == 1
        gpu_agent = tasks.Agent(db, store, srv, tasks.AgentArgs(task_stream=wire.PROVE_WORK_TYPE))
        claimed = tasks.poll_work(gpu_agent)
        assert gpu_agent.errors == [] and claimed == n + (n - 1) + 1
        assert tasks.poll_work(tasks.Agent(db, store, srv, tasks.AgentArgs(task_stream=wire.AUX_WORK_TYPE))) == 1
        assert db.job_state(job) == "done", db.job_error(job)
        assert srv.L.b200_kernel_launches() > launches0
        root, journal = wire.deserialize_rollup(store.assets["receipts/stark/%s.bincode" % job])
        assert root.claim == (0, n - 1) and json.loads(journal) == {"segments": n}

        def rec(kind, digest):
            return oracle.prove(rp, int(digest[0]) | (int(digest[1]) << 32), *RECURSION_WIDTHS, kind=kind, input_digest=digest)
        lifts = [rec(KIND_LIFT, oracle.seal_digest(oracle.prove(po2, 0xB2000000 + i))) for i in range(n)]
        j01 = rec(KIND_JOIN, oracle.hash_pair(oracle.seal_digest(lifts[0]), oracle.seal_digest(lifts[1])))
        j012 = rec(KIND_JOIN, oracle.hash_pair(oracle.seal_digest(j01), oracle.seal_digest(lifts[2])))
        assert np.array_equal(root.seal, j012)
        assert oracle.verify(root.seal) == 0

        # a corrupted stored receipt is caught by the device-side verify_integrity inside the join task
        job2 = db.create_job(execs, wire.task_type_to_value(wire.ExecutorReq(image=IMAGE, input="input:1", user_id="u")), user_id="u")
        tasks.poll_work(tasks.Agent(db, store, None, tasks.AgentArgs(task_stream=wire.EXEC_WORK_TYPE)))
        tasks.poll_work(gpu_agent, max_tasks=2)
        key = "job:%s:recursion_receipts:0" % job2
        bad = wire.deserialize_succinct(store.get_bytes(key)); bad.seal[bad.seal.size // 3] ^= 1
        store.set_bytes(key, wire.serialize_succinct(bad))
        tasks.poll_work(gpu_agent)
        assert db.job_state(job2) == "failed"
        assert "[BENTO-JOIN-003] Failed to verify left receipt integrity: seal does not verify" in db.job_error(job2)
    finally:
        srv.close()


def test_povw_job_and_pipelined_agent_on_the_gpu(gpu, oracle):
    """(a) The PoVW flow with real proofs: lift_povw / join_povw / unwrap_povw through the agent loop; the uploaded root equals the
    oracle's unwrap of the oracle's PoVW join.  (b) The pipelined agent loop (two Prove claims in flight, composite prove+lift tasks)
    in JOIN_STREAM mode leaves exactly the lifted receipts the synchronous loop would."""
    from boundless_b200 import ProverOpts, get_prover_server, tasks, wire
    from boundless_b200.prover_server import KIND_JOIN_POVW, KIND_LIFT, KIND_LIFT_POVW, KIND_UNWRAP_POVW, RECURSION_WIDTHS
    from boundless_b200.taskdb import MemoryTaskDb
    n, po2, rp = 2, 9, 11
    srv = get_prover_server(ProverOpts(segment_po2=12, recursion_po2=rp, slots=2))

    def rec(kind, digest):
        return oracle.prove(rp, int(digest[0]) | (int(digest[1]) << 32), *RECURSION_WIDTHS, kind=kind, input_digest=digest)
    try:
        db = MemoryTaskDb()
        db.create_stream(wire.PROVE_WORK_TYPE, user_id="u"); db.create_stream(wire.AUX_WORK_TYPE, user_id="u")
        db.create_stream(wire.JOIN_WORK_TYPE, user_id="u")
        execs = db.create_stream(wire.EXEC_WORK_TYPE, user_id="u")
        store = tasks.MemoryHotStore()
        store.set_bytes("input:1", json.dumps({"segments": n, "po2": po2}).encode())
        req = wire.task_type_to_value(wire.ExecutorReq(image=IMAGE, input="input:1", user_id="u"))
        # (a)
        job = db.create_job(execs, req, user_id="u")
        tasks.poll_work(tasks.Agent(db, store, None, tasks.AgentArgs(task_stream=wire.EXEC_WORK_TYPE)))
        agent = tasks.Agent(db, store, srv, tasks.AgentArgs(task_stream=wire.PROVE_WORK_TYPE, povw_job_number=7), povw="0x" + "11" * 20)
        assert tasks.poll_work(agent) == n + (n - 1) + 1 and agent.errors == []
        tasks.poll_work(tasks.Agent(db, store, srv, tasks.AgentArgs(task_stream=wire.AUX_WORK_TYPE)))
        assert db.job_state(job) == "done", db.job_error(job)
        lifts = [rec(KIND_LIFT_POVW, oracle.seal_digest(oracle.prove(po2, 0xB2000000 + i))) for i in range(n)]
        j = rec(KIND_JOIN_POVW, oracle.hash_pair(oracle.seal_digest(lifts[0]), oracle.seal_digest(lifts[1])))
        u = rec(KIND_UNWRAP_POVW, oracle.seal_digest(j))
        root, _ = wire.deserialize_rollup(store.assets["receipts/stark/%s.bincode" % job])
        assert np.array_equal(root.seal, u)
        assert np.array_equal(wire.deserialize_succinct(store.assets["work_receipts/%s.bincode" % job]).seal, j)
        assert json.loads(store.assets["work_receipts/%s_metadata.json" % job])["povw_job_number"] == "7"
        # (b)
        job2 = db.create_job(execs, req, user_id="u")
        tasks.poll_work(tasks.Agent(db, store, None, tasks.AgentArgs(task_stream=wire.EXEC_WORK_TYPE, join_stream=True)))
        pa = tasks.Agent(db, store, srv, tasks.AgentArgs(task_stream=wire.PROVE_WORK_TYPE))
        assert tasks.poll_work_pipelined(pa) == n and pa.errors == [] and len(pa.processed) == n
        for i in range(n):
            got = wire.deserialize_succinct(store.get_bytes("job:%s:recursion_receipts:%d" % (job2, i)))
            assert got.kind == KIND_LIFT and got.claim == (i, i)
            assert np.array_equal(got.seal, rec(KIND_LIFT, oracle.seal_digest(oracle.prove(po2, 0xB2000000 + i))))
        assert tasks.poll_work(tasks.Agent(db, store, srv, tasks.AgentArgs(task_stream=wire.JOIN_WORK_TYPE))) == (n - 1) + 1
        tasks.poll_work(tasks.Agent(db, store, srv, tasks.AgentArgs(task_stream=wire.AUX_WORK_TYPE)))
        assert db.job_state(job2) == "done", db.job_error(job2)
    finally:
        srv.close()
