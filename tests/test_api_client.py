"""The GPU worker's REST client (boundless_b200/api_client.py) against a local stub of the reference API server's worker routes
(/root/reference/prover/crates/api/src/lib.rs:902-1041) backed by the in-memory task database and hot store: a whole job is claimed,
proved (fake prover) and reported over HTTP, with the same URL validation and error behaviour as workflow/src/assets.rs."""
import json
import threading
import urllib.parse
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

import pytest

from boundless_b200 import tasks, wire
from boundless_b200.api_client import GPU_WORKER_STREAMS, ApiClient, ApiError, RestHotStore, RestTaskDb
from boundless_b200.taskdb import INIT_TASK, MemoryTaskDb
from test_tasks import IMAGE, FakeProver


class StubApi:
    """axum router of api/src/lib.rs:1181-1189 (worker part), over MemoryTaskDb / MemoryHotStore."""

    def __init__(self, db, store):
        self.db, self.store, self.requests = db, store, []
        outer = self

        class H(BaseHTTPRequestHandler):
            def log_message(self, *a):
                pass

            def _reply(self, code, body=b"", ctype="application/json"):
                self.send_response(code)
                self.send_header("Content-Type", ctype)
                self.send_header("Content-Length", str(len(body)))
                self.end_headers()
                self.wfile.write(body)

            def _handle(self, method):
                u = urllib.parse.urlsplit(self.path)
                q = dict(urllib.parse.parse_qsl(u.query))
                n = int(self.headers.get("Content-Length") or 0)
                body = self.rfile.read(n) if n else b""
                outer.requests.append((method, u.path, q))
                p = u.path.split("/")[1:]
                try:
                    if p[:2] == ["worker", "hot"]:
                        key = "/".join(p[2:])
                        if method == "GET":
                            if key not in outer.store.kv:
                                return self._reply(404, b"hot data missing: " + key.encode(), "text/plain")
                            return self._reply(200, outer.store.get_bytes(key), "application/octet-stream")
                        if method == "PUT":
                            outer.store.set_bytes(key, body)
                            outer.last_ttl = q.get("ttl_secs")
                            return self._reply(204)
                        if method == "DELETE":
                            outer.store.delete(key)
                            return self._reply(204)
                    if p[:2] == ["worker", "assets"] and method == "PUT":
                        outer.store.write_asset("/".join(p[2:]), body)
                        return self._reply(204)
                    if p[:1] == ["assets"] and method == "GET":
                        key = "/".join(p[1:])
                        if key not in outer.store.assets:
                            return self._reply(404, b"missing", "text/plain")
                        return self._reply(200, outer.store.assets[key], "application/octet-stream")
                    if p[:4] == ["worker", "gpu", "tasks", "claim"] and method == "POST":
                        stream = p[4]
                        if stream not in GPU_WORKER_STREAMS:
                            return self._reply(400, ("invalid gpu worker stream: %s" % stream).encode(), "text/plain")
                        t = outer.db.request_work(stream)
                        out = None if t is None else {"job_id": t.job_id, "task_id": t.task_id, "task_def": t.task_def,
                                                      "prereqs": t.prereqs, "max_retries": t.max_retries}
                        return self._reply(200, json.dumps(out).encode())
                    if p[:3] == ["worker", "gpu", "tasks"] and len(p) == 6:
                        job_id, task_id, action = p[3], p[4], p[5]
                        if action == "done" and method == "POST":
                            return self._reply(200, json.dumps({"updated": outer.db.update_task_done(job_id, task_id, json.loads(body)["output"])}).encode())
                        if action == "failed" and method == "POST":
                            return self._reply(200, json.dumps({"updated": outer.db.update_task_failed(job_id, task_id, json.loads(body)["error"])}).encode())
                        if action == "retry" and method == "POST":
                            return self._reply(200, json.dumps({"updated": outer.db.update_task_retry(job_id, task_id)}).encode())
                        if action == "retries-running" and method == "GET":
                            return self._reply(200, json.dumps({"retries": outer.db.get_task_retries_running(job_id, task_id)}).encode())
                except Exception as e:                      # AppError::InternalErr
                    return self._reply(500, str(e).encode(), "text/plain")
                self._reply(404, b"no route", "text/plain")

            def do_GET(self): self._handle("GET")
            def do_POST(self): self._handle("POST")
            def do_PUT(self): self._handle("PUT")
            def do_DELETE(self): self._handle("DELETE")

        self.httpd = ThreadingHTTPServer(("127.0.0.1", 0), H)
        self.url = "http://127.0.0.1:%d" % self.httpd.server_address[1]
        self.thread = threading.Thread(target=self.httpd.serve_forever, daemon=True)
        self.thread.start()

    def close(self):
        self.httpd.shutdown()
        self.httpd.server_close()


@pytest.fixture()
def cluster():
    db = MemoryTaskDb()
    streams = {w: db.create_stream(w, user_id="u") for w in (wire.PROVE_WORK_TYPE, wire.AUX_WORK_TYPE, wire.EXEC_WORK_TYPE)}
    store = tasks.MemoryHotStore()
    stub = StubApi(db, store)
    yield db, streams, store, stub
    stub.close()


def test_url_validation_matches_the_reference():
    """assets.rs:68-121: the base URL must parse, path components must be non-empty and must not start with '/'."""
    with pytest.raises(ApiError, match="must not be empty"):
        ApiClient("")
    with pytest.raises(ApiError, match="Failed to parse API URL"):
        ApiClient("not a url")
    api = ApiClient("http://127.0.0.1:9/")
    assert api.base_url == "http://127.0.0.1:9"
    assert api.worker_task_claim_url("prove") == "http://127.0.0.1:9/worker/gpu/tasks/claim/prove"
    assert api.worker_task_url("j", "t", "done") == "http://127.0.0.1:9/worker/gpu/tasks/j/t/done"
    assert api.worker_hot_url("job:1:segments:0") == "http://127.0.0.1:9/worker/hot/job:1:segments:0"
    assert api.worker_asset_url("receipts/stark/x.bincode") == "http://127.0.0.1:9/worker/assets/receipts/stark/x.bincode"
    assert api.asset_url("a/b") == "http://127.0.0.1:9/assets/a/b"
    for bad, what in ((lambda: api.worker_task_claim_url(""), "task stream"), (lambda: api.worker_task_claim_url("/prove"), "task stream"),
                      (lambda: api.worker_task_url("j", "", "done"), "task id"), (lambda: api.worker_task_url("j", "t", "/x"), "task action"),
                      (lambda: api.worker_hot_url("/k"), "hot-store key"), (lambda: api.worker_asset_url(""), "worker asset key"),
                      (lambda: api.asset_url("/a"), "asset key")):
        with pytest.raises(ApiError, match="Invalid " + what):
            bad()


def test_hot_store_and_assets_roundtrip(cluster):
    db, streams, store, stub = cluster
    api = ApiClient(stub.url)
    blob = bytes(range(256)) * 300
    api.hot_set_bytes("job:1:segments:0", blob)
    assert store.kv["job:1:segments:0"] == blob and api.hot_get_bytes("job:1:segments:0") == blob
    api.hot_set_bytes("k", b"x", ttl_secs=30)
    assert stub.last_ttl == "30"
    api.hot_delete("k")
    api.hot_delete("k")                                       # UNLINK of a missing key is not an error
    with pytest.raises(ApiError, match=r"Hot-store fetch failed for key k at .*HTTP status 404"):
        api.hot_get_bytes("k")
    api.write_asset_buf("receipts/stark/r.bincode", b"receipt")
    assert store.assets["receipts/stark/r.bincode"] == b"receipt" and api.read_asset_buf("receipts/stark/r.bincode") == b"receipt"
    with pytest.raises(ApiError, match="Asset request failed for key nope"):
        api.read_asset_buf("nope")


def test_claim_update_retry_over_http(cluster):
    db, streams, store, stub = cluster
    api = ApiClient(stub.url)
    assert api.claim_gpu_work("prove", 0) is None
    with pytest.raises(ApiError, match="GPU work claim failed for stream exec .*HTTP status 400.*invalid gpu worker stream"):
        api.claim_gpu_work("exec")                             # only prove / join / coproc / snark are GPU streams
    job = db.create_job(streams[wire.EXEC_WORK_TYPE], None, user_id="u")
    db.update_task_done(job, INIT_TASK, None) if db.request_work(wire.EXEC_WORK_TYPE) else None
    db.create_task(job, "0", streams[wire.PROVE_WORK_TYPE], wire.task_type_to_value(wire.ProveReq(0)), [], 2, 30)
    t = api.claim_gpu_work("prove", 5)
    assert (t.job_id, t.task_id, t.task_def, t.prereqs, t.max_retries) == (job, "0", {"Prove": {"index": 0}}, [], 2)
    assert ("POST", "/worker/gpu/tasks/claim/prove", {"wait_timeout_secs": "5"}) in stub.requests
    assert api.get_task_retries_running(job, "0") == 0
    assert api.update_task_retry(job, "0") is True
    assert api.claim_gpu_work("prove").task_id == "0" and api.get_task_retries_running(job, "0") == 1
    assert api.update_task_done(job, "0", {"x": 1}) is True and db.task_field(job, "0", "output") == {"x": 1}
    assert api.update_task_done(job, "0", None) is False       # already done: updated == false, not an error
    assert api.get_task_retries_running(job, "0") is None      # not running any more
    db.create_task(job, "1", streams[wire.PROVE_WORK_TYPE], wire.task_type_to_value(wire.ProveReq(1)), [], 0, 30)
    api.claim_gpu_work("prove")
    assert api.update_task_failed(job, "1", "boom") is True and db.job_state(job) == "failed" and db.job_error(job) == "boom"


@pytest.mark.parametrize("n", [1, 5])
def test_gpu_agent_runs_a_job_through_the_rest_routes(cluster, n):
    """exec and aux workers talk to the task database directly (they are not GPU streams); the GPU agent only speaks HTTP."""
    db, streams, store, stub = cluster
    prover = FakeProver()
    store.set_bytes("input:1", json.dumps({"segments": n, "po2": 10}).encode())
    job = db.create_job(streams[wire.EXEC_WORK_TYPE], wire.task_type_to_value(wire.ExecutorReq(image=IMAGE, input="input:1", user_id="u")),
                        user_id="u")
    exec_agent = tasks.Agent(db, store, None, tasks.AgentArgs(task_stream=wire.EXEC_WORK_TYPE, segment_po2=10))
    aux_agent = tasks.Agent(db, store, prover, tasks.AgentArgs(task_stream=wire.AUX_WORK_TYPE))
    api = ApiClient(stub.url)
    gpu_agent = tasks.Agent(RestTaskDb(api), RestHotStore(api), prover, tasks.AgentArgs(task_stream=wire.PROVE_WORK_TYPE))
    assert tasks.poll_work(exec_agent) == 1
    assert tasks.poll_work(gpu_agent) == n + (n - 1) + 1       # proves, joins, resolve: all claimed and reported over HTTP
    assert tasks.poll_work(aux_agent) == 1 and db.job_state(job) == "done"
    assert not [k for k in store.kv if k.startswith("job:")]   # hot keys were deleted through DELETE /worker/hot/...
    root, journal = wire.deserialize_rollup(store.assets["receipts/stark/%s.bincode" % job])
    assert root.claim == (0, n - 1)
    claims = [r for r in stub.requests if r[1] == "/worker/gpu/tasks/claim/prove"]
    assert len(claims) == 2 * n + 1                             # one per task plus the empty claim that ends the loop
    # the same job without HTTP gives the same root receipt
    db2 = MemoryTaskDb()
    s2 = {w: db2.create_stream(w, user_id="u") for w in (wire.PROVE_WORK_TYPE, wire.AUX_WORK_TYPE, wire.EXEC_WORK_TYPE)}
    st2 = tasks.MemoryHotStore(); st2.set_bytes("input:1", json.dumps({"segments": n, "po2": 10}).encode())
    p2 = FakeProver()
    job2 = db2.create_job(s2[wire.EXEC_WORK_TYPE], wire.task_type_to_value(wire.ExecutorReq(image=IMAGE, input="input:1", user_id="u")), user_id="u")
    for a in (tasks.Agent(db2, st2, None, tasks.AgentArgs(task_stream=wire.EXEC_WORK_TYPE, segment_po2=10)),
              tasks.Agent(db2, st2, p2, tasks.AgentArgs(task_stream=wire.PROVE_WORK_TYPE)),
              tasks.Agent(db2, st2, p2, tasks.AgentArgs(task_stream=wire.AUX_WORK_TYPE))):
        tasks.poll_work(a)
    root2, _ = wire.deserialize_rollup(st2.assets["receipts/stark/%s.bincode" % job2])
    assert (root2.seal == root.seal).all()


def test_failures_travel_over_http(cluster):
    db, streams, store, stub = cluster
    prover = FakeProver(); prover.fail_next["prove_segment"] = 99
    store.set_bytes("input:1", json.dumps({"segments": 2, "po2": 10}).encode())
    job = db.create_job(streams[wire.EXEC_WORK_TYPE], wire.task_type_to_value(wire.ExecutorReq(image=IMAGE, input="input:1", user_id="u")),
                        user_id="u")
    tasks.poll_work(tasks.Agent(db, store, None, tasks.AgentArgs(task_stream=wire.EXEC_WORK_TYPE, segment_po2=10)))
    api = ApiClient(stub.url)
    tasks.poll_work(tasks.Agent(RestTaskDb(api), RestHotStore(api), prover, tasks.AgentArgs(task_stream=wire.PROVE_WORK_TYPE)))
    assert db.job_state(job) == "failed"
    assert db.job_error(job) == "retry max hit: [BENTO-WF-115] Prove failed: injected prove_segment failure"
    assert [r for r in stub.requests if r[1].endswith("/retry")] and [r for r in stub.requests if r[1].endswith("/failed")]
    # a dead server is an error with the reference's context, not a hang
    dead = ApiClient("http://127.0.0.1:9", timeout=2)
    with pytest.raises(ApiError, match="GPU work claim failed for stream prove"):
        dead.claim_gpu_work("prove")
