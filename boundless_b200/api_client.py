"""The GPU worker's REST client: claim work, report done / failed / retry, and move blobs through the hot store, over the same HTTP
routes the reference agent uses, so that a B200 box running `tasks.poll_work` can sit behind an existing bento API server.

Reference: `ApiClient` in /root/reference/prover/crates/workflow/src/assets.rs:14-365 (client side) and the handlers in
/root/reference/prover/crates/api/src/lib.rs:902-1041 (server side):

    POST   {base}/worker/gpu/tasks/claim/{task_stream}?wait_timeout_secs=N   -> JSON `WorkerTask | null`
    POST   {base}/worker/gpu/tasks/{job_id}/{task_id}/done    {"output": ...} -> {"updated": bool}
    POST   {base}/worker/gpu/tasks/{job_id}/{task_id}/failed  {"error": str}  -> {"updated": bool}
    POST   {base}/worker/gpu/tasks/{job_id}/{task_id}/retry                   -> {"updated": bool}
    GET    {base}/worker/gpu/tasks/{job_id}/{task_id}/retries-running         -> {"retries": int | null}
    GET | PUT (?ttl_secs=N) | DELETE  {base}/worker/hot/{key}                 -> bytes | 204 | 204
    PUT    {base}/worker/assets/{key}                                         -> 204
    GET    {base}/assets/{key}                                                -> bytes

URL validation (empty or '/'-prefixed path components are rejected before any request is made) and the error contexts follow
assets.rs:80-121 and :193-365.  `RestTaskDb` / `RestHotStore` adapt the client to the two interfaces `tasks.Agent` is written
against (the same ones `taskdb.MemoryTaskDb` / `tasks.MemoryHotStore` implement), so `poll_work` runs unchanged over HTTP.
Only the Python standard library is used (urllib); the server side -- Redis, Postgres, object storage -- is the reference's
control plane and out of scope.
"""
import json
import urllib.error
import urllib.parse
import urllib.request
import uuid
from typing import Optional

from .taskdb import ReadyTask

GPU_WORKER_STREAMS = ("prove", "join", "coproc", "snark")      # api/src/lib.rs:129-131 is_gpu_worker_stream


class ApiError(RuntimeError):
    """anyhow-style error: "context: cause"."""


def _component(value: str, what: str) -> str:
    if not value or value.startswith("/"):
        raise ApiError("Invalid %s: %s" % (what, value))
    return value


class ApiClient:
    def __init__(self, base_url: str, timeout: float = 60.0):
        base_url = base_url.rstrip("/")
        if not base_url:
            raise ApiError("API URL must not be empty")
        parts = urllib.parse.urlsplit(base_url)
        if not parts.scheme or not parts.netloc:
            raise ApiError("Failed to parse API URL: %s" % base_url)
        self.base_url, self.timeout = base_url, timeout

    # ---- URL builders (assets.rs:80-121) --------------------------------------------------------------------------
    def asset_url(self, key):
        return "%s/assets/%s" % (self.base_url, _component(key, "asset key"))

    def worker_task_claim_url(self, task_stream):
        return "%s/worker/gpu/tasks/claim/%s" % (self.base_url, _component(task_stream, "task stream"))

    def worker_task_url(self, job_id, task_id, action):
        _component(task_id, "task id")
        _component(action, "task action")
        return "%s/worker/gpu/tasks/%s/%s/%s" % (self.base_url, job_id, task_id, action)

    def worker_hot_url(self, key):
        return "%s/worker/hot/%s" % (self.base_url, _component(key, "hot-store key"))

    def worker_asset_url(self, key):
        return "%s/worker/assets/%s" % (self.base_url, _component(key, "worker asset key"))

    # ---- transport -------------------------------------------------------------------------------------------------
    def _send(self, method, url, body=None, content_type=None, context=""):
        req = urllib.request.Request(url, data=body, method=method)
        if content_type:
            req.add_header("Content-Type", content_type)
        try:
            with urllib.request.urlopen(req, timeout=self.timeout) as resp:
                return resp.read()
        except urllib.error.HTTPError as e:            # error_for_status()
            detail = e.read().decode("utf-8", "replace")[:512]
            raise ApiError("%s: HTTP status %d for %s%s" % (context, e.code, url, (": " + detail) if detail else ""))
        except (urllib.error.URLError, OSError) as e:
            raise ApiError("%s: %s" % (context, e))

    def _json(self, method, url, payload, context, decode_context):
        body = None if payload is None else json.dumps(payload).encode()
        raw = self._send(method, url, body, "application/json" if body is not None else None, context)
        try:
            return json.loads(raw.decode() or "null")
        except ValueError as e:
            raise ApiError("%s: %s" % (decode_context, e))

    # ---- task routes (assets.rs:193-308) -----------------------------------------------------------------------------
    def claim_gpu_work(self, task_stream: str, wait_timeout_secs: int = 0) -> Optional[ReadyTask]:
        url = self.worker_task_claim_url(task_stream) + "?" + urllib.parse.urlencode({"wait_timeout_secs": int(wait_timeout_secs)})
        task = self._json("POST", url, None, "GPU work claim failed for stream %s at %s" % (task_stream, url),
                          "Failed to decode GPU work claim response from %s" % url)
        if task is None:
            return None
        try:
            job_id = str(uuid.UUID(task["job_id"]))
        except (ValueError, KeyError, TypeError) as e:
            raise ApiError("Invalid worker job_id %s: %s" % (task.get("job_id") if isinstance(task, dict) else task, e))
        return ReadyTask(job_id, task["task_id"], task["task_def"], task["prereqs"], int(task["max_retries"]))

    def _update(self, job_id, task_id, action, payload, what):
        url = self.worker_task_url(job_id, task_id, action)
        res = self._json("POST", url, payload, "Task %s update failed for %s:%s" % (what, job_id, task_id),
                         "Failed to decode task %s response for %s:%s" % (what, job_id, task_id))
        return bool(res["updated"])

    def update_task_done(self, job_id, task_id, output=None) -> bool:
        return self._update(job_id, task_id, "done", {"output": output}, "done")

    def update_task_failed(self, job_id, task_id, error: str) -> bool:
        return self._update(job_id, task_id, "failed", {"error": error}, "failed")

    def update_task_retry(self, job_id, task_id) -> bool:
        return self._update(job_id, task_id, "retry", None, "retry")

    def get_task_retries_running(self, job_id, task_id) -> Optional[int]:
        url = self.worker_task_url(job_id, task_id, "retries-running")
        res = self._json("GET", url, None, "Task retries fetch failed for %s:%s" % (job_id, task_id),
                         "Failed to decode retries-running response for %s:%s" % (job_id, task_id))
        return res["retries"]

    # ---- hot store and assets (assets.rs:123-191, :310-365) ------------------------------------------------------------
    def hot_get_bytes(self, key: str) -> bytes:
        url = self.worker_hot_url(key)
        return self._send("GET", url, context="Hot-store fetch failed for key %s at %s" % (key, url))

    def hot_set_bytes(self, key: str, value: bytes, ttl_secs: Optional[int] = None) -> None:
        url = self.worker_hot_url(key)
        if ttl_secs is not None:
            url += "?" + urllib.parse.urlencode({"ttl_secs": int(ttl_secs)})
        self._send("PUT", url, bytes(value), "application/octet-stream", "Hot-store write failed for key %s" % key)

    def hot_delete(self, key: str) -> None:
        self._send("DELETE", self.worker_hot_url(key), context="Hot-store delete failed for key %s" % key)

    def read_asset_buf(self, key: str) -> bytes:
        url = self.asset_url(key)
        return self._send("GET", url, context="Asset request failed for key %s at %s" % (key, url))

    def write_asset_buf(self, key: str, value: bytes) -> None:
        self._send("PUT", self.worker_asset_url(key), bytes(value), "application/octet-stream",
                   "Worker asset upload failed for key %s" % key)


class RestTaskDb:
    """The task-database face `tasks.poll_work` / `tasks.process_work` use, over the worker routes (lib.rs:611-677 calls these through
    `Agent::claim_work` / `update_task_*`, which pick the REST client whenever `--api-url` is set)."""

    def __init__(self, api: ApiClient, wait_timeout_secs: int = 0):
        self.api, self.wait_timeout_secs = api, wait_timeout_secs

    def request_work(self, worker_type: str) -> Optional[ReadyTask]:
        return self.api.claim_gpu_work(worker_type, self.wait_timeout_secs)

    def update_task_done(self, job_id, task_id, output=None) -> bool:
        return self.api.update_task_done(job_id, task_id, output)

    def update_task_failed(self, job_id, task_id, err: str) -> bool:
        return self.api.update_task_failed(job_id, task_id, err)

    def update_task_retry(self, job_id, task_id) -> bool:
        return self.api.update_task_retry(job_id, task_id)

    def get_task_retries_running(self, job_id, task_id) -> Optional[int]:
        return self.api.get_task_retries_running(job_id, task_id)


class RestHotStore:
    """The hot-store face of `tasks.Agent` over /worker/hot and /worker/assets."""

    def __init__(self, api: ApiClient, ttl_secs: Optional[int] = None):
        self.api, self.ttl_secs = api, ttl_secs

    def get_bytes(self, key: str) -> bytes:
        return self.api.hot_get_bytes(key)

    def set_bytes(self, key: str, value: bytes) -> None:
        self.api.hot_set_bytes(key, value, self.ttl_secs)

    def delete(self, key: str) -> None:
        self.api.hot_delete(key)

    def write_asset(self, key: str, value: bytes) -> None:
        self.api.write_asset_buf(key, value)
