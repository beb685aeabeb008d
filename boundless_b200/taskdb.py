"""In-process task database with the reference's scheduling semantics: this is how proving work is sharded over GPUs.

Reference: the Redis-Lua backend of `taskdb` (/root/reference/prover/crates/taskdb/src/redis_backend.rs): `create_task`
(:219-286), `request_work` (:288-338), `update_task_done` (:340-394), `update_task_failed` (:396-441), `update_task_retry`
(:463-503), ready queues keyed by (worker_type, priority) and scored by a global creation counter (`enqueue_task`, :122-134),
the running set keyed by deadline = claim time + timeout (:327) and `requeue_tasks` for expired claims.

What matters for the proving path: a claim pops the lowest priority number first, then the OLDEST task by
creation order; a task becomes ready when its `waiting_on` counter reaches 0; joins are created as soon as two peaks merge, so
they carry early sequence numbers and run before later segments once unblocked (depth-first reduction, bounded live receipts).
Every GPU agent is a puller on the same queues, which is the whole multi-GPU story of the reference (compose.yml:113).

This is the control-plane stand-in used by `boundless_b200.tasks` and the tests; a deployment keeps the reference's Redis taskdb
and only swaps the agent.  Thread-safe (one lock), no persistence.
"""
import json
import threading
import uuid
from dataclasses import dataclass
from typing import Dict, List, Optional

INIT_TASK = "init"          # taskdb/src/lib.rs:85


class TaskDbError(RuntimeError):
    pass


@dataclass
class ReadyTask:
    """taskdb::ReadyTask (taskdb/src/lib.rs:76-82)."""
    job_id: str
    task_id: str
    task_def: object
    prereqs: list
    max_retries: int


class MemoryTaskDb:
    def __init__(self, clock=None):
        self._lock = threading.Lock()
        self._clock = clock or __import__("time").time
        self._seq = 0
        self.streams: Dict[str, dict] = {}
        self.jobs: Dict[str, dict] = {}
        self.tasks: Dict[tuple, dict] = {}
        self.deps: Dict[tuple, List[str]] = {}
        self.ready: Dict[tuple, Dict[str, int]] = {}          # (worker_type, priority) -> {compound: sort_seq}
        self.running: Dict[str, float] = {}                   # compound -> deadline

    # -- streams / jobs ---------------------------------------------------------------------------------------------------
    def create_stream(self, worker_type: str, reserved: int = 0, be_mult: float = 1.0, user_id: str = "") -> str:
        with self._lock:
            sid = str(uuid.uuid4())
            self.streams[sid] = {"worker_type": worker_type, "reserved": reserved, "be_mult": be_mult, "user_id": user_id}
            return sid

    def get_stream(self, user_id: str, worker_type: str) -> Optional[str]:
        with self._lock:
            for sid, st in self.streams.items():
                if st["user_id"] == user_id and st["worker_type"] == worker_type:
                    return sid
            return None

    def create_job(self, stream_id: str, task_def, max_retries: int = 0, timeout_secs: int = 0, user_id: str = "",
                   priority: int = 1) -> str:
        """create_job: the job plus its `init` task (state ready, no prerequisites)."""
        with self._lock:
            if stream_id not in self.streams:
                raise TaskDbError("missing stream: %s" % stream_id)
            job_id = str(uuid.uuid4())
            self.jobs[job_id] = {"state": "running", "error": "", "user_id": user_id, "priority": int(priority), "tasks": []}
            self._insert(job_id, INIT_TASK, stream_id, task_def, [], max_retries, timeout_secs, "ready", 0)
            return job_id

    # -- tasks --------------------------------------------------------------------------------------------------------------
    def _insert(self, job_id, task_id, stream_id, task_def, prereqs, max_retries, timeout_secs, state, waiting_on):
        self._seq += 1
        wt = self.streams[stream_id]["worker_type"]
        t = {"stream_id": stream_id, "worker_type": wt, "priority": self.jobs[job_id]["priority"], "sort_seq": self._seq,
             "task_def": task_def, "prerequisites": list(prereqs), "state": state, "waiting_on": waiting_on, "retries": 0,
             "max_retries": int(max_retries), "timeout_secs": int(timeout_secs), "output": None, "error": "",
             "created_at": self._clock(), "started_at": None}
        self.tasks[(job_id, task_id)] = t
        self.jobs[job_id]["tasks"].append(task_id)
        if state == "ready":
            self._enqueue(job_id, task_id)

    def _enqueue(self, job_id, task_id):
        t = self.tasks[(job_id, task_id)]
        q = self.ready.setdefault((t["worker_type"], t["priority"]), {})
        q.setdefault(job_id + "|" + task_id, t["sort_seq"])               # ZADD NX

    def _dequeue(self, job_id, task_id):
        t = self.tasks[(job_id, task_id)]
        self.ready.get((t["worker_type"], t["priority"]), {}).pop(job_id + "|" + task_id, None)

    def create_task(self, job_id: str, task_id: str, stream_id: str, task_def, prereqs: List[str], max_retries: int,
                    timeout_secs: int) -> None:
        with self._lock:
            if job_id not in self.jobs:
                raise TaskDbError("missing job: %s" % job_id)
            if stream_id not in self.streams:
                raise TaskDbError("missing stream: %s" % stream_id)
            if (job_id, task_id) in self.tasks:
                raise TaskDbError("task already exists: %s" % task_id)
            waiting = 0
            for pre in prereqs:
                if (job_id, pre) not in self.tasks:
                    raise TaskDbError("missing prerequisite task: %s" % pre)
            for pre in prereqs:
                self.deps.setdefault((job_id, pre), []).append(task_id)
                if self.tasks[(job_id, pre)]["state"] != "done":
                    waiting += 1
            self._insert(job_id, task_id, stream_id, task_def, prereqs, max_retries, timeout_secs,
                         "ready" if waiting == 0 else "pending", waiting)

    def request_work(self, worker_type: str) -> Optional[ReadyTask]:
        """Lowest priority number first, then creation order; stale entries are skipped; a failed job cancels its tasks."""
        with self._lock:
            now = self._clock()
            for priority in (0, 1, 2):
                q = self.ready.get((worker_type, priority))
                while q:
                    compound = min(q, key=q.get)                              # ZPOPMIN
                    del q[compound]
                    job_id, task_id = compound.split("|", 1)
                    t = self.tasks.get((job_id, task_id))
                    if t is None or t["state"] != "ready":
                        continue
                    if self.jobs[job_id]["state"] == "failed":
                        t["state"] = "cancelled"
                        continue
                    t["state"] = "running"; t["started_at"] = now
                    self.running[compound] = now + max(t["timeout_secs"], 0)
                    return ReadyTask(job_id, task_id, t["task_def"], list(t["prerequisites"]), t["max_retries"])
            return None

    def update_task_done(self, job_id: str, task_id: str, output=None) -> bool:
        with self._lock:
            t = self.tasks.get((job_id, task_id))
            if t is None or t["state"] not in ("ready", "running"):
                return False
            was = t["state"]
            t["state"] = "done"; t["output"] = output
            self.running.pop(job_id + "|" + task_id, None)
            if was == "ready":
                self._dequeue(job_id, task_id)
            for dep in self.deps.get((job_id, task_id), []):
                d = self.tasks[(job_id, dep)]
                if d["state"] not in ("failed", "done"):
                    d["waiting_on"] -= 1
                    if d["waiting_on"] <= 0:
                        d["waiting_on"] = 0; d["state"] = "ready"
                        self._enqueue(job_id, dep)
            if all(self.tasks[(job_id, tid)]["state"] == "done" for tid in self.jobs[job_id]["tasks"]):
                self.jobs[job_id]["state"] = "done"
            return True

    def update_task_failed(self, job_id: str, task_id: str, err: str) -> bool:
        with self._lock:
            t = self.tasks.get((job_id, task_id))
            if t is None or t["state"] not in ("ready", "running", "pending"):
                return False
            was = t["state"]
            t["state"] = "failed"; t["error"] = err
            if was == "running":
                self.running.pop(job_id + "|" + task_id, None)
            elif was == "ready":
                self._dequeue(job_id, task_id)
            self.jobs[job_id]["state"] = "failed"; self.jobs[job_id]["error"] = err
            for tid in self.jobs[job_id]["tasks"]:                              # cancel the siblings that have not started
                s = self.tasks[(job_id, tid)]
                if tid != task_id and s["state"] in ("ready", "pending"):
                    if s["state"] == "ready":
                        self._dequeue(job_id, tid)
                    s["state"] = "cancelled"
            return True

    def update_task_retry(self, job_id: str, task_id: str) -> bool:
        """running -> ready with retries+1, or -> failed "retry max hit" once retries exceed max_retries (returns False)."""
        with self._lock:
            t = self.tasks.get((job_id, task_id))
            if t is None or t["state"] != "running":
                return False
            t["retries"] += 1
            self.running.pop(job_id + "|" + task_id, None)
            if t["retries"] > t["max_retries"]:
                t["state"] = "failed"; t["error"] = "retry max hit"
                self.jobs[job_id]["state"] = "failed"; self.jobs[job_id]["error"] = "retry max hit"
                return False
            t["state"] = "ready"; t["error"] = ""
            self._enqueue(job_id, task_id)
            return True

    def get_task_retries_running(self, job_id: str, task_id: str) -> Optional[int]:
        with self._lock:
            t = self.tasks.get((job_id, task_id))
            return t["retries"] if t is not None and t["state"] == "running" else None

    def requeue_tasks(self, limit: int = 100) -> int:
        """Claims whose deadline has passed go back to ready (or fail at max retries), oldest deadline first."""
        with self._lock:
            now = self._clock()
            expired = sorted((d, c) for c, d in self.running.items() if d <= now)[:limit]
        n = 0
        for _, compound in expired:
            job_id, task_id = compound.split("|", 1)
            if self.update_task_retry(job_id, task_id):
                n += 1
        return n

    # -- inspection ---------------------------------------------------------------------------------------------------------
    def job_state(self, job_id: str) -> str:
        with self._lock:
            return self.jobs[job_id]["state"]

    def job_error(self, job_id: str) -> str:
        with self._lock:
            return self.jobs[job_id]["error"]

    def task_state(self, job_id: str, task_id: str) -> str:
        with self._lock:
            return self.tasks[(job_id, task_id)]["state"]

    def task_field(self, job_id: str, task_id: str, name: str):
        with self._lock:
            return self.tasks[(job_id, task_id)][name]

    def dump(self) -> str:
        with self._lock:
            return json.dumps({j + "|" + t: v["state"] for (j, t), v in self.tasks.items()}, indent=1)
