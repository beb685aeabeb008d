"""Python face of the C++ join-tree planner (csrc/planner.cpp), mirroring taskdb::planner::Planner
(/root/reference/prover/crates/taskdb/src/planner/mod.rs:91-240): same method names, same errors."""
import ctypes as C

from . import lib as _lib
from .lib import CMD_FINALIZE, CMD_JOIN, CMD_KECCAK, CMD_SEGMENT, CMD_UNION, Task  # noqa: F401


class PlannerErr(Exception):
    pass


class PlanTask:
    def __init__(self, t: Task):
        self.task_number = t.task_number
        self.task_height = t.task_height
        self.command = t.command
        self.depends_on = [t.depends_on[i] for i in range(t.n_depends_on)]
        self.keccak_depends_on = [t.keccak_depends_on[i] for i in range(t.n_keccak_depends_on)]

    def __repr__(self):
        names = {CMD_KECCAK: "Keccak", CMD_FINALIZE: "Finalize", CMD_JOIN: "Join", CMD_SEGMENT: "Segment", CMD_UNION: "Union"}
        return "%d %s h=%d deps=%s kdeps=%s" % (self.task_number, names[self.command], self.task_height, self.depends_on,
                                                self.keccak_depends_on)


class Planner:
    def __init__(self):
        self.L = _lib.load()
        self.h = C.c_void_p(self.L.b200_planner_new())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.b200_planner_free(self.h)
            self.h = None

    def enqueue_segment(self):
        n = self.L.b200_planner_enqueue_segment(self.h)
        if n < 0:
            raise PlannerErr("PlanFinalized")
        return n

    def enqueue_keccak(self):
        n = self.L.b200_planner_enqueue_keccak(self.h)
        if n < 0:
            raise PlannerErr("PlanFinalized")
        return n

    def finish(self):
        n = self.L.b200_planner_finish(self.h)
        if n < 0:
            raise PlannerErr("PlanNotStartedString")
        return n

    def task_count(self):
        return self.L.b200_planner_task_count(self.h)

    def get_task(self, task_number):
        t = Task()
        if self.L.b200_planner_get_task(self.h, task_number, C.byref(t)) != 0:
            raise IndexError("Invalid task number %d" % task_number)
        return PlanTask(t)

    def next_task(self):
        t = Task()
        if self.L.b200_planner_next_task(self.h, C.byref(t)) != 0:
            return None
        return PlanTask(t)
