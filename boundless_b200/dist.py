"""Multi-GPU sharding of the segment-proving path: one process per GPU, segments partitioned across ranks, receipts
moved between ranks only where the join tree needs them.

Reference: one agent process per GPU pulling `Prove` tasks from a shared queue (compose.yml:113;
prover/crates/taskdb/src/redis_backend.rs:288) and `Join` tasks whose inputs travel as bincode blobs through Redis
(tasks/prove.rs:113-117 -> tasks/join.rs:27-34).  Here the same task DAG (taskdb Planner semantics, csrc/planner.cpp)
is executed by `world` ranks: segment i belongs to rank owner(i) (contiguous blocks, so the Planner's left-to-right
adjacency keeps most joins rank-local), a join runs on the rank that owns its LEFT input, and the right input is
sent rank-to-rank with torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests).  There is no collective
on the proving path itself: segments are independent (weak scaling).
"""
from typing import Callable, List, Optional

import numpy as np

from .planner import CMD_FINALIZE, CMD_JOIN, CMD_SEGMENT, Planner


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous block partition [lo, hi) of n_items over `world` ranks (first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of(segment_idx: int, n_items: int, world: int) -> int:
    base, extra = divmod(n_items, world)
    cut = extra * (base + 1)
    if segment_idx < cut:
        return segment_idx // (base + 1)
    return extra + (segment_idx - cut) // max(base, 1)


def plan_job(n_segments: int):
    """The reference's online plan for n segments: list of PlanTask in creation order, last one is Finalize."""
    pl = Planner()
    for _ in range(n_segments):
        pl.enqueue_segment()
    pl.finish()
    return [pl.get_task(i) for i in range(pl.task_count())]


class _Comm:
    """Point-to-point movement of fixed-size u32 buffers; torch.distributed when world > 1."""

    def __init__(self, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.rank = dist.get_rank() if self.on else 0
        self.world = dist.get_world_size() if self.on else 1
        self.device = device
        self.bytes_moved = 0

    def _tensor(self, words):
        import torch
        return torch.empty(words, dtype=torch.int32, device=self.device or "cpu")

    def send(self, arr: np.ndarray, dst: int):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int32))
        if self.device is not None:
            t = t.to(self.device)
        self.dist.send(t, dst)
        self.bytes_moved += arr.size * 4

    def recv(self, words: int, src: int) -> np.ndarray:
        t = self._tensor(words)
        self.dist.recv(t, src)
        return t.cpu().numpy().view(np.uint32).copy()


def prove_job(n_segments: int, prove_and_lift: Callable[[int], object], join: Callable[[object, object], object],
              seal_of: Callable[[object], np.ndarray], receipt_from_seal: Callable[[np.ndarray, tuple], object],
              recursion_seal_words: int, device=None, root_rank: int = 0,
              prove_and_lift_many: Optional[Callable[[List[int]], List[object]]] = None):
    """Run the whole job DAG.  Every rank calls this with the same arguments.

    prove_and_lift(i)  -> lifted receipt of segment i (ProverServer.prove_segment + lift, tasks/prove.rs:44-104)
    join(a, b)         -> joined receipt                  (tasks/join.rs:52-56)
    prove_and_lift_many(indices) -> receipts: optional batched form (lets the caller keep several proofs in flight); the
                          segment tasks have no prerequisites (tasks/executor.rs:86-120), so a rank may run all of its own first.
    Returns (root receipt on `root_rank` else None, stats dict)."""
    comm = _Comm(device)
    tasks = plan_job(n_segments)
    world, rank = comm.world, comm.rank
    # owner of every task = owner of its leftmost segment; segment tasks are numbered in arrival order
    seg_no = {}
    owner, span = {}, {}
    k = 0
    for t in tasks:
        if t.command == CMD_SEGMENT:
            seg_no[t.task_number] = k
            owner[t.task_number] = owner_of(k, n_segments, world)
            span[t.task_number] = (k, k)
            k += 1
        elif t.command == CMD_JOIN:
            l, r = t.depends_on
            owner[t.task_number] = owner[l]
            span[t.task_number] = (span[l][0], span[r][1])
    have = {}
    stats = {"proved": 0, "joined": 0, "sent": 0, "received": 0}
    root = None
    if prove_and_lift_many is not None:
        mine = [t.task_number for t in tasks if t.command == CMD_SEGMENT and owner[t.task_number] == rank]
        for tn, rcpt in zip(mine, prove_and_lift_many([seg_no[tn] for tn in mine])):
            have[tn] = rcpt
        stats["proved"] = len(mine)
    for t in tasks:
        if t.command == CMD_SEGMENT:
            if owner[t.task_number] == rank and t.task_number not in have:
                have[t.task_number] = prove_and_lift(seg_no[t.task_number])
                stats["proved"] += 1
        elif t.command == CMD_JOIN:
            l, r = t.depends_on
            if owner[r] != owner[l]:
                if owner[r] == rank:
                    comm.send(seal_of(have.pop(r)), owner[l]); stats["sent"] += 1
                elif owner[l] == rank:
                    have[r] = receipt_from_seal(comm.recv(recursion_seal_words, owner[r]), span[r]); stats["received"] += 1
            if owner[l] == rank:
                have[t.task_number] = join(have.pop(l), have.pop(r))
                stats["joined"] += 1
        elif t.command == CMD_FINALIZE:
            top = t.depends_on[0]
            if owner[top] != root_rank:
                if owner[top] == rank:
                    comm.send(seal_of(have[top]), root_rank); stats["sent"] += 1
                elif rank == root_rank:
                    have[top] = receipt_from_seal(comm.recv(recursion_seal_words, owner[top]), span[top]); stats["received"] += 1
            if rank == root_rank:
                root = have[top]
    stats["bytes_sent"] = comm.bytes_moved
    return root, stats


def gather_seals(local_seals: List[np.ndarray], words: int, device=None) -> Optional[List[np.ndarray]]:
    """all_gather of each rank's fixed-size seals (the "gather leaves" step of BASELINE.json): returns every rank's
    seals in rank order on all ranks.  One collective per call; payloads are ~0.1-0.25 MB each."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [s.copy() for s in local_seals]
    n = len(local_seals)
    buf = torch.zeros((n, words), dtype=torch.int32, device=device or "cpu")
    for i, s in enumerate(local_seals):
        buf[i].copy_(torch.from_numpy(np.ascontiguousarray(s).view(np.int32)))
    out = [torch.empty_like(buf) for _ in range(dist.get_world_size())]
    dist.all_gather(out, buf)
    return [o[i].cpu().numpy().view(np.uint32).copy() for o in out for i in range(n)]
