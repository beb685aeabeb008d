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


# ---- asynchronous form: every slot of the GPU busy, receipts device-resident, joins as soon as their inputs exist -----------------------
class B200Engine:
    """What JobRunner drives: the ProverServer's composite tasks (tasks::prove::prover and tasks::join::join as single enqueues, receipts
    in device memory).  Buffers are torch int32 tensors on the GPU so that torch.distributed (NCCL) can move them as they are."""

    def __init__(self, srv, make_segment: Callable[[int], object], device, verify: bool = True):
        import torch
        from .prover_server import KIND_JOIN, KIND_LIFT
        self.torch, self.srv, self.make_segment, self.device, self.verify = torch, srv, make_segment, device, verify
        self.slots = srv.opts.slots
        self.words = srv.seal_words(srv._rec_circuit(KIND_LIFT))
        self.KIND_JOIN = KIND_JOIN
        self._pool = []            # receipt buffers returned by release(), reused by new_buffer()
        self._out = {}             # slot -> the buffer its running task writes its receipt to

    def new_buffer(self):
        return self._pool.pop() if self._pool else self.torch.empty(self.words, dtype=self.torch.int32, device=self.device)

    def release(self, receipt):
        if receipt.owner is not None:
            self._pool.append(receipt.owner)
            receipt.owner = None

    def tensor_of(self, receipt):
        return receipt.owner

    def submit_segment(self, slot, index, out):
        self.srv.submit_prove_lift(slot, self.make_segment(index), d_out=out.data_ptr(), verify=self.verify, host_seals=False)
        self._out[slot] = out

    def submit_join(self, slot, a, b, out):
        self.srv.submit_recursion_dev(slot, self.KIND_JOIN, a, b, d_out=out.data_ptr(), verify=self.verify, host_seal=False)
        self._out[slot] = out

    def query(self, slot):
        return self.srv.query(slot)

    def finish(self, slot):
        r = self.srv.wait_task(slot)
        rec = r[1] if isinstance(r, tuple) else r
        rec.owner = self._out.pop(slot)
        return rec

    def receipt_from_buffer(self, buf, kind, claim):
        from .prover_server import DeviceReceipt
        return DeviceReceipt(buf.data_ptr(), self.words, kind, tuple(claim), [], buf, None)

    def to_host(self, receipt) -> np.ndarray:
        return receipt.owner.cpu().numpy().view(np.uint32).copy()


class _Pending:
    """Completion of a torch.distributed Work without blocking the event loop.  NCCL work is polled (a CUDA event query); gloo work never
    reports completion to a poll -- its is_completed() only turns true inside wait() -- so a helper thread blocks in wait() instead."""

    def __init__(self, work, poll: bool):
        self.work, self._done = work, False
        if not poll:
            import threading
            threading.Thread(target=self._wait, daemon=True).start()
        self._poll = poll

    def _wait(self):
        self.work.wait()
        self._done = True

    def done(self):
        if self._poll and not self._done:
            self._done = self.work.is_completed()
        return self._done


class _Link:
    """Receipt movement between ranks.  A small control message on a gloo group (CPU, costs no SM) announces "receipt t is ready, you are
    its consumer" together with the receipt's metadata; only then do both sides post the matching send / recv of the device tensor on
    the data group (NCCL over NVLink on GPUs), so no NCCL kernel sits spinning on a GPU while its peer is still proving.  Per (src, dst)
    pair the data transfers are posted in the sender's sequence order on both sides (NCCL matches point-to-point operations by order)."""
    _ctrl_cache = {}

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.rank = dist.get_rank() if self.on else 0
        self.world = dist.get_world_size() if self.on else 1
        self.ctrl = None
        self.nccl = bool(self.on and dist.get_backend() != "gloo")
        if self.nccl:
            key = id(dist.group.WORLD)
            if key not in _Link._ctrl_cache:
                _Link._ctrl_cache[key] = dist.new_group(backend="gloo")      # collective: every rank constructs its _Link
            self.ctrl = _Link._ctrl_cache[key]
        self.bytes_moved = 0
        self.send_seq = {}            # dst -> next sequence number
        self.recv_seq = {}            # src -> next sequence number to post
        self.pending_sends = []       # (completions, keepalive)
        self.flags = {}               # task -> (src, tensor, completion)
        self.arrived = {}             # src -> {seq: (task, kind, lo, hi)}
        self.data = {}                # task -> (completion, buffer, kind, claim)

    def expect(self, task, src):
        t = self.torch.zeros(5, dtype=self.torch.int64)
        self.flags[task] = (src, t, _Pending(self.dist.irecv(t, src, group=self.ctrl, tag=task), False))

    def send(self, task, dst, tensor, kind, claim):
        seq = self.send_seq.get(dst, 0); self.send_seq[dst] = seq + 1
        flag = self.torch.tensor([task, seq, kind, claim[0], claim[1]], dtype=self.torch.int64)
        w1 = _Pending(self.dist.isend(flag, dst, group=self.ctrl, tag=task), False)
        # the data tag keeps control and data apart when both travel on the same (gloo) group; NCCL ignores tags
        w2 = _Pending(self.dist.isend(tensor, dst, tag=(1 << 20) + task), self.nccl)
        self.pending_sends.append(((w1, w2), (flag, tensor)))
        self.bytes_moved += tensor.numel() * 4

    def poll(self, new_buffer):
        """-> list of (task, buffer, kind, claim) whose data has fully arrived"""
        for task in [t for t, (_, _, w) in self.flags.items() if w.done()]:
            src, t, _ = self.flags.pop(task)
            tk, seq, kind, lo, hi = (int(x) for x in t)
            self.arrived.setdefault(src, {})[seq] = (tk, kind, lo, hi)
        for src, slots in self.arrived.items():
            while self.recv_seq.get(src, 0) in slots:
                tk, kind, lo, hi = slots.pop(self.recv_seq.get(src, 0))
                self.recv_seq[src] = self.recv_seq.get(src, 0) + 1
                buf = new_buffer()
                w = _Pending(self.dist.irecv(buf, src, tag=(1 << 20) + tk), self.nccl)
                self.data[tk] = (w, buf, kind, (lo, hi))
        out = []
        for t in [t for t, (w, _, _, _) in self.data.items() if w.done()]:
            _, buf, kind, claim = self.data.pop(t)
            out.append((t, buf, kind, claim))
        self.pending_sends = [(ws, keep) for ws, keep in self.pending_sends if not all(w.done() for w in ws)]
        return out


class JobRunner:
    """One job = the reference Planner's DAG over n segments (segment -> prove + lift; join as soon as two peaks merge; finalize on the
    root rank), executed by `world` ranks with EVERY slot of every GPU kept busy: a rank launches whichever of its tasks is ready
    and oldest (the taskdb's claim order: creation order, so joins run before later segments -- depth-first reduction,
    redis_backend.rs:288-338), several at a time, and hands a finished receipt to the rank that owns its consumer.  Ownership is the
    one of prove_job above: segment i -> contiguous blocks, a join -> the owner of its LEFT input.

    engine: .slots, .new_buffer(), .submit_segment(slot, index, out), .submit_join(slot, a, b, out), .query(slot), .finish(slot) ->
    receipt (with .kind, .claim), .tensor_of(receipt), .receipt_from_buffer(buf, kind, claim), .release(receipt)."""

    def __init__(self, engine, n_segments: int, root_rank: int = 0, max_segments_in_flight: Optional[int] = None):
        """max_segments_in_flight: how many Prove tasks may run at once on this rank (default: every slot).  Fewer of them in flight
        make the first receipts exist earlier, so that lifts and joins run under the later segments -- but a single 2^20 proof runs at
        53 ms against 51 ms per proof with four in flight, and measured on one B200 that loses more than the shorter tail wins
        (4 segments to root: 305 / 291 / 291 / 283 ms with 1 / 2 / 3 / 4 in flight; profiles/tree_inflight_r02.txt)."""
        self.engine, self.n, self.root_rank = engine, n_segments, root_rank
        self.max_seg = engine.slots if max_segments_in_flight is None else max(1, int(max_segments_in_flight))
        self.link = _Link()
        self.rank, self.world = self.link.rank, self.link.world
        self.tasks = plan_job(n_segments)
        self.owner, self.span, self.seg_no, self.consumer = {}, {}, {}, {}
        k = 0
        for t in self.tasks:
            if t.command == CMD_SEGMENT:
                self.seg_no[t.task_number] = k
                self.owner[t.task_number] = owner_of(k, n_segments, self.world)
                self.span[t.task_number] = (k, k); k += 1
            elif t.command == CMD_JOIN:
                l, r = t.depends_on
                self.owner[t.task_number] = self.owner[l]
                self.span[t.task_number] = (self.span[l][0], self.span[r][1])
                self.consumer[l] = self.consumer[r] = t
            elif t.command == CMD_FINALIZE:
                self.owner[t.task_number] = root_rank
                self.consumer[t.depends_on[0]] = t
        self.stats = {"proved": 0, "joined": 0, "sent": 0, "received": 0, "max_in_flight": 0}

    def run(self):
        import heapq
        import time
        eng, link, me = self.engine, self.link, self.rank
        by_no = {t.task_number: t for t in self.tasks}
        have, ready, running = {}, [], {}
        free = list(range(eng.slots))
        todo = 0
        root = None
        for t in self.tasks:
            if t.command == CMD_SEGMENT and self.owner[t.task_number] == me:
                heapq.heappush(ready, t.task_number); todo += 1
            elif t.command == CMD_JOIN and self.owner[t.task_number] == me:
                todo += 1
                for d in t.depends_on:
                    if self.owner[d] != me:
                        link.expect(d, self.owner[d])
            elif t.command == CMD_FINALIZE and me == self.root_rank and self.owner[t.depends_on[0]] != me:
                link.expect(t.depends_on[0], self.owner[t.depends_on[0]])
        need_root = me == self.root_rank

        def available(tn, rec):
            nonlocal root
            c = self.consumer[tn]
            dst = self.owner[c.task_number]
            if dst != me:
                link.send(tn, dst, eng.tensor_of(rec), rec.kind, rec.claim); self.stats["sent"] += 1
                have[tn] = rec                        # keeps the tensor alive until the job ends
                return
            have[tn] = rec
            if c.command == CMD_FINALIZE:
                root = rec
            elif all(d in have for d in c.depends_on):
                heapq.heappush(ready, c.task_number)

        while todo or running or (need_root and root is None) or link.pending_sends:
            progressed = False
            for slot in [s for s in running if eng.query(s)]:
                tn = running.pop(slot)
                rec = eng.finish(slot)
                free.append(slot); todo -= 1; progressed = True
                t = by_no[tn]
                if t.command == CMD_JOIN:
                    self.stats["joined"] += 1
                    for d in t.depends_on:
                        eng.release(have.pop(d))
                else:
                    self.stats["proved"] += 1
                available(tn, rec)
            if link.on:
                for tn, buf, kind, claim in link.poll(eng.new_buffer):
                    self.stats["received"] += 1; progressed = True
                    available(tn, eng.receipt_from_buffer(buf, kind, claim))
            held = []
            while free and ready:
                tn = heapq.heappop(ready)
                t = by_no[tn]
                if t.command == CMD_SEGMENT and sum(1 for x in running.values() if by_no[x].command == CMD_SEGMENT) >= self.max_seg:
                    held.append(tn)                       # enough segment proofs in flight: keep the slot for recursion work
                    continue
                slot = free.pop(0)
                if t.command == CMD_SEGMENT:
                    eng.submit_segment(slot, self.seg_no[tn], eng.new_buffer())
                else:
                    l, r = t.depends_on
                    eng.submit_join(slot, have[l], have[r], eng.new_buffer())
                running[slot] = tn; progressed = True
                self.stats["max_in_flight"] = max(self.stats["max_in_flight"], len(running))
            for tn in held:
                heapq.heappush(ready, tn)
            if not progressed:
                time.sleep(0)
        self.stats["bytes_sent"] = link.bytes_moved
        return root, self.stats


def gather_seals(local_seals: List[np.ndarray], words: int, device=None) -> Optional[List[np.ndarray]]:
    """all_gather of each rank's fixed-size seals (the "gather leaves" step of BASELINE.json): returns every rank's
    seals in rank order on all ranks.  Ranks may hold different numbers of seals (shard_bounds gives the first n % world ranks one
    more): the counts are gathered first and the payload is padded to the largest.  Payloads are ~0.1-0.25 MB each."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [s.copy() for s in local_seals]
    world = dist.get_world_size()
    dev = device or "cpu"
    cnt = torch.tensor([len(local_seals)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c) for c in counts]
    n = max(max(counts), 1)
    buf = torch.zeros((n, words), dtype=torch.int32, device=dev)
    for i, s in enumerate(local_seals):
        buf[i].copy_(torch.from_numpy(np.ascontiguousarray(s).view(np.int32)))
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return [o[i].cpu().numpy().view(np.uint32).copy() for o, c in zip(out, counts) for i in range(c)]
