"""Multi-rank job logic on CPU (gloo, world_size 2 and 3): sharding, the Planner-shaped join DAG executed across ranks,
rank-to-rank receipt movement, and the seal all_gather.  The prover is replaced by a hash stub (no GPU here); the real
prover runs the same code path under NCCL in bench.py --mode tree on the B200 box."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORDS = 64


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _h(*parts):
    d = hashlib.sha256(b"|".join(parts)).digest() * 8
    return np.frombuffer(d[:WORDS * 4], dtype=np.uint32).copy()


class R:
    def __init__(self, seal, claim): self.seal, self.claim = seal, claim


def _expected_root(n):
    sys.path.insert(0, ROOT)
    from boundless_b200.dist import plan_job
    from boundless_b200.planner import CMD_JOIN, CMD_SEGMENT
    val, k = {}, 0
    for t in plan_job(n):
        if t.command == CMD_SEGMENT:
            val[t.task_number] = _h(b"seg", str(k).encode()); k += 1
        elif t.command == CMD_JOIN:
            l, r = t.depends_on
            val[t.task_number] = _h(b"join", val[l].tobytes(), val[r].tobytes())
        else:
            return val[t.depends_on[0]]


def _worker(rank, world, port, n_segments, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from boundless_b200.dist import gather_seals, prove_job, shard_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(n_segments, rank, world)
        proved = []
        def prove_and_lift(i):
            assert lo <= i < hi, "segment %d proved on the wrong rank" % i
            proved.append(i)
            return R(_h(b"seg", str(i).encode()), (i, i))
        def join(a, b):
            assert a.claim[1] + 1 == b.claim[0]
            return R(_h(b"join", a.seal.tobytes(), b.seal.tobytes()), (a.claim[0], b.claim[1]))
        root, stats = prove_job(n_segments, prove_and_lift, join, lambda r: r.seal, lambda s, c: R(s, c), WORDS)
        # every rank contributes ALL its seals: the counts differ whenever n_segments % world != 0
        gathered = gather_seals([_h(b"seg", str(i).encode()) for i in range(lo, hi)], WORDS)
        q.put((rank, None if root is None else (root.seal.tobytes(), root.claim), stats, proved, [g.tobytes() for g in gathered]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_segments", [(2, 8), (2, 5), (3, 7), (2, 1)])
def test_prove_job_across_ranks(b200lib, world, n_segments):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_segments, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        rank, root, stats, proved, gathered = q.get(timeout=120)
        res[rank] = (root, stats, proved, gathered)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the root lands on rank 0 only, covers all segments, and equals the single-process evaluation of the same DAG
    assert res[0][0] is not None and all(res[r][0] is None for r in range(1, world))
    seal_bytes, claim = res[0][0]
    assert claim == (0, n_segments - 1)
    assert seal_bytes == _expected_root(n_segments).tobytes()
    # every segment proved exactly once, on its owner; joins add up; each cross-rank edge is one send and one receive
    allp = sorted(i for r in res for i in res[r][2])
    assert allp == list(range(n_segments))
    assert sum(res[r][1]["joined"] for r in res) == n_segments - 1
    assert sum(res[r][1]["sent"] for r in res) == sum(res[r][1]["received"] for r in res)
    if n_segments >= world:
        assert sum(res[r][1]["sent"] for r in res) >= world - 1
    # all_gather returns every rank's seals in rank order on all ranks, uneven shards included
    assert all(res[r][3] == res[0][3] for r in res)
    assert res[0][3] == [_h(b"seg", str(i).encode()).tobytes() for i in range(n_segments)]


def test_shard_bounds_and_owner():
    sys.path.insert(0, ROOT)
    from boundless_b200.dist import owner_of, shard_bounds
    for n in (1, 7, 8, 256, 257):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                lo, hi = shard_bounds(n, r, world)
                covered += list(range(lo, hi))
                for i in range(lo, hi):
                    assert owner_of(i, n, world) == r
            assert covered == list(range(n))


# ---- the asynchronous runner (JobRunner): several tasks in flight per rank, device-resident receipts, control messages on gloo ----
class _StubEngine:
    """Stands in for B200Engine: `slots` tasks in flight, each "completes" a few polls after submission (so completion order differs
    from submission order), receipts are CPU tensors of WORDS int32 words."""

    def __init__(self, slots, lo, hi):
        import torch
        self.torch, self.slots, self.lo, self.hi = torch, slots, lo, hi
        self.pending, self.proved, self.max_running = {}, [], 0

    class Rec:
        def __init__(self, buf, kind, claim): self.owner, self.kind, self.claim = buf, kind, tuple(claim)

    def new_buffer(self): return self.torch.zeros(WORDS, dtype=self.torch.int32)
    def release(self, rec): rec.owner = None
    def tensor_of(self, rec): return rec.owner

    def _fill(self, out, arr): out.copy_(self.torch.from_numpy(arr.view(np.int32)))

    def submit_segment(self, slot, index, out):
        assert self.lo <= index < self.hi, "segment %d proved on the wrong rank" % index
        self.proved.append(index)
        self._fill(out, _h(b"seg", str(index).encode()))
        self.pending[slot] = [3 + (index * 7) % 5, self.Rec(out, 1, (index, index))]
        self.max_running = max(self.max_running, len(self.pending))

    def submit_join(self, slot, a, b, out):
        assert a.claim[1] + 1 == b.claim[0]
        self._fill(out, _h(b"join", a.owner.numpy().view(np.uint32).tobytes(), b.owner.numpy().view(np.uint32).tobytes()))
        self.pending[slot] = [2, self.Rec(out, 2, (a.claim[0], b.claim[1]))]
        self.max_running = max(self.max_running, len(self.pending))

    def query(self, slot):
        self.pending[slot][0] -= 1
        return self.pending[slot][0] <= 0

    def finish(self, slot): return self.pending.pop(slot)[1]
    def receipt_from_buffer(self, buf, kind, claim): return self.Rec(buf, kind, claim)


def _async_worker(rank, world, port, n_segments, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from boundless_b200.dist import JobRunner, shard_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(n_segments, rank, world)
        eng = _StubEngine(3, lo, hi)
        root, stats = JobRunner(eng, n_segments, max_segments_in_flight=2).run()
        q.put((rank, None if root is None else (root.owner.numpy().view(np.uint32).tobytes(), root.claim), stats, eng.proved, eng.max_running))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_segments", [(2, 8), (2, 5), (3, 7), (2, 1), (3, 16)])
def test_async_job_runner_across_ranks(b200lib, world, n_segments):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_async_worker, args=(r, world, port, n_segments, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        rank, root, stats, proved, max_running = q.get(timeout=120)
        res[rank] = (root, stats, proved, max_running)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][0] is not None and all(res[r][0] is None for r in range(1, world))
    seal_bytes, claim = res[0][0]
    assert claim == (0, n_segments - 1)
    assert seal_bytes == _expected_root(n_segments).tobytes()           # the same root as the single-process evaluation of the DAG
    assert sorted(i for r in res for i in res[r][2]) == list(range(n_segments))
    assert sum(res[r][1]["joined"] for r in res) == n_segments - 1
    assert sum(res[r][1]["sent"] for r in res) == sum(res[r][1]["received"] for r in res)
    if n_segments >= 2 * world:
        assert max(res[r][3] for r in res) >= 2                           # several tasks really were in flight on one rank


def test_async_job_runner_single_process(b200lib):
    """world 1 (no process group): the runner reduces 9 segments to the Planner's root with 3 slots busy."""
    sys.path.insert(0, ROOT)
    from boundless_b200.dist import JobRunner
    eng = _StubEngine(3, 0, 9)
    root, stats = JobRunner(eng, 9, max_segments_in_flight=3).run()
    assert root.claim == (0, 8) and stats["proved"] == 9 and stats["joined"] == 8 and stats["sent"] == 0
    assert root.owner.numpy().view(np.uint32).tobytes() == _expected_root(9).tobytes()
    assert eng.max_running == 3
    # max_segments_in_flight caps the Prove tasks that run at once (the other slots are kept for lifts and joins); same root
    eng2 = _StubEngine(4, 0, 9)
    seg_running = []
    orig_submit, orig_finish = eng2.submit_segment, eng2.finish
    live = set()
    def submit_segment(slot, index, out):
        live.add(slot); seg_running.append(len(live)); orig_submit(slot, index, out)
    def finish(slot):
        live.discard(slot); return orig_finish(slot)
    eng2.submit_segment, eng2.finish = submit_segment, finish
    root2, _ = JobRunner(eng2, 9, max_segments_in_flight=2).run()
    assert max(seg_running) == 2 and root2.owner.numpy().view(np.uint32).tobytes() == _expected_root(9).tobytes()
