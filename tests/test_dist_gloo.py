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
        gathered = gather_seals([_h(b"seg", str(i).encode()) for i in range(lo, lo + 1)] if hi > lo else [_h(b"none")], WORDS)
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
    # all_gather returns every rank's seal in rank order on all ranks
    assert all(res[r][3] == res[0][3] for r in res) and len(res[0][3]) == world


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
