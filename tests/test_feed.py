"""prove_stream / PinnedRing (boundless_b200/feed.py): scheduling logic on the CPU with a recording stand-in for the prover; the GPU twin
with real proofs is tests/test_gpu_prover.py::test_prove_stream_matches_the_oracle."""
import threading

import pytest

from boundless_b200.feed import PinnedRing, prove_stream


class Opts:
    def __init__(self, slots): self.slots = slots


class Recorder:
    """submit / prefetch / wait with the ProverServer rules: one proof per slot, wait only on a busy slot."""

    def __init__(self, slots, fail_on=None):
        self.opts, self.busy, self.log, self.fail_on, self.max_inflight = Opts(slots), {}, [], fail_on, 0

    def submit_segment(self, slot, seg):
        assert slot not in self.busy, "slot %d busy" % slot
        self.busy[slot] = seg
        self.max_inflight = max(self.max_inflight, len(self.busy))
        self.log.append(("submit", slot, seg))

    def prefetch_segment(self, slot, seg):
        assert slot in self.busy, "prefetch goes under a running proof"
        self.log.append(("prefetch", slot, seg))

    def wait(self, slot):
        seg = self.busy.pop(slot)
        self.log.append(("wait", slot, seg))
        if seg == self.fail_on:
            raise RuntimeError("proof of %r failed" % (seg,))
        return "receipt-%s" % seg


@pytest.mark.parametrize("n,slots", [(0, 2), (1, 4), (3, 4), (4, 4), (5, 2), (9, 4), (7, 1)])
def test_order_slots_and_prefetch(n, slots):
    srv = Recorder(slots)
    assert list(prove_stream(srv, range(n))) == ["receipt-%d" % i for i in range(n)]
    assert srv.max_inflight == min(n, slots) and not srv.busy
    subs = [(e[1], e[2]) for e in srv.log if e[0] == "submit"]
    assert subs == [(i % slots, i) for i in range(n)]
    # segment i + slots is prefetched on segment i's slot, immediately after segment i is submitted
    for k, e in enumerate(srv.log):
        if e[0] == "submit" and e[2] + slots < n:
            assert srv.log[k + 1] == ("prefetch", e[1], e[2] + slots)
    assert sum(1 for e in srv.log if e[0] == "prefetch") == max(0, n - slots)


def test_producer_is_pulled_lazily_and_failures_drain():
    pulled = []

    def producer():
        for i in range(100):
            pulled.append(i)
            yield i
    srv = Recorder(2)
    g = prove_stream(srv, producer())
    assert next(g) == "receipt-0"
    assert len(pulled) <= 2 + 3 + 1                       # the two in flight, slots + 1 ahead, one being submitted
    g.close()
    assert not srv.busy                                    # early exit leaves no slot busy
    srv = Recorder(2, fail_on=3)
    got = []
    with pytest.raises(RuntimeError, match="proof of 3 failed"):
        for r in prove_stream(srv, range(8)):
            got.append(r)
    assert got == ["receipt-0", "receipt-1", "receipt-2"] and not srv.busy
    with pytest.raises(ValueError):
        list(prove_stream(Recorder(0), [1]))


def test_pinned_ring_recycles_buffers_behind_the_consumer():
    bufs = [bytearray(8) for _ in range(4)]
    ring = PinnedRing(bufs, hold=2)
    seen = []

    def produce():
        for i in range(20):
            b = ring.acquire(timeout=10)
            b[0] = i
            ring.put(b, meta=i)
        ring.close()
    t = threading.Thread(target=produce)
    t.start()
    ids = set()
    for buf, meta in ring.items(timeout=10):
        assert buf[0] == meta                               # a buffer is never refilled while the consumer still holds it
        seen.append(meta); ids.add(id(buf))
    t.join()
    assert seen == list(range(20)) and ids == {id(b) for b in bufs}
    assert len(ring._free) == 4
    with pytest.raises(RuntimeError):
        ring.acquire(timeout=1)


def test_numa_helpers_parse_sysfs(tmp_path):
    """feed.gpu_numa_node / _parse_cpulist against a fake sysfs tree (the real lookup needs a GPU; bench.py records what it finds)."""
    from boundless_b200 import feed
    assert feed._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    d = tmp_path / "bus" / "pci" / "devices" / "0000:1b:00.0"
    d.mkdir(parents=True)
    (d / "numa_node").write_text("1\n")
    assert feed.gpu_numa_node("0000:1B:00.0", str(tmp_path)) == 1
    assert feed.gpu_numa_node("0000:ff:00.0", str(tmp_path)) == -1
