"""Streaming segments into the prover: the consumer side of the executor -> GPU hand-off.

In the reference the executor emits `Segment`s while it runs (`tasks/executor.rs:721-757`, `segment_limit_po2`) and every segment
travels through Redis to whichever GPU agent claims its `prove` task (`tasks/prove.rs:24-40`).  On one box the hand-off can be a
bounded set of pinned host buffers instead: the producer fills a buffer, `prove_stream` keeps `slots` proofs in
flight on the GPU, and each slot's NEXT witness is copied host -> device under the proof that slot is running
(`ProverServer.prefetch_segment` / `b200_prefetch_trace_async`), so the 940 MB copy per 2^20-row segment never sits on the critical path.
"""
import collections
import threading
from typing import Iterable, Iterator, Optional


def prove_stream(srv, segments: Iterable, slots: Optional[int] = None) -> Iterator:
    """Yield one SegmentReceipt per segment, in input order, with up to `slots` proofs in flight.

    Segment i runs on slot i % slots; right after it is submitted, the witness of segment i + slots (the next one for that slot) starts
    its prefetch.  `segments` may be a generator: at most slots + 1 segments are pulled ahead of the one being submitted.
    `srv` is a ProverServer (or anything with submit_segment / prefetch_segment / wait and `opts.slots`)."""
    slots = int(slots or srv.opts.slots)
    if slots < 1:
        raise ValueError("slots must be positive")
    it = iter(segments)
    ahead = collections.deque()          # pulled from the producer, not yet submitted
    inflight = collections.deque()       # slots in submission order
    done = False
    i = 0
    try:
        while True:
            while not done and len(ahead) < slots + 1:
                try:
                    ahead.append(next(it))
                except StopIteration:
                    done = True
            if not ahead:
                break
            slot = i % slots
            if len(inflight) == slots:
                yield srv.wait(inflight.popleft())        # the oldest proof in flight is the one on `slot`
            seg = ahead.popleft()
            srv.submit_segment(slot, seg)
            inflight.append(slot)
            if len(ahead) >= slots:
                srv.prefetch_segment(slot, ahead[slots - 1])   # segment i + slots: the next witness for this slot
            i += 1
        while inflight:
            yield srv.wait(inflight.popleft())
    finally:
        while inflight:                                    # consumer stopped early or a proof failed: leave no slot busy
            try:
                srv.wait(inflight.popleft())
            except Exception:                              # noqa: BLE001
                pass


class PinnedRing:
    """A bounded ring of pinned host witness buffers shared by one producer (the executor side) and one consumer (`prove_stream`).

    acquire() blocks while every buffer is still owned by a proof that has not finished; items() hands out filled buffers in order.
    A buffer goes back to the producer with release(), which items() does automatically `hold` items later.  Feeding prove_stream, use
    hold = 2 * slots + 1 (the proof of item i has been waited for once item i + slots is submitted, and prove_stream pulls slots + 1
    items ahead of that) and at least hold + 2 buffers."""

    def __init__(self, buffers, hold: int):
        self._free = collections.deque(buffers)
        self._filled = collections.deque()
        self._cv = threading.Condition()
        self._closed = False
        self.hold = int(hold)

    def acquire(self, timeout: Optional[float] = None):
        """Producer: a buffer to fill (blocks until one is free)."""
        with self._cv:
            if not self._cv.wait_for(lambda: self._free or self._closed, timeout):
                raise TimeoutError("no free witness buffer")
            if self._closed:
                raise RuntimeError("ring closed")
            return self._free.popleft()

    def put(self, buf, meta=None):
        """Producer: `buf` is filled; `meta` travels with it (e.g. the Segment header)."""
        with self._cv:
            self._filled.append((buf, meta))
            self._cv.notify_all()

    def close(self):
        """Producer: no more segments."""
        with self._cv:
            self._closed = True
            self._cv.notify_all()

    def release(self, buf):
        with self._cv:
            self._free.append(buf)
            self._cv.notify_all()

    def items(self, timeout: Optional[float] = None):
        """Consumer: (buf, meta) pairs in order until the producer closes the ring; buffers are recycled `hold` items later."""
        held = collections.deque()
        while True:
            with self._cv:
                if not self._cv.wait_for(lambda: self._filled or self._closed, timeout):
                    raise TimeoutError("producer stalled")
                if not self._filled:
                    break
                buf, meta = self._filled.popleft()
            held.append(buf)
            if len(held) > self.hold:
                self.release(held.popleft())
            yield buf, meta
        # the consumer drains prove_stream before dropping the generator, so the remaining buffers are free again
        while held:
            self.release(held.popleft())


# ---- NUMA placement of the pinned witness buffers ---------------------------------------------------------------------------------
def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(pci_bus_id: str, sysfs: str = "/sys") -> int:
    """NUMA node of the PCI device (sysfs numa_node), -1 when the platform does not say (single node, most VMs)."""
    import os
    for name in (pci_bus_id.lower(), pci_bus_id.lower().split(":", 1)[-1] if pci_bus_id.count(":") == 2 else pci_bus_id.lower()):
        for dom in ("", "0000:"):
            path = os.path.join(sysfs, "bus", "pci", "devices", dom + name, "numa_node")
            try:
                with open(path) as f:
                    return int(f.read().strip())
            except (OSError, ValueError):
                continue
    return -1


def bind_to_gpu_numa_node(device: int, L=None, sysfs: str = "/sys") -> dict:
    """Restrict this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned witness buffers it allocates afterwards
    (first touch) and the threads that fill them sit next to the GPU's PCIe root: with one process per GPU (compose.yml:113) every rank
    then streams its 940 MB witnesses from its own memory controller instead of all ranks sharing node 0.  Returns what was found and
    done; a platform that reports no NUMA node (-1) or a single node is left alone."""
    import os
    from . import lib as _lib
    L = L or _lib.load()
    info = {"device": device, "pci_bus_id": None, "gpu_numa_node": -1, "nodes_online": None, "bound": False, "cpus": None}
    import ctypes as C
    buf = C.create_string_buffer(32)
    if L.b200_device_pci_bus_id(device, buf, 32) is not None:
        return info
    info["pci_bus_id"] = buf.value.decode()
    node = gpu_numa_node(info["pci_bus_id"], sysfs)
    info["gpu_numa_node"] = node
    try:
        with open(os.path.join(sysfs, "devices", "system", "node", "online")) as f:
            info["nodes_online"] = f.read().strip()
    except OSError:
        pass
    if node < 0 or info["nodes_online"] in (None, "0"):
        return info
    try:
        with open(os.path.join(sysfs, "devices", "system", "node", "node%d" % node, "cpulist")) as f:
            cpus = _parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"], info["cpus"] = True, len(allowed)
    except (OSError, ValueError):
        pass
    return info
