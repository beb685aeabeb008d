"""Seal-level parity: the CUDA pipeline behind the ProverServer mirror must emit, word for word, the seal the
CPU oracle emits for the same segment, and that seal must pass the oracle's independent verifier.

Mirrors the reference's own test pattern, "verify after every step" (tasks/prove.rs:56-58,106-108; join.rs:77-79),
with bit-exactness added on top (the reference pins validity only).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_server(gpu):
    from boundless_b200 import ProverOpts, get_prover_server
    srv = get_prover_server(ProverOpts(segment_po2=14, recursion_po2=12, slots=2))
    yield srv
    srv.close()


@pytest.mark.parametrize("po2", [9, 10, 11, 12, 13, 14])
def test_segment_seal_bit_exact(small_server, oracle, po2):
    from boundless_b200 import Segment, VerifierContext
    seg = Segment(index=po2, po2=po2)
    rcpt = small_server.prove_segment(VerifierContext(), seg)
    ref = oracle.prove(po2, seg.seed)
    assert rcpt.seal.size == ref.size
    bad = np.nonzero(rcpt.seal != ref)[0]
    assert bad.size == 0, "first mismatch at word %d of %d" % (bad[0], ref.size)
    assert oracle.verify(rcpt.seal) == 0


def test_segment_from_host_trace(small_server, oracle):
    """The reference-facing form: the witness arrives in host memory (H2D inside the call)."""
    from boundless_b200 import Segment, VerifierContext
    po2 = 12
    seed = 0xB2000000 + 77
    trace = oracle.gen_trace(seed, po2, 16 + 208)
    rcpt = small_server.prove_segment(VerifierContext(), Segment(index=77, po2=po2, seed=seed, trace=trace))
    assert np.array_equal(rcpt.seal, oracle.prove(po2, seed))
    # a different witness under the same seed gives a different, still valid, seal
    trace2 = trace.copy(); trace2[5] = (int(trace2[5]) + 1) % oracle.P
    rcpt2 = small_server.prove_segment(VerifierContext(), Segment(index=77, po2=po2, seed=seed, trace=trace2))
    assert not np.array_equal(rcpt2.seal, rcpt.seal)
    assert np.array_equal(rcpt2.seal, oracle.prove(po2, seed, trace=trace2))
    assert oracle.verify(rcpt2.seal) == 0


def test_prefetched_witness_gives_the_same_seal(small_server, oracle):
    """b200_prefetch_trace_async: the next witness is copied on the copy stream into the slot's second coefficient region while the
    slot is proving; seals are bit-identical to the plain path, the regions alternate, and an unrelated trace is not confused."""
    from boundless_b200 import Segment
    po2, srv = 12, small_server
    seeds = [0xB2000000 + 900 + i for i in range(4)]
    traces = [oracle.gen_trace(sd, po2, 16 + 208) for sd in seeds]
    segs = [Segment(index=900 + i, po2=po2, seed=sd, trace=t) for i, (sd, t) in enumerate(zip(seeds, traces))]
    refs = [oracle.prove(po2, sd) for sd in seeds]
    seals = []
    srv.submit_segment(0, segs[0])
    for i in range(1, 4):
        srv.prefetch_segment(0, segs[i])             # under the running proof
        seals.append(srv.wait(0).seal)
        srv.submit_segment(0, segs[i])               # same buffer -> picks the staged copy up
    seals.append(srv.wait(0).seal)
    for got, ref in zip(seals, refs):
        assert np.array_equal(got, ref)
    # a prefetch that is NOT followed by the matching submit is simply dropped
    srv.prefetch_segment(0, segs[1])
    other = Segment(index=950, po2=po2, seed=seeds[2], trace=traces[2].copy())
    srv.submit_segment(0, other)
    assert np.array_equal(srv.wait(0).seal, refs[2])
    srv.submit_segment(0, segs[3])
    assert np.array_equal(srv.wait(0).seal, refs[3])


def test_prove_stream_matches_the_oracle(small_server, oracle):
    """feed.prove_stream: segments arrive from a generator with host witnesses, two proofs in flight, next witness prefetched."""
    from boundless_b200 import Segment
    from boundless_b200.feed import prove_stream
    po2, n = 10, 5
    seeds = [0xB2000000 + 700 + i for i in range(n)]

    def producer():
        for i, sd in enumerate(seeds):
            yield Segment(index=700 + i, po2=po2, seed=sd, trace=oracle.gen_trace(sd, po2, 16 + 208))
    receipts = list(prove_stream(small_server, producer()))
    assert [r.index for r in receipts] == [700 + i for i in range(n)]
    for r, sd in zip(receipts, seeds):
        assert np.array_equal(r.seal, oracle.prove(po2, sd))


def test_lift_join_bit_exact(small_server, oracle):
    from boundless_b200 import Segment, VerifierContext
    from boundless_b200.prover_server import KIND_JOIN, KIND_LIFT, RECURSION_WIDTHS
    ctx = VerifierContext()
    rp = small_server.opts.recursion_po2
    segs = [small_server.prove_segment(ctx, Segment(index=i, po2=10)) for i in range(2)]
    lifted = [small_server.lift(s) for s in segs]
    for s, l in zip(segs, lifted):
        d = oracle.seal_digest(s.seal)
        seed = int(d[0]) | (int(d[1]) << 32)
        ref = oracle.prove(rp, seed, *RECURSION_WIDTHS, kind=KIND_LIFT, input_digest=d)
        assert np.array_equal(l.seal, ref)
        assert oracle.verify(l.seal) == 0
    j = small_server.join(lifted[0], lifted[1])
    d = oracle.hash_pair(oracle.seal_digest(lifted[0].seal), oracle.seal_digest(lifted[1].seal))
    seed = int(d[0]) | (int(d[1]) << 32)
    ref = oracle.prove(rp, seed, *RECURSION_WIDTHS, kind=KIND_JOIN, input_digest=d)
    assert np.array_equal(j.seal, ref)
    assert oracle.verify(j.seal) == 0
    assert j.claim == (0, 1)


def test_job_dag_with_real_proofs(small_server, oracle):
    """BASELINE config 4 on one GPU: 3 segments -> prove + lift -> Planner-shaped joins -> root; every receipt on the
    way is recomputed by the oracle (tasks/prove.rs:44-104 + tasks/join.rs:52-56 executed through dist.prove_job)."""
    from boundless_b200 import Segment, VerifierContext
    from boundless_b200.dist import prove_job
    from boundless_b200.prover_server import KIND_JOIN, KIND_LIFT, RECURSION_WIDTHS, SuccinctReceipt
    ctx = VerifierContext()
    rp = small_server.opts.recursion_po2
    root, stats = prove_job(3, lambda i: small_server.lift(small_server.prove_segment(ctx, Segment(index=i, po2=9))),
                            small_server.join, lambda r: r.seal, lambda s, c: SuccinctReceipt(s, KIND_JOIN, c),
                            small_server.seal_words(small_server._rec_circuit(KIND_LIFT)))
    assert stats == {"proved": 3, "joined": 2, "sent": 0, "received": 0, "bytes_sent": 0}
    assert root.claim == (0, 2)
    # the batched / pipelined form (all slots busy) yields the same root
    root2, _ = prove_job(3, None, small_server.join, lambda r: r.seal, lambda s, c: SuccinctReceipt(s, KIND_JOIN, c),
                         small_server.seal_words(small_server._rec_circuit(KIND_LIFT)),
                         prove_and_lift_many=lambda idx: small_server.prove_and_lift_many([Segment(index=i, po2=9) for i in idx]))
    assert np.array_equal(root2.seal, root.seal)
    def rec(kind, digest):
        return oracle.prove(rp, int(digest[0]) | (int(digest[1]) << 32), *RECURSION_WIDTHS, kind=kind, input_digest=digest)
    lifts = [rec(KIND_LIFT, oracle.seal_digest(oracle.prove(9, 0xB2000000 + i))) for i in range(3)]
    j01 = rec(KIND_JOIN, oracle.hash_pair(oracle.seal_digest(lifts[0]), oracle.seal_digest(lifts[1])))
    j012 = rec(KIND_JOIN, oracle.hash_pair(oracle.seal_digest(j01), oracle.seal_digest(lifts[2])))
    assert np.array_equal(root.seal, j012)
    assert oracle.verify(root.seal) == 0


def test_two_slots_in_flight(small_server, oracle):
    from boundless_b200 import Segment
    a, b = Segment(index=100, po2=12), Segment(index=101, po2=13)
    small_server.submit_segment(0, a)
    small_server.submit_segment(1, b)
    rb = small_server.wait(1)
    ra = small_server.wait(0)
    assert np.array_equal(ra.seal, oracle.prove(12, a.seed))
    assert np.array_equal(rb.seal, oracle.prove(13, b.seed))


def test_errors_are_reported_not_fatal(small_server):
    from boundless_b200 import B200Error, Segment
    with pytest.raises(B200Error):
        small_server.submit_segment(0, Segment(index=0, po2=20))      # exceeds this server's max circuit
    with pytest.raises(B200Error):
        small_server.submit_segment(7, Segment(index=0, po2=10))      # no such slot
    small_server.submit_segment(0, Segment(index=0, po2=10))
    with pytest.raises(B200Error):
        small_server.submit_segment(0, Segment(index=1, po2=10))      # slot busy
    small_server.wait(0)


def _golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz"))


def _sha(a):
    import hashlib
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


@pytest.fixture(scope="module")
def full_server(gpu):
    """The BASELINE configuration: 2^20-row segments (16/208/32), recursion at 2^18 (16/128/16)."""
    from boundless_b200 import ProverOpts, get_prover_server
    srv = get_prover_server(ProverOpts(segment_po2=20, recursion_po2=18, slots=2))
    yield srv
    srv.close()


@pytest.mark.parametrize("po2", [16, 18, 20])
def test_segment_seal_bit_exact_at_baseline_sizes(full_server, oracle, po2):
    """north_star: the seal is bit-exact.  Word for word against the CPU oracle proving the same segment on the box's host cores, and
    against the sha256 pinned in tests/golden (generated by tests/golden/make_golden.py) -- at 2^16, 2^18 and the BASELINE config 2
    size 2^20 (the kernels that carry the headline: radix-32 strided NTT passes, 2^22 x 208 leaf hashing, 2^22-leaf trees)."""
    from boundless_b200 import Segment, VerifierContext
    seg = Segment(index=po2, po2=po2)
    rcpt = full_server.prove_segment(VerifierContext(), seg)
    assert np.array_equal(_sha(rcpt.seal), _golden()["seal_sha_po2_%d" % po2]), "seal differs from the pinned oracle seal"
    ref = oracle.prove(po2, seg.seed)
    assert rcpt.seal.size == ref.size
    bad = np.nonzero(rcpt.seal != ref)[0]
    assert bad.size == 0, "first mismatch at word %d of %d" % (bad[0], ref.size)
    assert oracle.verify(rcpt.seal) == 0
    full_server.verify_integrity(rcpt)          # the device-side verify_integrity accepts it too (tasks/prove.rs:56-58)


def test_lift_join_bit_exact_at_recursion_size(full_server, oracle):
    """lift and join at the real recursion size (po2 18, 16/128/16): two segments -> two lifts -> one join, each seal word for word
    against the oracle and against the pinned sha256 (tasks/prove.rs:96-108, tasks/join.rs:52-79)."""
    from boundless_b200 import Segment, VerifierContext
    from boundless_b200.prover_server import KIND_JOIN, KIND_LIFT, RECURSION_WIDTHS
    ctx = VerifierContext()
    g = _golden()
    segs = [full_server.prove_segment(ctx, Segment(index=40 + i, po2=12)) for i in range(2)]
    lifted = [full_server.lift(s) for s in segs]
    for i, (s, l) in enumerate(zip(segs, lifted)):
        d = oracle.seal_digest(s.seal)
        ref = oracle.prove(18, int(d[0]) | (int(d[1]) << 32), *RECURSION_WIDTHS, kind=KIND_LIFT, input_digest=d)
        assert np.array_equal(l.seal, ref)
        assert np.array_equal(_sha(l.seal), g["lift18_sha_%d" % i])
        assert oracle.verify(l.seal) == 0
    j = full_server.join(lifted[0], lifted[1])
    d = oracle.hash_pair(oracle.seal_digest(lifted[0].seal), oracle.seal_digest(lifted[1].seal))
    ref = oracle.prove(18, int(d[0]) | (int(d[1]) << 32), *RECURSION_WIDTHS, kind=KIND_JOIN, input_digest=d)
    assert np.array_equal(j.seal, ref)
    assert np.array_equal(_sha(j.seal), g["join18_sha"])
    assert oracle.verify(j.seal) == 0
    full_server.verify_integrity(j)


def test_full_size_segment_is_deterministic_and_lifts(full_server, oracle):
    """BASELINE config 2 at full size: the same segment proves to the same seal twice, and its lift at 2^18 verifies."""
    from boundless_b200 import Segment, VerifierContext
    r0 = full_server.prove_segment(VerifierContext(), Segment(index=0))
    r0b = full_server.prove_segment(VerifierContext(), Segment(index=0))
    assert np.array_equal(r0.seal, r0b.seal)
    assert oracle.verify(r0.seal) == 0
    l0 = full_server.lift(r0)
    assert oracle.verify(l0.seal) == 0


def test_po2_22_segment_bit_exact(gpu, oracle):
    """The first size whose evaluation matrix exceeds 2^32 words (272 columns x 2^24): every word of the seal against the oracle."""
    from boundless_b200 import ProverOpts, Segment, VerifierContext, get_prover_server
    srv = get_prover_server(ProverOpts(segment_po2=22, recursion_po2=18, slots=1))
    try:
        seg = Segment(index=22, po2=22)
        r = srv.prove_segment(VerifierContext(), seg)
        ref = oracle.prove(22, seg.seed)
        assert np.array_equal(r.seal, ref)
        srv.verify_integrity(r)
    finally:
        srv.close()


@pytest.mark.parametrize("po2", [23, 24])
def test_po2_23_24_segments_verify(gpu, oracle, po2):
    """Upstream MAX_CYCLES_PO2 = 24: segments of 2^23 and 2^24 rows (evaluation domains 2^25 / 2^26: three-pass transforms, Merkle
    trees of 2^25 / 2^26 leaves, 55 / 111 GB of device memory for the one slot).  The oracle's PROVER is not run at these sizes (its
    evaluation matrix alone is 37 / 73 GB of host memory); the seal must pass the oracle's verifier and the device verifier, be
    deterministic, and the kernels underneath are compared with the oracle at these sizes in tests/test_gpu_kernels.py."""
    from boundless_b200 import ProverOpts, Segment, VerifierContext, get_prover_server
    srv = get_prover_server(ProverOpts(segment_po2=po2, recursion_po2=18, slots=1))
    try:
        seg = Segment(index=po2, po2=po2)
        r = srv.prove_segment(VerifierContext(), seg)
        assert r.seal.size == oracle.seal_words(po2)
        assert oracle.verify(r.seal) == 0
        srv.verify_integrity(r)
        if po2 == 23:
            assert np.array_equal(srv.prove_segment(VerifierContext(), seg).seal, r.seal)
            l = srv.lift(r)
            assert oracle.verify(l.seal) == 0
    finally:
        srv.close()


def test_po2_21_segment_verifies(gpu, oracle):
    """compose.yml:67 runs the agents with --segment-po2 21: one 2^21-row segment (twice BASELINE config 2; 2^23-point evaluation
    domain, 2^23-leaf Merkle trees, the 2^23 check-polynomial iNTT) must pass the oracle's verifier."""
    from boundless_b200 import ProverOpts, Segment, VerifierContext, get_prover_server
    srv = get_prover_server(ProverOpts(segment_po2=21, recursion_po2=18, slots=1))
    try:
        r = srv.prove_segment(VerifierContext(), Segment(index=3, po2=21))
        assert r.seal.size == oracle.seal_words(21)
        assert oracle.verify(r.seal) == 0
    finally:
        srv.close()


# ---- the agent's task bodies as single enqueues over device-resident receipts (round 2) -----------------------------------------------
def _dev_buf(torch, words):
    return torch.zeros(words, dtype=torch.int32, device="cuda")


def test_prove_lift_composite_is_bit_exact(gpu, small_server, oracle):
    """b200_prove_lift_async = tasks::prove::prover (prove.rs:44-108) in one enqueue: both seals equal the oracle's, both verdicts are 0,
    and the lifted seal also lands in the caller's device buffer."""
    from boundless_b200 import Segment
    from boundless_b200.prover_server import KIND_LIFT, RECURSION_WIDTHS
    torch, srv = gpu, small_server
    rp = srv.opts.recursion_po2
    words = srv.seal_words(srv._rec_circuit(KIND_LIFT))
    bufs = [_dev_buf(torch, words) for _ in range(2)]
    segs = [Segment(index=300 + i, po2=11 + i) for i in range(2)]
    for slot, (seg, buf) in enumerate(zip(segs, bufs)):           # both slots in flight at once
        srv.submit_prove_lift(slot, seg, d_out=buf.data_ptr())
    for slot, (seg, buf) in enumerate(zip(segs, bufs)):
        assert isinstance(srv.query(slot), bool)
        seg_r, lift = srv.wait_task(slot)
        ref = oracle.prove(seg.po2, seg.seed)
        assert np.array_equal(seg_r.seal, ref)
        d = oracle.seal_digest(ref)
        ref_l = oracle.prove(rp, int(d[0]) | (int(d[1]) << 32), *RECURSION_WIDTHS, kind=KIND_LIFT, input_digest=d)
        assert np.array_equal(lift.seal, ref_l)
        assert np.array_equal(buf.cpu().numpy().view(np.uint32), ref_l)
        assert lift.kind == KIND_LIFT and lift.claim == (seg.index, seg.index) and lift.ptr == buf.data_ptr()
        assert srv.query(slot) is True


def test_join_over_device_receipts_is_bit_exact_and_verifies_its_inputs(gpu, small_server, oracle):
    """b200_recursion_verified_async = tasks::join::join (join.rs:41-79): verify left, verify right, join, verify the result, all on
    device-resident receipts.  A corrupted input is caught by the verification inside the task ([BENTO-JOIN-003/004])."""
    from boundless_b200 import Segment
    from boundless_b200.prover_server import KIND_JOIN, KIND_LIFT, RECURSION_WIDTHS, DeviceReceipt, VerificationError
    torch, srv = gpu, small_server
    rp = srv.opts.recursion_po2
    words = srv.seal_words(srv._rec_circuit(KIND_LIFT))
    lifts = []
    for i in range(2):
        buf = _dev_buf(torch, words)
        srv.submit_prove_lift(0, Segment(index=310 + i, po2=10), d_out=buf.data_ptr(), host_seals=False)
        _, l = srv.wait_task(0)
        l.owner = buf
        assert l.seal is None
        lifts.append(l)
    out = _dev_buf(torch, words)
    srv.submit_recursion_dev(1, KIND_JOIN, lifts[0], lifts[1], d_out=out.data_ptr())
    j = srv.wait_task(1)
    ls = [b.owner.cpu().numpy().view(np.uint32) for b in lifts]
    d = oracle.hash_pair(oracle.seal_digest(ls[0]), oracle.seal_digest(ls[1]))
    ref = oracle.prove(rp, int(d[0]) | (int(d[1]) << 32), *RECURSION_WIDTHS, kind=KIND_JOIN, input_digest=d)
    assert np.array_equal(j.seal, ref) and np.array_equal(out.cpu().numpy().view(np.uint32), ref)
    assert j.claim == (310, 311) and j.kind == KIND_JOIN
    # device-resident verify_integrity of a device receipt
    srv.verify_integrity(DeviceReceipt(out.data_ptr(), words, KIND_JOIN, j.claim, [], out, None))
    # right input corrupted on the device -> the task fails in its own verification of that input, and names it
    lifts[1].owner[words // 2] ^= 1
    srv.submit_recursion_dev(1, KIND_JOIN, lifts[0], lifts[1], d_out=out.data_ptr())
    with pytest.raises(VerificationError) as ei:
        srv.wait_task(1)
    assert ei.value.what == "right receipt" and ei.value.code != 0
    lifts[1].owner[words // 2] ^= 1
    # a receipt whose metadata claims another kind than its seal was proved with is rejected (header bound to the expected circuit)
    wrong = DeviceReceipt(lifts[0].ptr, words, KIND_JOIN, lifts[0].claim, [], lifts[0].owner, None)
    srv.submit_recursion_dev(1, KIND_JOIN, wrong, lifts[1], d_out=out.data_ptr())
    with pytest.raises(VerificationError) as ei:
        srv.wait_task(1)
    assert ei.value.what == "left receipt" and ei.value.code == 103


def test_verify_integrity_binds_the_expected_circuit(small_server, oracle):
    """a valid seal of ANOTHER circuit / kind must not pass as the receipt it is presented as."""
    from boundless_b200 import Segment, VerifierContext
    from boundless_b200.prover_server import KIND_JOIN, KIND_LIFT, SegmentReceipt, SuccinctReceipt, VerificationError
    srv = small_server
    seg = srv.prove_segment(VerifierContext(), Segment(index=320, po2=10))
    lifted = srv.lift(seg)
    srv.verify_integrity(seg); srv.verify_integrity(lifted)
    with pytest.raises(VerificationError) as ei:          # a lift seal presented as a join receipt
        srv.verify_integrity(SuccinctReceipt(lifted.seal, KIND_JOIN, lifted.claim))
    assert ei.value.code == 103
    with pytest.raises(VerificationError) as ei:          # a po2-10 segment seal presented as a po2-11 one: wrong length
        srv.verify_integrity(SegmentReceipt(seg.seal, seg.index, 11))
    assert ei.value.code == 102
    small = srv.prove_segment(VerifierContext(), Segment(index=321, po2=9))
    with pytest.raises(VerificationError):                # a (valid) segment seal presented as a lift receipt
        srv.verify_integrity(SuccinctReceipt(small.seal, KIND_LIFT, (321, 321)))
    bad = seg.seal.copy(); bad[5] = 1                     # spare header words are bound too
    with pytest.raises(VerificationError) as ei:
        srv.verify_integrity(SegmentReceipt(bad, seg.index, seg.po2))
    assert ei.value.code == 103


def test_async_job_runner_with_real_proofs(gpu, small_server, oracle):
    """dist.JobRunner on one GPU: 5 segments, both slots busy, joins launched as soon as their inputs exist, receipts never leave the
    device; the root equals the oracle's evaluation of the same Planner DAG."""
    from boundless_b200 import Segment
    from boundless_b200.dist import B200Engine, JobRunner
    from boundless_b200.prover_server import KIND_JOIN, KIND_LIFT, RECURSION_WIDTHS
    torch, srv = gpu, small_server
    rp = srv.opts.recursion_po2
    eng = B200Engine(srv, lambda i: Segment(index=i, po2=9), torch.device("cuda", 0))
    root, stats = JobRunner(eng, 5).run()
    assert stats["proved"] == 5 and stats["joined"] == 4 and stats["sent"] == 0 and stats["max_in_flight"] == 2
    assert root.claim == (0, 4)
    def rec(kind, digest):
        return oracle.prove(rp, int(digest[0]) | (int(digest[1]) << 32), *RECURSION_WIDTHS, kind=kind, input_digest=digest)
    l = [rec(KIND_LIFT, oracle.seal_digest(oracle.prove(9, 0xB2000000 + i))) for i in range(5)]
    def jn(a, b):
        return rec(KIND_JOIN, oracle.hash_pair(oracle.seal_digest(a), oracle.seal_digest(b)))
    want = jn(jn(jn(l[0], l[1]), jn(l[2], l[3])), l[4])          # Planner: peaks (0..3) and 4, finish joins them
    got = eng.to_host(root)
    assert np.array_equal(got, want)
    assert oracle.verify(got) == 0


def test_povw_kinds_bit_exact(small_server, oracle):
    """lift_povw / join_povw / unwrap_povw (tasks/prove.rs:70-78, join_povw.rs:55, resolve_povw.rs:57): recursion proofs of kinds 5, 6, 7;
    each seal equals the oracle's for the same kind and input digest, verifies under its own kind, and is rejected under the plain one."""
    from boundless_b200 import B200Error, Segment, VerifierContext
    from boundless_b200.prover_server import (KIND_JOIN, KIND_JOIN_POVW, KIND_LIFT, KIND_LIFT_POVW, KIND_UNWRAP_POVW, RECURSION_WIDTHS,
                                              SuccinctReceipt, VerificationError)
    srv, ctx = small_server, VerifierContext()
    rp = srv.opts.recursion_po2
    def rec(kind, digest):
        return oracle.prove(rp, int(digest[0]) | (int(digest[1]) << 32), *RECURSION_WIDTHS, kind=kind, input_digest=digest)
    segs = [srv.prove_segment(ctx, Segment(index=500 + i, po2=10)) for i in range(2)]
    lifts = [srv.lift_povw(s) for s in segs]
    for s, l in zip(segs, lifts):
        assert l.kind == KIND_LIFT_POVW and np.array_equal(l.seal, rec(KIND_LIFT_POVW, oracle.seal_digest(s.seal)))
        srv.verify_integrity(l)
        assert oracle.verify(l.seal) == 0
    j = srv.join_povw(lifts[0], lifts[1])
    want = rec(KIND_JOIN_POVW, oracle.hash_pair(oracle.seal_digest(lifts[0].seal), oracle.seal_digest(lifts[1].seal)))
    assert j.kind == KIND_JOIN_POVW and j.claim == (500, 501) and np.array_equal(j.seal, want)
    srv.verify_integrity(j)
    u = srv.unwrap_povw(j)
    assert u.kind == KIND_UNWRAP_POVW and u.claim == j.claim and np.array_equal(u.seal, rec(KIND_UNWRAP_POVW, oracle.seal_digest(j.seal)))
    srv.verify_integrity(u)
    # a PoVW receipt does not pass as a plain one, and the plain programs refuse to stand in for the PoVW ones
    with pytest.raises(VerificationError):
        srv.verify_integrity(SuccinctReceipt(j.seal, KIND_JOIN, j.claim))
    plain = srv.lift(segs[0])
    assert plain.kind == KIND_LIFT and not np.array_equal(plain.seal, lifts[0].seal)
    with pytest.raises(B200Error):
        srv.join_povw(plain, lifts[1])
    with pytest.raises(B200Error):
        srv.unwrap_povw(plain)


def test_prove_keccak_and_union_bit_exact(small_server, oracle):
    """prove_keccak (tasks/keccak.rs:71-75): a kind-8 recursion-shaped proof of the requested size over the digest of the keccak input
    states; two such receipts reduce with `union` (tasks/union.rs:43-47).  All three seals equal the oracle's."""
    from boundless_b200 import B200Error
    from boundless_b200.prover_server import KIND_KECCAK, KIND_UNION, RECURSION_WIDTHS
    srv = small_server
    rp = srv.opts.recursion_po2
    def rec(po2, kind, digest):
        return oracle.prove(po2, int(digest[0]) | (int(digest[1]) << 32), *RECURSION_WIDTHS, kind=kind, input_digest=digest)
    rng = np.random.default_rng(8)
    recs = []
    for i, po2 in enumerate((10, 11)):
        states = rng.integers(0, 256, 200 * (3 + i), dtype=np.uint8).tobytes()
        r = srv.prove_keccak("%064x" % i, po2, "ee" * 32, states)
        words = np.frombuffer(states, dtype="<u2").astype(np.uint32)
        assert r.kind == KIND_KECCAK and np.array_equal(r.seal, rec(po2, KIND_KECCAK, oracle.seal_digest(words)))
        srv.verify_integrity(r)
        assert oracle.verify(r.seal) == 0
        recs.append(r)
    u = srv.union(recs[0], recs[1])
    want = rec(rp, KIND_UNION, oracle.hash_pair(oracle.seal_digest(recs[0].seal), oracle.seal_digest(recs[1].seal)))
    assert np.array_equal(u.seal, want)
    srv.verify_integrity(u)
    with pytest.raises(B200Error):
        srv.prove_keccak("00" * 32, 10, "ee" * 32, b"")
    with pytest.raises(B200Error):
        srv.prove_keccak("00" * 32, 10, "ee" * 32, bytes(199))
