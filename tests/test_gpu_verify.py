"""The device-side verify_integrity (csrc/verify.cu) against the CPU oracle's independent verifier: both must accept every
honest seal (GPU-made or oracle-made, every circuit kind) and reject every tampered one with the same verdict code.

Mirrors the reference's "verify after every step" (tasks/prove.rs:56-58, :81-83, :106-108; tasks/join.rs:77-79)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 2013265921
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz"))


@pytest.fixture(scope="module")
def srv(gpu):
    from boundless_b200 import ProverOpts, get_prover_server
    s = get_prover_server(ProverOpts(segment_po2=14, recursion_po2=12, slots=2))
    yield s
    s.close()


def verdict(srv, seal, slot=0):
    from boundless_b200 import VerificationError
    try:
        srv.verify_seal(seal, slot)          # header-trusting form: the same question oracle.verify answers
        return 0
    except VerificationError as e:
        return e.code


@pytest.mark.parametrize("po2", [9, 10, 12, 13, 14])
def test_accepts_own_and_oracle_seals(srv, oracle, po2):
    from boundless_b200 import Segment, VerifierContext
    seg = Segment(index=40 + po2, po2=po2)
    r = srv.prove_segment(VerifierContext(), seg)
    srv.verify_integrity(r)
    assert verdict(srv, oracle.prove(po2, seg.seed)) == 0
    # the resident form: verification enqueued right behind the proof, no host round trip in between
    srv.submit_segment(1, seg)
    srv.submit_verify(1)
    r2 = srv.wait(1)
    assert int(srv._verdict[1].value) == 0 and np.array_equal(r2.seal, r.seal)


def test_accepts_recursion_kinds_and_golden(srv, oracle):
    from boundless_b200 import Segment, VerifierContext
    a = srv.lift(srv.prove_segment(VerifierContext(), Segment(index=1, po2=10)))
    b = srv.lift(srv.prove_segment(VerifierContext(), Segment(index=2, po2=10)))
    j = srv.join(a, b)
    for r in (a, b, j, srv.resolve(j, a), srv.union(a, b)):
        srv.verify_integrity(r)
        assert oracle.verify(r.seal) == 0
    assert verdict(srv, GOLD["seal_po2_9"]) == 0


def test_rejects_tampering_with_the_oracle_verdict(srv, oracle):
    """Single-word corruptions across every region of the seal: header, top layers, tap evaluations, FRI tops, final
    polynomial, query openings (leaf values, sibling digests, FRI leaves)."""
    seal = GOLD["seal_po2_9"]
    rng = np.random.default_rng(7)
    picks = list(range(0, 16)) + [int(v) for v in rng.integers(16, seal.size, 150)] + [seal.size - 1]
    codes = set()
    for k in picks:
        bad = seal.copy(); bad[k] = (int(bad[k]) + 1) % P
        want = oracle.verify(bad)
        got = verdict(srv, bad)
        assert want != 0 and got == want, (k, got, want)
        codes.add(got)
    assert len(codes) >= 5, codes                      # the sample really exercises different checks
    # non-canonical word, wrong length, garbage header
    bad = seal.copy(); bad[100] = P
    assert verdict(srv, bad) == oracle.verify(bad) == 106
    assert verdict(srv, seal[:-1]) == oracle.verify(seal[:-1]) == 102
    assert verdict(srv, np.concatenate([seal, seal[:1]])) == 102
    assert verdict(srv, seal[:8]) == oracle.verify(seal[:8]) == 100
    hdr = seal.copy(); hdr[1] = 3
    assert verdict(srv, hdr) == oracle.verify(hdr) == 101
    assert verdict(srv, seal) == 0                      # and the slot still works afterwards


def test_verify_errors_are_reported(srv, oracle):
    from boundless_b200 import B200Error, ProverOpts, get_prover_server
    small = get_prover_server(ProverOpts(segment_po2=9, recursion_po2=9, recursion_widths=(16, 32, 8), segment_widths=(16, 32, 8), slots=1))
    try:
        with pytest.raises(B200Error):
            small.submit_verify(0)                      # nothing proved on this slot yet
        with pytest.raises(B200Error):
            verdict(small, oracle.prove(9, 5))          # 16/208/32 columns exceed this prover's max circuit
        assert verdict(small, GOLD["seal_po2_9"]) == 0  # the golden seal (16/32/8 columns) fits
    finally:
        small.close()


def test_full_size_segment_and_lift_verify_on_device(gpu, oracle):
    """BASELINE config 2 size: the chain of tasks/prove.rs -- prove_segment, verify, lift, verify -- all on the device."""
    from boundless_b200 import ProverOpts, Segment, get_prover_server
    from boundless_b200.prover_server import KIND_LIFT
    s = get_prover_server(ProverOpts(segment_po2=20, recursion_po2=18, slots=1))
    try:
        s.submit_segment(0, Segment(index=5))
        s.submit_verify(0)
        r = s.wait(0)
        assert int(s._verdict[0].value) == 0
        s.submit_recursion(0, KIND_LIFT, r)
        s.submit_verify(0)
        l = s.wait(0)
        assert int(s._verdict[0].value) == 0
        bad = r.seal.copy(); bad[bad.size // 2] ^= 1
        assert verdict(s, bad) == oracle.verify(bad) != 0
    finally:
        s.close()
