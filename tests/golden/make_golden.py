#!/usr/bin/env python3
"""Regenerate tests/golden/golden.npz from the CPU oracle.

The reference holds NO STARK golden vectors (SURVEY.md 0 finding 5), and its arithmetic (risc0-zkp 3.0.3) cannot be built or
imported here, so these fixtures pin (a) the published constants the oracle must reproduce -- the Poseidon2 known-answer vector
and constant prefixes recorded in SURVEY.md 8c / Appendix B, the BabyBear roots of unity of Appendix C -- and (b) the oracle's
own outputs on fixed seeds, so that any later change to the oracle (the parity anchor of every GPU test) is caught.
    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as o  # noqa: E402


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def main():
    rng = np.random.default_rng(20260925)
    g = {}
    # published pins
    g["p2_kat_in"] = np.arange(24, dtype=np.uint32)
    g["p2_kat_out"] = np.array([0x2ed3e23d, 0x12921fb0, 0x0e659e79, 0x61d81dc9, 0x32bae33b, 0x62486ae3, 0x1e681b60, 0x24b91325,
                                0x2a2ef5b9, 0x50e8593e, 0x5bc818ec, 0x10691997, 0x35a14520, 0x2ba6a3c5, 0x279d47ec, 0x55014e81,
                                0x5953a67f, 0x2f403111, 0x6b8828ff, 0x1801301f, 0x2749207a, 0x3dc9cf21, 0x3c985ba2, 0x57a99864], dtype=np.uint32)
    g["p2_rc_first8"] = np.array([0x0fa20c37, 0x0795bb97, 0x12c60b9c, 0x0eabd88e, 0x096485ca, 0x07093527, 0x1b1d4e50, 0x30a01ace], dtype=np.uint32)
    g["p2_diag"] = np.array([0x409133f0, 0x1667a8a1, 0x06a6c7b6, 0x6f53160e, 0x273b11d1, 0x03176c5d, 0x72f9bbf9, 0x73ceba91, 0x5cdef81d,
                             0x01393285, 0x46daee06, 0x065d7ba6, 0x52d72d6f, 0x05dd05e0, 0x3bab4b63, 0x6ada3842, 0x2fc5fbec, 0x770d61b0,
                             0x5715aae9, 0x03ef0e90, 0x75b6c770, 0x242adf5f, 0x00d0ca4c, 0x36c0e388], dtype=np.uint32)
    g["rou_fwd"] = np.array([1, 2013265920, 284861408, 1801542727, 567209306, 740045640, 918899846, 1881002012, 1453957774, 65325759,
                             1538055801, 515192888, 483885487, 157393079, 1695124103, 2005211659, 1540072241, 88064245, 1542985445,
                             1269900459, 1461624142, 825701067, 682402162, 1311873874, 1164520853, 352275361, 18769, 137], dtype=np.uint32)
    # oracle outputs on fixed inputs (regression pins)
    x = o.to_mont(rng.integers(0, o.P, 3 << 10, dtype=np.int64))
    g["ntt_in"] = x
    g["intt_out_sha"] = sha(o.batch_intt(x, 10, 3))
    g["expand_ntt_out_sha"] = sha(o.batch_expand_ntt(x, 10, 3, 2))
    g["zk_shift_out_sha"] = sha(o.batch_zk_shift(x, 10, 3))
    m = o.to_mont(rng.integers(0, o.P, 64 * 20, dtype=np.int64))
    g["merkle_in"] = m
    g["merkle_nodes"] = o.merkle_build(m, 64, 20)
    mix = o.to_mont(rng.integers(0, o.P, 4, dtype=np.int64))
    f = o.to_mont(rng.integers(0, o.P, 4 * 256, dtype=np.int64))
    g["fri_in"], g["fri_mix"], g["fri_out"] = f, mix, o.fri_fold(f, 256, mix)
    g["seal_po2_9"] = o.prove(9, 0xB2000000, 16, 32, 8)
    for po2 in (10, 12):
        g["seal_sha_po2_%d" % po2] = sha(o.prove(po2, 0xB2000000 + po2))
    d = o.seal_digest(g["seal_po2_9"])
    g["seal_digest_po2_9"] = d
    g["lift_sha"] = sha(o.prove(10, int(d[0]) | (int(d[1]) << 32), 16, 128, 16, kind=1, input_digest=d))
    # BASELINE-size seals: segments at po2 16 / 18 / 20 (seed 0xB2000000 + po2, 16/208/32), and lift + join at the
    # recursion size po2 18 (16/128/16) over two po2-12 segments (indices 40, 41).  Minutes of CPU; sha256 only.
    for po2 in (16, 18, 20):
        g["seal_sha_po2_%d" % po2] = sha(o.prove(po2, 0xB2000000 + po2))
    lifts = []
    for i in range(2):
        d = o.seal_digest(o.prove(12, 0xB2000000 + 40 + i))
        lifts.append(o.prove(18, int(d[0]) | (int(d[1]) << 32), 16, 128, 16, kind=1, input_digest=d))
        g["lift18_sha_%d" % i] = sha(lifts[i])
    d = o.hash_pair(o.seal_digest(lifts[0]), o.seal_digest(lifts[1]))
    g["join18_sha"] = sha(o.prove(18, int(d[0]) | (int(d[1]) << 32), 16, 128, 16, kind=2, input_digest=d))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
