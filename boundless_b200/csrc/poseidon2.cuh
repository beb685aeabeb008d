// Poseidon2 permutation over BabyBear, t=24 (rate 16, capacity 8), x^7, 8 full + 21 partial rounds.
//
// Device counterpart of risc0-sys' `sppark_poseidon2_rows` / `sppark_poseidon2_fold` inner permutation
// (un-vendored; SURVEY.md 2.1 kernel 2, Appendix A "Poseidon2"); reached from
// /root/reference/prover/crates/workflow/src/tasks/prove.rs:44-52 via ProverServer::prove_segment.
//
// One thread owns one 24-cell state in registers.  The kernel is INT32-pipe bound (1356 Montgomery
// multiplications per permutation), so the code is organised around instruction count:
//   * add/sub are IADD + VIADDMNMX (field.cuh);
//   * the internal layer s + d_i*c_i folds the "+ s" into the 64-bit IMAD.WIDE accumulator, so a cell
//     update costs one Montgomery multiply and no separate modular add;
//   * round constants / diagonal come from __constant__ (uniform across the warp -> c[bank][ofs] operands).
#pragma once
#include "field.cuh"
#include "constants.inc"

namespace b200 {

static __constant__ uint32_t c_rc[213] = B200_P2_RC_INIT;      // Montgomery form
static __constant__ uint32_t c_diag[24] = B200_P2_DIAG_INIT;   // Montgomery form
#ifndef B200_P2_SHOUP
#define B200_P2_SHOUP 1
#endif
#if B200_P2_SHOUP
static __constant__ uint32_t c_diag_plain[24] = B200_P2_DIAG_PLAIN_INIT;   // d (canonical integer)
static __constant__ uint32_t c_diag_shoup[24] = B200_P2_DIAG_SHOUP_INIT;   // floor(d * 2^32 / p)
#endif

// fp_add written so that ptxas cannot encode the sum as IMAD.IADD (multiplier pipe): both halves are VIADDMNMX (ALU pipe).
// Which adds of the linear layers use it is a measured trade-off (B200_P2_V, tools/microbench.cu).
#ifndef B200_P2_ZALL
#define B200_P2_ZALL 1
#endif
#ifndef B200_P2_V
#if B200_P2_ZALL
#define B200_P2_V 63
#else
#define B200_P2_V 0
#endif
#endif
__device__ __forceinline__ uint32_t fp_add_alu(uint32_t a, uint32_t b) {
    uint32_t s = addmin(a, b, 0xffffffffu);
    return addmin(s, 0u - P, s);
}
// B200_P2_Z: bitmask of add classes written as three-input adds (fp_add_z: IADD3 only, never IMAD.IADD):
// 1 M4 blocks, 2 column sums, 4 final adds of the external layer, 8 sum tree of the internal layer, 16 round-constant adds,
// 32 the "+ s" of the internal layer.
#ifndef B200_P2_Z
#define B200_P2_Z 0
#endif
#define P2_ADD_M4(a, b)  ((B200_P2_Z & 1) ? fp_add_z(a, b) : (B200_P2_V & 1) ? fp_add_alu(a, b) : fp_add(a, b))
#define P2_ADD_SUM(a, b) ((B200_P2_Z & 2) ? fp_add_z(a, b) : (B200_P2_V & 2) ? fp_add_alu(a, b) : fp_add(a, b))
#define P2_ADD_FIN(a, b) ((B200_P2_Z & 4) ? fp_add_z(a, b) : (B200_P2_V & 4) ? fp_add_alu(a, b) : fp_add(a, b))
#define P2_ADD_INT(a, b) ((B200_P2_Z & 8) ? fp_add_z(a, b) : (B200_P2_V & 8) ? fp_add_alu(a, b) : fp_add(a, b))
#define P2_ADD_RC(a, b)  ((B200_P2_Z & 16) ? fp_add_z(a, b) : (B200_P2_V & 16) ? fp_add_alu(a, b) : fp_add(a, b))
#define P2_ADD_IS(a, b)  ((B200_P2_Z & 32) ? fp_add_z(a, b) : (B200_P2_V & 32) ? fp_add_alu(a, b) : fp_add(a, b))
// B200_P2_ZALL: EVERY plain 32-bit add of the permutation is a VIADDMNMX (min(a + b, 2^32 - 1) == a + b), an ALU-pipe-only instruction,
// so that the multiplier pipe -- the unit that binds this kernel (ncu: fmaheavy 92 % busy, ALU 60 %) -- carries multiplies only.  Forcing
// a subset does nothing (ptxas re-balances by encoding OTHER adds as IMAD.IADD), three-input adds with a runtime zero get re-associated
// back into two-input ones, and with round 1's instruction mix forcing all adds overloaded the ALU pipe instead
// (profiles/microbench_zadd_r01.txt); with the fused reduction and the lazy internal rounds the ALU pipe has the room.
#if B200_P2_ZALL
__device__ __forceinline__ uint32_t p2_add32(uint32_t a, uint32_t b) { return addmin(a, b, 0xffffffffu); }
__device__ __forceinline__ uint32_t p2_fp_add(uint32_t a, uint32_t b) { return fp_add_alu(a, b); }
#else
__device__ __forceinline__ uint32_t p2_add32(uint32_t a, uint32_t b) { return a + b; }
__device__ __forceinline__ uint32_t p2_fp_add(uint32_t a, uint32_t b) { return fp_add(a, b); }
#endif

// B200_P2_LAZY: instruction-count variants prepared for measurement (arithmetic checked on the host by
// tests/test_device_code_on_host.py).  Bitmask:
//   1  S-box: x^4 is left uncorrected in (0, 2p) -- it only feeds one side of the last multiply          (-1 instruction per S-box)
//   2  internal layer: the 24-term sum is taken in 64 bits and reduced once                              (~ -10 per round)
//   4  internal rounds: cells 1..23 stay lazy in [0, 2p) between rounds, the sum is carried along        (~ -7 per round)
#ifndef B200_P2_LAZY
#define B200_P2_LAZY 7
#endif
__device__ __forceinline__ uint32_t p2_sbox(uint32_t x) {
    uint32_t x2 = fp_mul(x, x);
    uint32_t x3 = fp_mul(x2, x);
#if B200_P2_LAZY & 1
    uint32_t x4 = fp_mul_lazy(x2, x2);
#else
    uint32_t x4 = fp_mul(x2, x2);
#endif
    return fp_mul(x3, x4);
}
// V mod p for a 64-bit V < 2^38 (sums of up to ~100 field elements): q = floor((V >> 8) * floor(2^40 / p) / 2^32) is floor(V / p) or
// one less (the constant 546 is 0.99976 of 2^40 / p and V / p < 128, so the estimate falls short by less than 1), hence
// V - q*p is in [0, 2p), fits 32 bits, and one correction makes it canonical.  SHF + IMAD.HI + IMAD + VIADDMNMX.
__device__ __forceinline__ uint32_t fp_reduce38(uint64_t v) {
    const uint32_t q = __umulhi((uint32_t)(v >> 8), 546u);
    const uint32_t r = (uint32_t)v - q * P;
    return addmin(r, 0u - P, r);
}
// sum of N canonical values (+ a 64-bit starting value) without per-term corrections: two canonical values add without a carry
// (2p < 2^32), the pair sums are accumulated in 64 bits, the total is reduced once.  acc0 + N*p must stay below 2^38.
template <int N>
__device__ __forceinline__ uint32_t fp_sum64(const uint32_t (&v)[N], uint64_t acc0 = 0) {
    static_assert(N <= 64, "total must stay below 2^38");
    uint64_t acc = acc0;
#pragma unroll
    for (int i = 0; i + 3 < N; i += 4) acc += (uint64_t)p2_add32(v[i], v[i + 1]) + (uint64_t)p2_add32(v[i + 2], v[i + 3]);
    if ((N & 3) == 3) acc += (uint64_t)p2_add32(v[N - 3], v[N - 2]) + (uint64_t)v[N - 1];
    else if ((N & 3) == 2) acc += (uint64_t)p2_add32(v[N - 2], v[N - 1]);
    else if ((N & 3) == 1) acc += (uint64_t)v[N - 1];
    return fp_reduce38(acc);
}

// external linear layer: circ(2*M4, M4, ..., M4), M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]]
__device__ __forceinline__ void p2_m_ext(uint32_t (&c)[24]) {
#pragma unroll
    for (int k = 0; k < 6; k++) {
        uint32_t x0 = c[4 * k], x1 = c[4 * k + 1], x2 = c[4 * k + 2], x3 = c[4 * k + 3];
        uint32_t t0 = P2_ADD_M4(x0, x1), t1 = P2_ADD_M4(x2, x3);
        uint32_t t2 = P2_ADD_M4(P2_ADD_M4(x1, x1), t1), t3 = P2_ADD_M4(P2_ADD_M4(x3, x3), t0);
        uint32_t t1_2 = P2_ADD_M4(t1, t1), t0_2 = P2_ADD_M4(t0, t0);
        uint32_t t4 = P2_ADD_M4(P2_ADD_M4(t1_2, t1_2), t3), t5 = P2_ADD_M4(P2_ADD_M4(t0_2, t0_2), t2);
        c[4 * k] = P2_ADD_M4(t3, t5);
        c[4 * k + 1] = t5;
        c[4 * k + 2] = P2_ADD_M4(t2, t4);
        c[4 * k + 3] = t4;
    }
    uint32_t s0 = c[0], s1 = c[1], s2 = c[2], s3 = c[3];
#pragma unroll
    for (int k = 1; k < 6; k++) {
        s0 = P2_ADD_SUM(s0, c[4 * k]); s1 = P2_ADD_SUM(s1, c[4 * k + 1]);
        s2 = P2_ADD_SUM(s2, c[4 * k + 2]); s3 = P2_ADD_SUM(s3, c[4 * k + 3]);
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        c[4 * k] = P2_ADD_FIN(c[4 * k], s0); c[4 * k + 1] = P2_ADD_FIN(c[4 * k + 1], s1);
        c[4 * k + 2] = P2_ADD_FIN(c[4 * k + 2], s2); c[4 * k + 3] = P2_ADD_FIN(c[4 * k + 3], s3);
    }
}

// internal linear layer: c_i <- s + diag_i * c_i, s = sum c.
// "+ s" rides in the IMAD.WIDE accumulator: X == s*2^32 (mod p) with hi(X) < p/2, so d*c + X < p*2^32.
__device__ __forceinline__ void p2_m_int(uint32_t (&c)[24]) {
    uint32_t a0 = P2_ADD_INT(c[0], c[1]), a1 = P2_ADD_INT(c[2], c[3]), a2 = P2_ADD_INT(c[4], c[5]), a3 = P2_ADD_INT(c[6], c[7]);
    uint32_t a4 = P2_ADD_INT(c[8], c[9]), a5 = P2_ADD_INT(c[10], c[11]), a6 = P2_ADD_INT(c[12], c[13]), a7 = P2_ADD_INT(c[14], c[15]);
    uint32_t a8 = P2_ADD_INT(c[16], c[17]), a9 = P2_ADD_INT(c[18], c[19]), a10 = P2_ADD_INT(c[20], c[21]), a11 = P2_ADD_INT(c[22], c[23]);
    a0 = P2_ADD_INT(a0, a1); a2 = P2_ADD_INT(a2, a3); a4 = P2_ADD_INT(a4, a5); a6 = P2_ADD_INT(a6, a7); a8 = P2_ADD_INT(a8, a9); a10 = P2_ADD_INT(a10, a11);
    a0 = P2_ADD_INT(a0, a2); a4 = P2_ADD_INT(a4, a6); a8 = P2_ADD_INT(a8, a10);
    uint32_t s = P2_ADD_INT(P2_ADD_INT(a0, a4), a8);
#if B200_P2_LAZY & 2
    s = fp_sum64<24>(c);       // the add tree above is dead code then
#endif
#if B200_P2_SHOUP
    // Shoup multiplication by the constant d_i: q = hi(c * d'), r = c*d - q*p in [0, 2p); IMAD.HI + 2 IMAD instead of
    // IMAD.WIDE + IMAD + IMAD.HI (8 instead of 10 multiplier-pipe cycles)
#pragma unroll
    for (int i = 0; i < 24; i++) {
        const uint32_t q = __umulhi(c[i], c_diag_shoup[i]);
        uint32_t r = c[i] * c_diag_plain[i] - q * P;
        r = addmin(r, 0u - P, r);
        c[i] = P2_ADD_IS(r, s);
    }
    return;
#endif
    // X = s<<32 if s < (p+1)/2 else ((2s-p)<<31) = {hi: s-(p+1)/2, lo: 0x80000000}
    constexpr uint32_t HALF = (P + 1) / 2;
    bool big = s >= HALF;
    uint32_t xhi = big ? s - HALF : s;
    uint32_t xlo = big ? 0x80000000u : 0u;
    uint64_t X = ((uint64_t)xhi << 32) | xlo;
#pragma unroll
    for (int i = 0; i < 24; i++) c[i] = fp_mul_acc(c_diag[i], c[i], X);
}

// loop unrolling of the round loops (instruction-cache footprint against scheduling freedom across rounds; measured with
// tools/microbench.cu, profiles/microbench_unroll_r01.txt)
#ifndef B200_P2_UNROLL_EXT
#define B200_P2_UNROLL_EXT 1
#endif
#ifndef B200_P2_UNROLL_INT
#define B200_P2_UNROLL_INT 1
#endif
#define B200_PRAGMA(x) _Pragma(#x)
#define B200_UNROLL(n) B200_PRAGMA(unroll n)
// All 21 internal rounds with cells 1..23 kept lazy in [0, 2p) between rounds (B200_P2_LAZY & 4).  Invariant at the top of a round:
// c[0] canonical, c[1..23] in [0, 2p), T == sum_{i>=1} c[i] (mod p) canonical.  A Shoup multiply takes any u32, so the lazy cells feed
// it directly and the "+ s" needs no correction (one instruction less per cell and round); the next T is the sum of the CORRECTED
// Shoup outputs plus 23*s, taken in 64 bits (23*s is the IMAD.WIDE that starts the accumulator) and reduced once.
#if B200_P2_SHOUP
__device__ __forceinline__ void p2_internal_rounds_lazy(uint32_t (&c)[24]) {
    uint32_t T;
    {
        uint32_t v[23];
#pragma unroll
        for (int i = 0; i < 23; i++) v[i] = c[i + 1];
        T = fp_sum64<23>(v);
    }
#pragma unroll 1
    for (int r = 0; r < 21; r++) {
        const uint32_t y = p2_sbox(p2_fp_add(c[0], c_rc[96 + r]));
        const uint32_t s = p2_fp_add(y, T);                    // sum of the whole state after the S-box
        uint32_t rr[23];
        {
            const uint32_t q = __umulhi(y, c_diag_shoup[0]);
            uint32_t t = y * c_diag_plain[0] - q * P;
            t = addmin(t, 0u - P, t);
            c[0] = p2_fp_add(t, s);
        }
#pragma unroll
        for (int i = 1; i < 24; i++) {
            const uint32_t q = __umulhi(c[i], c_diag_shoup[i]);
            const uint32_t t = c[i] * c_diag_plain[i] - q * P;
            rr[i - 1] = addmin(t, 0u - P, t);                  // canonical
            c[i] = p2_add32(rr[i - 1], s);                     // lazy: < 2p < 2^32
        }
        T = fp_sum64<23>(rr, (uint64_t)s * 23u);               // < 46p
    }
#pragma unroll
    for (int i = 1; i < 24; i++) c[i] = addmin(c[i], 0u - P, c[i]);
}
#endif

// Hybrid form of the 21 internal rounds (B200_P2_NMACC = K in 1..24): cells 0..K-1 are updated as ONE Montgomery multiply-accumulate
// REDC(diag_i * c_i + X), X == s * 2^32 (mod p) with hi(X) < p/2 riding in the IMAD.WIDE accumulator (4 instructions, 10
// multiplier-pipe cycles, canonical result), cells K..23 as in p2_internal_rounds_lazy (Shoup multiply, 5 instructions, 8 cycles,
// lazy result).  K trades multiplier-pipe cycles for instruction count; the best K is a measurement (tools/microbench.cu).
// Invariant at the top of a round: c[0..K-1] canonical, c[K..23] in [0, 2p), T == sum_{i>=1} c[i] (mod p) canonical.
#ifndef B200_P2_NMACC
#define B200_P2_NMACC 0
#endif
#if B200_P2_NMACC > 0
__device__ __forceinline__ void p2_internal_rounds_hybrid(uint32_t (&c)[24]) {
    constexpr int K = B200_P2_NMACC;
    constexpr uint32_t HALF = (P + 1) / 2;
    uint32_t T;
    {
        uint32_t v[23];
#pragma unroll
        for (int i = 0; i < 23; i++) v[i] = c[i + 1];
        T = fp_sum64<23>(v);
    }
#pragma unroll 1
    for (int r = 0; r < 21; r++) {
        const uint32_t y = p2_sbox(p2_fp_add(c[0], c_rc[96 + r]));
        const uint32_t s = p2_fp_add(y, T);
        // X = s << 32 when s < (p+1)/2, else (2s - p) << 31 = {hi: s - (p+1)/2, lo: 2^31}: both are s * 2^32 mod p
        const uint32_t xhi = addmin(s, 0u - HALF, s);
        const uint32_t xlo = (s - xhi) << 31;                  // (p+1)/2 is odd
        const uint64_t X = ((uint64_t)xhi << 32) | xlo;
        uint32_t o[23];                                        // canonical summands of the next T
        c[0] = fp_mul_acc(c_diag[0], y, X);
#pragma unroll
        for (int i = 1; i < 24; i++) {
            if (i < K) {
                c[i] = fp_mul_acc(c_diag[i], c[i], X);
                o[i - 1] = c[i];
            } else {
                const uint32_t q = __umulhi(c[i], c_diag_shoup[i]);
                const uint32_t t = c[i] * c_diag_plain[i] - q * P;
                o[i - 1] = addmin(t, 0u - P, t);
                c[i] = p2_add32(o[i - 1], s);
            }
        }
        T = fp_sum64<23>(o, (uint64_t)s * (uint32_t)(24 - (K > 1 ? K : 1)));
    }
#pragma unroll
    for (int i = (K > 1 ? K : 1); i < 24; i++) c[i] = addmin(c[i], 0u - P, c[i]);
}
#endif

#ifndef B200_P2_MERGED
#define B200_P2_MERGED 0
#endif
__device__ __forceinline__ void p2_permute(uint32_t (&c)[24]) {
    p2_m_ext(c);
#if B200_P2_MERGED
    // one copy of the full-round body serves the first and the last four rounds (smaller instruction footprint)
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
        const int base = phase ? 117 : 0;
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 24; i++) c[i] = p2_sbox(P2_ADD_RC(c[i], c_rc[base + 24 * r + i]));
            p2_m_ext(c);
        }
        if (phase == 0) {
#pragma unroll 1
            for (int r = 0; r < 21; r++) {
                c[0] = p2_sbox(P2_ADD_RC(c[0], c_rc[96 + r]));
                p2_m_int(c);
            }
        }
    }
#else
B200_UNROLL(B200_P2_UNROLL_EXT)
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 24; i++) c[i] = p2_sbox(P2_ADD_RC(c[i], c_rc[24 * r + i]));
        p2_m_ext(c);
    }
#if B200_P2_NMACC > 0
    p2_internal_rounds_hybrid(c);
#elif (B200_P2_LAZY & 4) && B200_P2_SHOUP
    p2_internal_rounds_lazy(c);
#else
B200_UNROLL(B200_P2_UNROLL_INT)
    for (int r = 0; r < 21; r++) {
        c[0] = p2_sbox(P2_ADD_RC(c[0], c_rc[96 + r]));
        p2_m_int(c);
    }
#endif
B200_UNROLL(B200_P2_UNROLL_EXT)
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 24; i++) c[i] = p2_sbox(P2_ADD_RC(c[i], c_rc[117 + 24 * r + i]));
        p2_m_ext(c);
    }
#endif
}

// ---- warp-cooperative permutation (latency form) ----------------------------------------------------------------
// One WARP per state: lane i < 24 holds cell i, lanes 24..31 hold 0 and take part in the shuffles.  Used where a single
// permutation chain is on a proof's critical path (Fiat-Shamir transcript, hash of the tap evaluations, the top of a Merkle
// tree): 24 S-boxes / 24 diagonal multiplies run in parallel and the linear layers are shuffles, so one permutation takes
// ~4 us instead of ~23 us for the one-thread form.  Same field operations on canonical values => bit-identical results.
#ifndef B200_HOST_EMULATION      // warp shuffles have no host form; tests/host_emul covers the one-thread permutation
static __device__ const uint32_t g_rc_w[213] = B200_P2_RC_INIT;      // lane-indexed reads: global memory, not the constant bank
static __device__ const uint32_t g_diag_w[24] = B200_P2_DIAG_INIT;

#ifndef B200_P2W_REDUX
#define B200_P2W_REDUX 1
#endif
#ifndef B200_P2W_LAT
#define B200_P2W_LAT 1      // 1: the latency form of P2Warp below (round 2); 0: the first form (kept for A/B builds)
#endif
struct P2Warp {
    uint32_t lane, m4[4], diag;       // this lane's M4 row (Montgomery form of the small integers) and diagonal entry
    uint32_t rc[8];                   // this lane's round constant of each full round
    __device__ __forceinline__ void init() {
        lane = threadIdx.x & 31u;
        const uint32_t r = lane & 3u;
        // M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]]
        const uint32_t row = r == 0 ? 0x03010705u : r == 1 ? 0x01010604u : r == 2 ? 0x07050301u : 0x06040101u;
#pragma unroll
        for (int j = 0; j < 4; j++) m4[j] = fp_to_mont((row >> (8 * j)) & 0xffu);
        diag = lane < 24 ? g_diag_w[lane] : 0u;
#pragma unroll
        for (int k = 0; k < 8; k++) rc[k] = lane < 24 ? g_rc_w[(k < 4 ? 24 * k : 117 + 24 * (k - 4)) + lane] : 0u;
    }
#if B200_P2W_LAT == 0
    __device__ __forceinline__ uint32_t m_ext(uint32_t x) const {
        const uint32_t q = lane & ~3u;
        const uint32_t x0 = __shfl_sync(0xffffffffu, x, q), x1 = __shfl_sync(0xffffffffu, x, q + 1);
        const uint32_t x2 = __shfl_sync(0xffffffffu, x, q + 2), x3 = __shfl_sync(0xffffffffu, x, q + 3);
        const uint32_t y = fp_add(fp_add(fp_mul(m4[0], x0), fp_mul(m4[1], x1)), fp_add(fp_mul(m4[2], x2), fp_mul(m4[3], x3)));
        uint32_t s = y;                                   // column sums over the 8 quads (quads 6, 7 are zero)
        s = fp_add(s, __shfl_xor_sync(0xffffffffu, s, 4));
        s = fp_add(s, __shfl_xor_sync(0xffffffffu, s, 8));
        s = fp_add(s, __shfl_xor_sync(0xffffffffu, s, 16));
        return lane < 24 ? fp_add(y, s) : 0u;
    }
    __device__ __forceinline__ uint32_t permute(uint32_t x) const {
        x = m_ext(x);
#pragma unroll
        for (int r = 0; r < 4; r++) x = m_ext(p2_sbox(fp_add(x, rc[r])));      // lanes >= 24: sbox(0) = 0
#pragma unroll 1
        for (int r = 0; r < 21; r++) {
            // sum of cells 1..23, concurrently with the S-box of cell 0.  B200_P2W_REDUX (default): two warp-wide integer reductions
            // (REDUX.SUM, sm_80+) over the 16-bit halves -- each partial sum stays below 2^21 -- recombined in 64 bits and reduced once;
            // ~75 cycles on the critical path instead of ~160 for five shuffle + modular-add steps.
            uint32_t s = lane == 0 ? 0u : x;
            const uint32_t x0 = p2_sbox(fp_add(x, c_rc[96 + r]));                 // only lane 0's value is used
#if B200_P2W_REDUX
            {
                const uint32_t lo = __reduce_add_sync(0xffffffffu, s & 0xffffu), hi = __reduce_add_sync(0xffffffffu, s >> 16);
                s = fp_reduce38(((uint64_t)hi << 16) + lo);
            }
#else
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) s = fp_add(s, __shfl_xor_sync(0xffffffffu, s, d));
#endif
            const uint32_t c0 = __shfl_sync(0xffffffffu, x0, 0);
            s = fp_add(s, c0);
            if (lane == 0) x = c0;
            x = lane < 24 ? fp_add(fp_mul(diag, x), s) : 0u;
        }
#pragma unroll
        for (int r = 4; r < 8; r++) x = m_ext(p2_sbox(fp_add(x, rc[r])));
        return x;
    }
#else
    // Latency form (round 2): one permutation per warp is a dependent chain, so what counts is its length, not the instruction count.
    //  * external layer: the column sum over the six quads is five INDEPENDENT shuffles and a three-level add tree instead of three
    //    dependent shuffle + add steps;
    //  * internal round: new x_i = diag_i x_i + (sum of cells 1..23) + sbox(x_0).  diag_i x_i (lanes 1..23) and the REDUX sum run beside
    //    the S-box; lane 0 never waits for a shuffle, it takes (diag_0 + 1) sbox(x_0) + sum from its own registers; the other lanes add
    //    the broadcast S-box output last.  The next round constant is fetched a round ahead.
    __device__ __forceinline__ uint32_t m_ext(uint32_t x) const {
        const uint32_t q = lane & ~3u;
        const uint32_t x0 = __shfl_sync(0xffffffffu, x, q), x1 = __shfl_sync(0xffffffffu, x, q + 1);
        const uint32_t x2 = __shfl_sync(0xffffffffu, x, q + 2), x3 = __shfl_sync(0xffffffffu, x, q + 3);
        const uint32_t y = fp_add(fp_add(fp_mul(m4[0], x0), fp_mul(m4[1], x1)), fp_add(fp_mul(m4[2], x2), fp_mul(m4[3], x3)));
        // lanes 24..31 hold y = 0 (their x is 0), so a rotation by 4k over all 32 lanes visits the six quads and two zero quads
        const uint32_t y1 = __shfl_sync(0xffffffffu, y, lane + 4), y2 = __shfl_sync(0xffffffffu, y, lane + 8);
        const uint32_t y3 = __shfl_sync(0xffffffffu, y, lane + 12), y4 = __shfl_sync(0xffffffffu, y, lane + 16);
        const uint32_t y5 = __shfl_sync(0xffffffffu, y, lane + 20), y6 = __shfl_sync(0xffffffffu, y, lane + 24);
        const uint32_t y7 = __shfl_sync(0xffffffffu, y, lane + 28);
        const uint32_t s = fp_add(fp_add(fp_add(y, y), fp_add(y1, y2)), fp_add(fp_add(y3, y4), fp_add(fp_add(y5, y6), y7)));
        return lane < 24 ? s : 0u;
    }
    __device__ __forceinline__ uint32_t permute(uint32_t x) const {
        x = m_ext(x);
#pragma unroll
        for (int r = 0; r < 4; r++) x = m_ext(p2_sbox(fp_add(x, rc[r])));      // lanes >= 24: sbox(0) = 0
        const uint32_t dplus0 = fp_add(__shfl_sync(0xffffffffu, diag, 0), R1);    // diag_0 + 1 (Montgomery one)
        uint32_t rcr = c_rc[96];
#pragma unroll 1
        for (int r = 0; r < 21; r++) {
            const uint32_t rcn = c_rc[96 + (r < 20 ? r + 1 : r)];
            const uint32_t c0 = p2_sbox(fp_add(x, rcr));                           // lane 0's is the round's S-box output
            const uint32_t s = lane == 0 ? 0u : x;
            const uint32_t lo = __reduce_add_sync(0xffffffffu, s & 0xffffu), hi = __reduce_add_sync(0xffffffffu, s >> 16);
            const uint32_t srest = fp_reduce38(((uint64_t)hi << 16) + lo);         // cells 1..23, same value in every lane
            const uint32_t t = fp_add(fp_mul(diag, x), srest);                     // lanes 1..23: everything but the S-box output
            const uint32_t m0 = fp_add(fp_mul(dplus0, c0), srest);                 // lane 0: complete
            const uint32_t c0b = __shfl_sync(0xffffffffu, c0, 0);
            x = lane == 0 ? m0 : (lane < 24 ? fp_add(t, c0b) : 0u);
            rcr = rcn;
        }
#pragma unroll
        for (int r = 4; r < 8; r++) x = m_ext(p2_sbox(fp_add(x, rc[r])));
        return x;
    }
#endif
};

#endif  // B200_HOST_EMULATION

// Two independent states per thread, software-pipelined by half a round: while state A is in its (ALU-heavy) linear
// layer, state B is in its (multiplier-heavy) S-box layer, so both integer pipes stay busy inside one warp.
__device__ __forceinline__ void p2_sbox_full(uint32_t (&c)[24], int rc_base) {
#pragma unroll
    for (int i = 0; i < 24; i++) c[i] = p2_sbox(fp_add(c[i], c_rc[rc_base + i]));
}
__device__ __forceinline__ void p2_permute2(uint32_t (&a)[24], uint32_t (&b)[24]) {
    p2_m_ext(a); p2_m_ext(b);
    p2_sbox_full(a, 0);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
        p2_m_ext(a); p2_sbox_full(b, 24 * r);
        if (r < 3) p2_sbox_full(a, 24 * (r + 1));
        p2_m_ext(b);
    }
    a[0] = p2_sbox(fp_add(a[0], c_rc[96]));
#pragma unroll 1
    for (int r = 0; r < 21; r++) {
        p2_m_int(a); b[0] = p2_sbox(fp_add(b[0], c_rc[96 + r]));
        if (r < 20) a[0] = p2_sbox(fp_add(a[0], c_rc[97 + r]));
        p2_m_int(b);
    }
    p2_sbox_full(a, 117);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
        p2_m_ext(a); p2_sbox_full(b, 117 + 24 * r);
        if (r < 3) p2_sbox_full(a, 117 + 24 * (r + 1));
        p2_m_ext(b);
    }
}

}  // namespace b200
