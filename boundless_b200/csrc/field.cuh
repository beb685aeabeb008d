// BabyBear (p = 15*2^27+1) Montgomery arithmetic and the degree-4 extension for sm_100a.
//
// Replaces the device field code of risc0-sys 1.5.0 / sppark 0.1.14 (un-vendored CUDA behind
// risc0_zkvm::ProverServer; reached from /root/reference/prover/crates/workflow/src/tasks/prove.rs:44-52).
// Representation is identical to the reference's: u32 Montgomery form a*2^32 mod p, canonical in [0,p),
// so buffers are interchangeable with a risc0 `DeviceBuffer<BabyBearElem>` (SURVEY.md 8b "Data layout").
//
// B200 pipe model (measured, profiles/microbench_r01.txt + ncu): every integer multiply (IMAD 2 cyc, IMAD.HI /
// IMAD.WIDE 4 cyc per warp per SMSP) runs on the single "fmaheavy" pipe, which is what saturates in the Poseidon2 and
// NTT kernels; plain adds that ptxas encodes as IMAD.IADD land on that pipe too.  So: a multiply is exactly
// IMAD.WIDE + IMAD + IMAD.HI (10 fmaheavy cycles) and every add/sub/correction is written with the DPX fused
// add-min (VIADDMNMX, ALU pipe), which ptxas cannot turn into an IMAD.
#pragma once
#include <stdint.h>
#ifndef B200_HOST_EMULATION      // tests/host_emul supplies the qualifiers and vector types itself
#include <cuda_runtime.h>
#endif

namespace b200 {

constexpr uint32_t P = 2013265921u;        // 0x78000001
constexpr uint32_t PINV = 0x88000001u;     // p^-1 mod 2^32
constexpr uint32_t R1 = 268435454u;        // 2^32 mod p  (Montgomery one)
constexpr uint32_t R2 = 1172168163u;       // 2^64 mod p
constexpr uint32_t NBETA_M = 1073741848u;  // mont(p - 11): X^4 = -11

// min(x + y, z) as one VIADDMNMX on sm_90+/sm_100
__device__ __forceinline__ uint32_t addmin(uint32_t x, uint32_t y, uint32_t z) { return __viaddmin_u32(x, y, z); }

// Build-time variants (tools/microbench.cu measures them; the defaults are the fastest measured on B200).
#ifndef B200_ADD_V
#define B200_ADD_V 0
#endif
#ifndef B200_REDC_V
#define B200_REDC_V 4      // the fused form below: 4 instructions per Montgomery multiply instead of 5 (round 2; variant 3 was round 1's)
#endif
__device__ __forceinline__ uint32_t fp_add(uint32_t a, uint32_t b) {
#if B200_ADD_V == 0
    uint32_t s = a + b;                     // < 2p < 2^32
#else
    uint32_t s = addmin(a, b, 0xffffffffu); // same sum, forced onto the ALU pipe
#endif
    return addmin(s, 0u - P, s);            // min(s - p, s): s - p wraps high when s < p
}
__device__ __forceinline__ uint32_t fp_sub(uint32_t a, uint32_t b) {
#if B200_ADD_V == 0
    uint32_t d = a - b;                     // wraps high when a < b
    return addmin(d, P, d);                 // min(d + p, d)
#else
    uint32_t d = addmin(a, P, 0xffffffffu) - b;   // a + p - b in (0, 2p)
    return addmin(d, 0u - P, d);            // min(d - p, d)
#endif
}
// A runtime zero the compiler cannot see through (%ctaid.z of a 1-deep grid; a uniform register, read once per thread): adding it
// makes a sum a THREE-input add, which only IADD3 (ALU pipe) can encode -- ptxas cannot turn it into IMAD.IADD (multiplier pipe).
// Used to steer chosen adds off the multiplier pipe in the multiplier-bound kernels (B200_P2_Z in poseidon2.cuh).
// INVARIANT: correct only in kernels launched with gridDim.z == 1, which every launch of this library is.  Since round 2
// NO default code path uses it -- the default reduction is B200_REDC_V == 4 and the default add placement is VIADDMNMX-based
// (B200_P2_ZALL) -- it survives only in the measured-and-rejected build variants (B200_REDC_V == 3, B200_P2_Z != 0).
#ifdef B200_HOST_EMULATION        // tests/host_emul compiles this header for the host: the runtime zero is a plain zero there
__device__ __forceinline__ uint32_t zreg() { return 0u; }
#else
__device__ __forceinline__ uint32_t zreg() { uint32_t z; asm("mov.u32 %0, %%ctaid.z;" : "=r"(z)); return z; }
#endif
__device__ __forceinline__ uint32_t fp_add_z(uint32_t a, uint32_t b) {
    uint32_t s = a + b + zreg();
    return addmin(s, 0u - P, s);
}
__device__ __forceinline__ uint32_t fp_sub_z(uint32_t a, uint32_t b) {
    uint32_t d = a - b + zreg();
    return addmin(d, P, d);
}
__device__ __forceinline__ uint32_t fp_neg(uint32_t a) { return a ? P - a : 0u; }
__device__ __forceinline__ uint32_t fp_dbl(uint32_t a) { return fp_add(a, a); }

// (T + m*p) >> 32 for T = hi:lo and m = -lo/p mod 2^32, i.e. the whole Montgomery reduction step, as ONE instruction: ptxas fuses the
// carry-chained pair mad.lo.cc / madc.hi into IMAD.HI.U32 Rd, m, p, T with the 64-bit product T as its addend (the discarded low
// word is 0 by construction).  Written in C, the same expression leaves a stray IADD3 behind; this form does not (checked in SASS:
// a Montgomery multiply is IMAD.WIDE + IMAD + IMAD.HI + VIADDMNMX).  Result in [0, T/2^32 + p).
__device__ __forceinline__ uint32_t fp_redc_step(uint32_t hi, uint32_t lo) {
    const uint32_t m = lo * (0u - PINV);
#ifdef B200_HOST_EMULATION
    return (uint32_t)(((uint64_t)m * P + (((uint64_t)hi << 32) | lo)) >> 32);
#else
    uint32_t r, zero;
    asm("{ mad.lo.cc.u32 %1, %2, %3, %4; madc.hi.u32 %0, %2, %3, %5; }" : "=r"(r), "=r"(zero) : "r"(m), "r"(P), "r"(lo), "r"(hi));
    return r;
#endif
}
// Montgomery reduction of T < p*2^32 given as (hi, lo): returns T / 2^32 mod p, canonical.
__device__ __forceinline__ uint32_t fp_redc(uint32_t hi, uint32_t lo) {
#if B200_REDC_V == 4
    const uint32_t r = fp_redc_step(hi, lo);     // [0, 2p)
    return addmin(r, 0u - P, r);
#elif B200_REDC_V == 0
    uint32_t m = lo * PINV;                 // m*p == lo (mod 2^32)
    uint32_t t = __umulhi(m, P);            // (T - m*p) / 2^32 = hi - t, in (-p, p)
    uint32_t r = hi - t;
    return addmin(r, P, r);                 // min(r + p, r)
#elif B200_REDC_V == 1
    uint32_t m = lo * (0u - PINV);          // m*p == -lo (mod 2^32)
    uint64_t o2 = (uint64_t)m * P + (((uint64_t)hi << 32) | lo);   // IMAD.HI(m, P, lo) + hi; low word cancels
    uint32_t r = (uint32_t)(o2 >> 32);      // (T + m*p) / 2^32 in [0, 2p)
    return addmin(r, 0u - P, r);            // min(r - p, r)
#elif B200_REDC_V == 3
    uint32_t m = lo * PINV;
    uint32_t t = __umulhi(m, P);
    uint32_t r = hi - t + zreg();           // three-input form: stays an IADD3 (ALU pipe)
    return addmin(r, P, r);
#else
    uint32_t m = lo + ((lo + (lo << 4)) << 27);   // lo * 0x88000001 with two LEAs (ALU pipe) instead of an IMAD
    uint32_t t = __umulhi(m, P);
    uint32_t r = hi - t;
    return addmin(r, P, r);
#endif
}
__device__ __forceinline__ uint32_t fp_mul(uint32_t a, uint32_t b) {
    uint64_t o = (uint64_t)a * b;
    return fp_redc((uint32_t)(o >> 32), (uint32_t)o);
}
// a*b + c*2^32-ish accumulate: returns (a*b + acc64) / 2^32 mod p; caller guarantees a*b + acc64 < p*2^32
__device__ __forceinline__ uint32_t fp_mul_acc(uint32_t a, uint32_t b, uint64_t acc64) {
    uint64_t o = (uint64_t)a * b + acc64;
    return fp_redc((uint32_t)(o >> 32), (uint32_t)o);
}
// a*b/2^32 mod p WITHOUT the final correction.  Same precondition as fp_mul (a*b < p*2^32); the result is in [0, a*b/2^32 + p), i.e.
// below 1.47p for canonical a, b and below 2p always.  A lazy value may feed ONE side of a following fp_mul / Shoup multiply (the
// other side canonical keeps the product below p*2^32), never an add.
__device__ __forceinline__ uint32_t fp_mul_lazy(uint32_t a, uint32_t b) {
    uint64_t o = (uint64_t)a * b;
#if B200_REDC_V == 4
    return fp_redc_step((uint32_t)(o >> 32), (uint32_t)o);
#else
    uint32_t m = (uint32_t)o * PINV;
    uint32_t t = __umulhi(m, P);
    return (uint32_t)(o >> 32) - t + P;      // (-p, p) + p, one three-input add
#endif
}
__device__ __forceinline__ uint32_t fp_sqr(uint32_t a) { return fp_mul(a, a); }
__device__ __forceinline__ uint32_t fp_to_mont(uint32_t x) { return fp_mul(x, R2); }      // x < p
__device__ __forceinline__ uint32_t fp_from_mont(uint32_t a) { return fp_mul(a, 1u); }
__device__ __forceinline__ uint32_t fp_pow(uint32_t a, uint64_t e) {
    uint32_t r = R1;
    while (e) { if (e & 1) r = fp_mul(r, a); a = fp_mul(a, a); e >>= 1; }
    return r;
}

struct Fp4 { uint32_t c[4]; };

__device__ __forceinline__ Fp4 ld_fp4(const uint32_t* p) {        // 16-byte aligned AoS element
    uint4 v = *reinterpret_cast<const uint4*>(p);
    return Fp4{{v.x, v.y, v.z, v.w}};
}
__device__ __forceinline__ void st_fp4(uint32_t* p, const Fp4& a) {
    *reinterpret_cast<uint4*>(p) = make_uint4(a.c[0], a.c[1], a.c[2], a.c[3]);
}
__device__ __forceinline__ Fp4 fp4_zero() { return Fp4{{0u, 0u, 0u, 0u}}; }
__device__ __forceinline__ Fp4 fp4_one() { return Fp4{{R1, 0u, 0u, 0u}}; }
__device__ __forceinline__ Fp4 fp4_add(const Fp4& a, const Fp4& b) {
    return Fp4{{fp_add(a.c[0], b.c[0]), fp_add(a.c[1], b.c[1]), fp_add(a.c[2], b.c[2]), fp_add(a.c[3], b.c[3])}};
}
__device__ __forceinline__ Fp4 fp4_sub(const Fp4& a, const Fp4& b) {
    return Fp4{{fp_sub(a.c[0], b.c[0]), fp_sub(a.c[1], b.c[1]), fp_sub(a.c[2], b.c[2]), fp_sub(a.c[3], b.c[3])}};
}
__device__ __forceinline__ Fp4 fp4_mul_fp(const Fp4& a, uint32_t b) {
    return Fp4{{fp_mul(a.c[0], b), fp_mul(a.c[1], b), fp_mul(a.c[2], b), fp_mul(a.c[3], b)}};
}
// acc += a * b (b in Fp)
__device__ __forceinline__ void fp4_fma_fp(Fp4& acc, const Fp4& a, uint32_t b) {
#pragma unroll
    for (int i = 0; i < 4; i++) acc.c[i] = fp_add(acc.c[i], fp_mul(a.c[i], b));
}
// ExtElem multiply over X^4 = -11 (same formula as risc0-zkp field/baby_bear.rs ExtElem::mul)
__device__ __forceinline__ Fp4 fp4_mul(const Fp4& a, const Fp4& b) {
    Fp4 r;
    r.c[0] = fp_add(fp_mul(a.c[0], b.c[0]),
                    fp_mul(NBETA_M, fp_add(fp_add(fp_mul(a.c[1], b.c[3]), fp_mul(a.c[2], b.c[2])), fp_mul(a.c[3], b.c[1]))));
    r.c[1] = fp_add(fp_add(fp_mul(a.c[0], b.c[1]), fp_mul(a.c[1], b.c[0])),
                    fp_mul(NBETA_M, fp_add(fp_mul(a.c[2], b.c[3]), fp_mul(a.c[3], b.c[2]))));
    r.c[2] = fp_add(fp_add(fp_add(fp_mul(a.c[0], b.c[2]), fp_mul(a.c[1], b.c[1])), fp_mul(a.c[2], b.c[0])),
                    fp_mul(NBETA_M, fp_mul(a.c[3], b.c[3])));
    r.c[3] = fp_add(fp_add(fp_mul(a.c[0], b.c[3]), fp_mul(a.c[1], b.c[2])),
                    fp_add(fp_mul(a.c[2], b.c[1]), fp_mul(a.c[3], b.c[0])));
    return r;
}
__device__ __forceinline__ Fp4 fp4_pow(Fp4 a, uint64_t e) {
    Fp4 r = fp4_one();
    while (e) { if (e & 1) r = fp4_mul(r, a); a = fp4_mul(a, a); e >>= 1; }
    return r;
}

__device__ __forceinline__ uint32_t bitrev(uint32_t x, uint32_t bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

}  // namespace b200
