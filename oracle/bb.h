/*
 * oracle/bb.h -- BabyBear field + degree-4 extension, CPU restatement.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED at seal level: the arithmetic of the reference's hot path lives in un-vendored
 * crates (risc0-zkp 3.0.3, risc0-sys 1.5.0, sppark 0.1.14; /root/reference/Cargo.lock:9155,9131,10315)
 * that are absent from /root/reference and cannot be built here (no cargo).  This file restates the
 * published algorithm (SURVEY.md Appendix A "Field") and is pinned by the field constants and the
 * Poseidon2 known-answer vector, not by reference-produced seals.
 *
 * Reference call sites that reach this arithmetic: prover/crates/workflow/src/tasks/prove.rs:44-52
 * (prove_segment), :96-104 (lift), tasks/join.rs:52-56 (join).
 */
#ifndef ORACLE_BB_H
#define ORACLE_BB_H
#include <stdint.h>
#include <stddef.h>

#define BB_P 2013265921u        /* 15 * 2^27 + 1 */
#define BB_M 0x88000001u        /* p^-1 mod 2^32 (risc0-zkp field/baby_bear.rs `M`) */
#define BB_R2 1172168163u       /* 2^64 mod p */
#define BB_BETA 11u             /* extension: X^4 + 11 (risc0 BETA = 11, NBETA = p - 11) */

typedef uint32_t fp;            /* Montgomery form: a * 2^32 mod p, always canonical in [0,p) */

/* risc0-zkp baby_bear.rs `mul`: o = a*b; low = -o; red = M*low; o += red*P; r = o>>32; r>=P ? r-P : r */
static inline fp fp_mul(fp a, fp b) {
    uint64_t o = (uint64_t)a * b;
    uint32_t low = 0u - (uint32_t)o;
    uint32_t red = BB_M * low;
    o += (uint64_t)red * BB_P;
    uint32_t r = (uint32_t)(o >> 32);
    return r >= BB_P ? r - BB_P : r;
}
static inline fp fp_add(fp a, fp b) { uint32_t r = a + b; return r >= BB_P ? r - BB_P : r; }
static inline fp fp_sub(fp a, fp b) { uint32_t r = a - b; return r > BB_P ? r + BB_P : r; }
static inline fp fp_neg(fp a) { return a ? BB_P - a : 0; }
static inline fp fp_from_u32(uint32_t x) { return fp_mul(x % BB_P, BB_R2); }   /* Elem::new */
static inline uint32_t fp_to_u32(fp a) { return fp_mul(a, 1); }                 /* as_u32 */
static inline fp fp_pow(fp a, uint64_t e) {
    fp r = fp_from_u32(1);
    while (e) { if (e & 1) r = fp_mul(r, a); a = fp_mul(a, a); e >>= 1; }
    return r;
}
static inline fp fp_inv(fp a) { return fp_pow(a, BB_P - 2); }

typedef struct { fp c[4]; } fp4;

static inline fp4 fp4_zero(void) { fp4 r = {{0, 0, 0, 0}}; return r; }
static inline fp4 fp4_from_fp(fp a) { fp4 r = {{a, 0, 0, 0}}; return r; }
static inline fp4 fp4_one(void) { return fp4_from_fp(fp_from_u32(1)); }
static inline fp4 fp4_add(fp4 a, fp4 b) { fp4 r; for (int i = 0; i < 4; i++) r.c[i] = fp_add(a.c[i], b.c[i]); return r; }
static inline fp4 fp4_sub(fp4 a, fp4 b) { fp4 r; for (int i = 0; i < 4; i++) r.c[i] = fp_sub(a.c[i], b.c[i]); return r; }
static inline fp4 fp4_mul_fp(fp4 a, fp b) { fp4 r; for (int i = 0; i < 4; i++) r.c[i] = fp_mul(a.c[i], b); return r; }
static inline int fp4_eq(fp4 a, fp4 b) { return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2] && a.c[3] == b.c[3]; }
/* ExtElem mul over X^4 = -11:
 * c0=a0b0+NB(a1b3+a2b2+a3b1) c1=a0b1+a1b0+NB(a2b3+a3b2) c2=a0b2+a1b1+a2b0+NB*a3b3 c3=a0b3+a1b2+a2b1+a3b0 */
static inline fp4 fp4_mul(fp4 a, fp4 b) {
    const fp nb = fp_from_u32(BB_P - BB_BETA);
    fp4 r;
    r.c[0] = fp_add(fp_mul(a.c[0], b.c[0]),
                    fp_mul(nb, fp_add(fp_add(fp_mul(a.c[1], b.c[3]), fp_mul(a.c[2], b.c[2])), fp_mul(a.c[3], b.c[1]))));
    r.c[1] = fp_add(fp_add(fp_mul(a.c[0], b.c[1]), fp_mul(a.c[1], b.c[0])),
                    fp_mul(nb, fp_add(fp_mul(a.c[2], b.c[3]), fp_mul(a.c[3], b.c[2]))));
    r.c[2] = fp_add(fp_add(fp_add(fp_mul(a.c[0], b.c[2]), fp_mul(a.c[1], b.c[1])), fp_mul(a.c[2], b.c[0])),
                    fp_mul(nb, fp_mul(a.c[3], b.c[3])));
    r.c[3] = fp_add(fp_add(fp_mul(a.c[0], b.c[3]), fp_mul(a.c[1], b.c[2])),
                    fp_add(fp_mul(a.c[2], b.c[1]), fp_mul(a.c[3], b.c[0])));
    return r;
}
static inline fp4 fp4_pow(fp4 a, uint64_t e) {
    fp4 r = fp4_one();
    while (e) { if (e & 1) r = fp4_mul(r, a); a = fp4_mul(a, a); e >>= 1; }
    return r;
}
/* a^-1 = a^(p^4-2): p^4 - 2 = (p^4 - 1) - 1; done as a^(p-2) * (a^p)^(p^3-ish) is overkill; use
 * the tower: a = A + X*B with A,B in Fp2=Fp[Y]/(Y^2+11), Y=X^2.  a^-1 = (A - X*B)/(A^2 - Y*B^2). */
static inline fp4 fp4_inv(fp4 a) {
    const fp nb = fp_from_u32(BB_P - BB_BETA);
    /* A = a0 + a2 Y, B = a1 + a3 Y; Y^2 = -11 */
    fp A0 = a.c[0], A1 = a.c[2], B0 = a.c[1], B1 = a.c[3];
    /* A^2 = (A0^2 + nb*A1^2) + 2 A0 A1 Y */
    fp A2_0 = fp_add(fp_mul(A0, A0), fp_mul(nb, fp_mul(A1, A1)));
    fp A2_1 = fp_mul(fp_add(A0, A0), A1);
    /* B^2 */
    fp B2_0 = fp_add(fp_mul(B0, B0), fp_mul(nb, fp_mul(B1, B1)));
    fp B2_1 = fp_mul(fp_add(B0, B0), B1);
    /* Y*B^2 = nb*B2_1 + B2_0 Y */
    fp D0 = fp_sub(A2_0, fp_mul(nb, B2_1));
    fp D1 = fp_sub(A2_1, B2_0);
    /* (D0 + D1 Y)^-1 = (D0 - D1 Y)/(D0^2 + 11 D1^2) */
    fp den = fp_sub(fp_mul(D0, D0), fp_mul(nb, fp_mul(D1, D1)));
    fp di = fp_inv(den);
    fp I0 = fp_mul(D0, di), I1 = fp_neg(fp_mul(D1, di));
    /* result = (A - X B) * (I0 + I1 Y): A*I = (A0 I0 + nb A1 I1) + (A0 I1 + A1 I0) Y */
    fp4 r;
    r.c[0] = fp_add(fp_mul(A0, I0), fp_mul(nb, fp_mul(A1, I1)));
    r.c[2] = fp_add(fp_mul(A0, I1), fp_mul(A1, I0));
    r.c[1] = fp_neg(fp_add(fp_mul(B0, I0), fp_mul(nb, fp_mul(B1, I1))));
    r.c[3] = fp_neg(fp_add(fp_mul(B0, I1), fp_mul(B1, I0)));
    return r;
}

static inline uint32_t bit_reverse(uint32_t x, unsigned bits) {
    uint32_t r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

#endif
