/*
 * oracle/ntt.c -- radix-2 NTT / iNTT over BabyBear, CPU restatement.  TEST INFRASTRUCTURE ONLY (see bb.h).
 * PARITY UNPINNED at seal level; pinned by the roots-of-unity table and by the
 * O(n^2) DFT / interpolate-evaluate identities exercised in tests/test_oracle.py.
 *
 * Restates risc0-zkp 3.0.3 core/ntt.rs (un-vendored, Cargo.lock:9155) per SURVEY.md Appendix A "NTT":
 *   interpolate_ntt = rev_butterfly (DIF, natural in -> bit-reversed out) then scale by 1/2^n
 *   evaluate_ntt(io, expand_bits) = fwd_butterfly (DIT, bit-reversed in -> natural out) skipping the
 *                                   levels n <= expand_bits
 *   expand(out, in, e): out[i] = in[i >> e]
 *   zk_shift: coefficient of x^d *= 3^d, with slot j holding degree bitrev_n(j)
 * These are the CPU counterparts of risc0-sys' sppark_batch_iNTT / sppark_batch_NTT /
 * sppark_batch_expand / sppark_batch_zk_shift, reached from
 * prover/crates/workflow/src/tasks/prove.rs:44-52.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

static fp ROU_FWD[28], ROU_REV[28];
static fp *TW_FWD[28], *TW_REV[28];   /* TW[n][i] = ROU[n]^i, i < 2^(n-1) */
static int rou_ready = 0;

static void rou_init(void) {
    if (rou_ready) return;
#pragma omp critical(oracle_rou)
    {
        if (!rou_ready) {
            ROU_FWD[27] = fp_from_u32(137);
            for (int k = 26; k >= 0; k--) ROU_FWD[k] = fp_mul(ROU_FWD[k + 1], ROU_FWD[k + 1]);
            for (int k = 0; k < 28; k++) ROU_REV[k] = fp_inv(ROU_FWD[k]);
            rou_ready = 1;
        }
    }
}
fp oracle_rou_fwd(unsigned k) { rou_init(); return ROU_FWD[k]; }
fp oracle_rou_rev(unsigned k) { rou_init(); return ROU_REV[k]; }

static const fp *tw_table(int fwd, unsigned n) {
    rou_init();
    fp **T = fwd ? TW_FWD : TW_REV;
    if (!T[n]) {
#pragma omp critical(oracle_tw)
        {
            if (!T[n]) {
                size_t half = n ? ((size_t)1 << (n - 1)) : 1;
                fp *t = (fp *)malloc(half * sizeof(fp));
                fp step = fwd ? ROU_FWD[n] : ROU_REV[n], cur = fp_from_u32(1);
                for (size_t i = 0; i < half; i++) { t[i] = cur; cur = fp_mul(cur, step); }
                T[n] = t;
            }
        }
    }
    return T[n];
}
void oracle_ntt_prepare(unsigned max_n) { for (unsigned n = 1; n <= max_n; n++) { tw_table(0, n); tw_table(1, n); } }

static void rev_butterfly(fp *io, unsigned n) {
    if (n == 0) return;
    size_t half = (size_t)1 << (n - 1);
    const fp *tw = TW_REV[n];
    for (size_t i = 0; i < half; i++) {
        fp a = io[i], b = io[i + half];
        io[i] = fp_add(a, b);
        io[i + half] = fp_mul(fp_sub(a, b), tw[i]);
    }
    rev_butterfly(io, n - 1);
    rev_butterfly(io + half, n - 1);
}
static void fwd_butterfly(fp *io, unsigned n, unsigned expand_bits) {
    if (n == 0 || n == expand_bits) return;
    size_t half = (size_t)1 << (n - 1);
    fwd_butterfly(io, n - 1, expand_bits);
    fwd_butterfly(io + half, n - 1, expand_bits);
    const fp *tw = TW_FWD[n];
    for (size_t i = 0; i < half; i++) {
        fp a = io[i], b = fp_mul(io[i + half], tw[i]);
        io[i] = fp_add(a, b);
        io[i + half] = fp_sub(a, b);
    }
}

/* Same butterflies as rev_butterfly / fwd_butterfly, scheduled level-by-level so that ONE large transform can use every
 * host thread (the batched entry points use it when there are fewer polynomials than threads; results are identical). */
#define PAR_LEVELS 6
static void rev_butterfly_par(fp *io, unsigned n) {
    if (n <= PAR_LEVELS + 8) { rev_butterfly(io, n); return; }
    for (unsigned lev = n; lev > n - PAR_LEVELS; lev--) {
        const size_t half = (size_t)1 << (lev - 1), blocks = (size_t)1 << (n - lev);
        const fp *tw = TW_REV[lev];
#pragma omp parallel for schedule(static)
        for (long t = 0; t < (long)(blocks * half); t++) {
            size_t b = (size_t)t / half, i = (size_t)t % half;
            fp *p = io + (b << lev);
            fp a = p[i], c = p[i + half];
            p[i] = fp_add(a, c);
            p[i + half] = fp_mul(fp_sub(a, c), tw[i]);
        }
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (long b = 0; b < (long)((size_t)1 << PAR_LEVELS); b++) rev_butterfly(io + ((size_t)b << (n - PAR_LEVELS)), n - PAR_LEVELS);
}
static void fwd_butterfly_par(fp *io, unsigned n, unsigned expand_bits) {
    if (n <= PAR_LEVELS + 8 || n - PAR_LEVELS <= expand_bits) { fwd_butterfly(io, n, expand_bits); return; }
#pragma omp parallel for schedule(dynamic, 1)
    for (long b = 0; b < (long)((size_t)1 << PAR_LEVELS); b++) fwd_butterfly(io + ((size_t)b << (n - PAR_LEVELS)), n - PAR_LEVELS, expand_bits);
    for (unsigned lev = n - PAR_LEVELS + 1; lev <= n; lev++) {
        const size_t half = (size_t)1 << (lev - 1), blocks = (size_t)1 << (n - lev);
        const fp *tw = TW_FWD[lev];
#pragma omp parallel for schedule(static)
        for (long t = 0; t < (long)(blocks * half); t++) {
            size_t b = (size_t)t / half, i = (size_t)t % half;
            fp *p = io + (b << lev);
            fp a = p[i], c = fp_mul(p[i + half], tw[i]);
            p[i] = fp_add(a, c);
            p[i + half] = fp_sub(a, c);
        }
    }
}

void oracle_interpolate_ntt(fp *io, unsigned n) {
    oracle_ntt_prepare(n);
    rev_butterfly(io, n);
    fp norm = fp_inv(fp_from_u32(1u << n));
    size_t size = (size_t)1 << n;
    for (size_t i = 0; i < size; i++) io[i] = fp_mul(io[i], norm);
}
void oracle_evaluate_ntt(fp *io, unsigned n, unsigned expand_bits) {
    oracle_ntt_prepare(n);
    fwd_butterfly(io, n, expand_bits);
}
void oracle_expand(fp *out, const fp *in, unsigned n_in, unsigned e) {
    size_t size = (size_t)1 << (n_in + e);
    for (size_t i = 0; i < size; i++) out[i] = in[i >> e];
}
void oracle_zk_shift(fp *io, unsigned n) {
    size_t size = (size_t)1 << n;
    /* pw[d] = 3^d built by doubling in natural degree order, applied at slot bitrev(d) */
    fp three = fp_from_u32(3), cur = fp_from_u32(1);
    for (size_t d = 0; d < size; d++) {
        size_t j = bit_reverse((uint32_t)d, n);
        io[j] = fp_mul(io[j], cur);
        cur = fp_mul(cur, three);
    }
}
void oracle_bit_reverse(fp *io, unsigned n) {
    size_t size = (size_t)1 << n;
    for (size_t i = 0; i < size; i++) {
        size_t j = bit_reverse((uint32_t)i, n);
        if (i < j) { fp t = io[i]; io[i] = io[j]; io[j] = t; }
    }
}

/* batched forms (K1, K2, K3): `count` polynomials, column-major */
static int few_polys(size_t count) { return count < 16; }     /* fewer polynomials than a typical host has threads */

void oracle_batch_interpolate_ntt(fp *io, unsigned n, size_t count) {
    oracle_ntt_prepare(n);
    if (few_polys(count)) {
        fp norm = fp_inv(fp_from_u32(1u << n));
        for (size_t c = 0; c < count; c++) {
            fp *p = io + (c << n);
            rev_butterfly_par(p, n);
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)((size_t)1 << n); i++) p[i] = fp_mul(p[i], norm);
        }
        return;
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (long c = 0; c < (long)count; c++) oracle_interpolate_ntt(io + ((size_t)c << n), n);
}
void oracle_batch_zk_shift(fp *io, unsigned n, size_t count) {
#pragma omp parallel for schedule(dynamic, 1)
    for (long c = 0; c < (long)count; c++) oracle_zk_shift(io + ((size_t)c << n), n);
}
void oracle_batch_evaluate_ntt(fp *io, unsigned n, size_t count) {
    oracle_ntt_prepare(n);
#pragma omp parallel for schedule(dynamic, 1)
    for (long c = 0; c < (long)count; c++) oracle_evaluate_ntt(io + ((size_t)c << n), n, 0);
}
void oracle_batch_expand_into_evaluate_ntt(fp *out, const fp *in, unsigned n_in, size_t count, unsigned e) {
    oracle_ntt_prepare(n_in + e);
    if (few_polys(count)) {
        for (size_t c = 0; c < count; c++) {
            fp *o = out + (c << (n_in + e));
            const fp *src = in + (c << n_in);
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)((size_t)1 << (n_in + e)); i++) o[i] = src[(size_t)i >> e];
            fwd_butterfly_par(o, n_in + e, e);
        }
        return;
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (long c = 0; c < (long)count; c++) {
        fp *o = out + ((size_t)c << (n_in + e));
        oracle_expand(o, in + ((size_t)c << n_in), n_in, e);
        oracle_evaluate_ntt(o, n_in + e, e);
    }
}
