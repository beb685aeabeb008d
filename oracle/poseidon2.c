/*
 * oracle/poseidon2.c -- Poseidon2 over BabyBear (t=24, rate 16, out 8, x^7, R_F=8, R_P=21): CPU restatement.
 * TEST INFRASTRUCTURE ONLY (see bb.h header).  PARITY UNPINNED at seal level; this file is pinned by the
 * round-constant prefix, the internal diagonal and the full-permutation known-answer vector recorded in
 * SURVEY.md section 8c / Appendix B (all checked by oracle_selftest()).
 *
 * Restates risc0-zkp 3.0.3 core/hash/poseidon2 (un-vendored; Cargo.lock:9155) as described in
 * SURVEY.md Appendix A "Poseidon2": constants from the Horizen-Labs Grain LFSR, sponge that OVERWRITES
 * the rate cells, hash_pair, and the Poseidon2Rng used for Fiat-Shamir.
 */
#include "oracle.h"
#include <string.h>
#include <stdlib.h>

static fp RC[213];      /* Montgomery */
static fp DIAG[24];     /* Montgomery */
static uint32_t RC_CANON[213], DIAG_CANON[24];
static int consts_ready = 0;

/* ---- Grain LFSR ---- */
typedef struct { uint8_t b[80]; } grain_t;
static int grain_step(grain_t *g) {
    int nb = g->b[62] ^ g->b[51] ^ g->b[38] ^ g->b[23] ^ g->b[13] ^ g->b[0];
    memmove(g->b, g->b + 1, 79);
    g->b[79] = (uint8_t)nb;
    return nb;
}
static int grain_out(grain_t *g) {
    int nb = grain_step(g);
    while (nb == 0) { grain_step(g); nb = grain_step(g); }
    return grain_step(g);
}
static uint32_t grain_bits(grain_t *g, int k) {
    uint32_t v = 0;
    for (int i = 0; i < k; i++) v = (v << 1) | (uint32_t)grain_out(g);
    return v;
}
static void grain_put(grain_t *g, int *pos, uint32_t v, int k) {
    for (int i = k - 1; i >= 0; i--) g->b[(*pos)++] = (v >> i) & 1;
}

void oracle_p2_init(void) {
    if (consts_ready) return;
    grain_t g; int pos = 0;
    grain_put(&g, &pos, 1, 2); grain_put(&g, &pos, 0, 4); grain_put(&g, &pos, 31, 12);
    grain_put(&g, &pos, 24, 12); grain_put(&g, &pos, 8, 10); grain_put(&g, &pos, 21, 10);
    while (pos < 80) g.b[pos++] = 1;
    for (int i = 0; i < 160; i++) grain_step(&g);
    int n = 0;
    while (n < 213) { uint32_t x = grain_bits(&g, 31); if (x < BB_P) RC_CANON[n++] = x; }
    uint32_t batch[24];
    for (int b = 0; b < 5; b++) for (int i = 0; i < 24; i++) batch[i] = grain_bits(&g, 31);
    for (int i = 0; i < 24; i++) DIAG_CANON[i] = (uint32_t)(((uint64_t)batch[i] + BB_P - 1) % BB_P);
    for (int i = 0; i < 213; i++) RC[i] = fp_from_u32(RC_CANON[i]);
    for (int i = 0; i < 24; i++) DIAG[i] = fp_from_u32(DIAG_CANON[i]);
    consts_ready = 1;
}
const uint32_t *oracle_p2_rc_canon(void) { oracle_p2_init(); return RC_CANON; }
const uint32_t *oracle_p2_diag_canon(void) { oracle_p2_init(); return DIAG_CANON; }

static inline fp sbox(fp x) { fp x2 = fp_mul(x, x); fp x3 = fp_mul(x2, x); fp x4 = fp_mul(x2, x2); return fp_mul(x3, x4); }

static inline void m_ext(fp *c) {
    for (int k = 0; k < 6; k++) {
        fp *x = c + 4 * k;
        fp t0 = fp_add(x[0], x[1]), t1 = fp_add(x[2], x[3]);
        fp t2 = fp_add(fp_add(x[1], x[1]), t1), t3 = fp_add(fp_add(x[3], x[3]), t0);
        fp t1_4 = fp_add(t1, t1); t1_4 = fp_add(t1_4, t1_4);
        fp t0_4 = fp_add(t0, t0); t0_4 = fp_add(t0_4, t0_4);
        fp t4 = fp_add(t1_4, t3), t5 = fp_add(t0_4, t2);
        fp t6 = fp_add(t3, t5), t7 = fp_add(t2, t4);
        x[0] = t6; x[1] = t5; x[2] = t7; x[3] = t4;
    }
    fp s[4];
    for (int j = 0; j < 4; j++) {
        fp a = 0;
        for (int k = 0; k < 6; k++) a = fp_add(a, c[4 * k + j]);
        s[j] = a;
    }
    for (int i = 0; i < 24; i++) c[i] = fp_add(c[i], s[i & 3]);
}
static inline void m_int(fp *c) {
    fp s = 0;
    for (int i = 0; i < 24; i++) s = fp_add(s, c[i]);
    for (int i = 0; i < 24; i++) c[i] = fp_add(s, fp_mul(DIAG[i], c[i]));
}

/* poseidon2_mix: m_ext; 4 full; 21 partial; 4 full */
void oracle_p2_mix(fp *c) {
    m_ext(c);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 24; i++) c[i] = sbox(fp_add(c[i], RC[24 * r + i]));
        m_ext(c);
    }
    for (int r = 0; r < 21; r++) {
        c[0] = sbox(fp_add(c[0], RC[96 + r]));
        m_int(c);
    }
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 24; i++) c[i] = sbox(fp_add(c[i], RC[117 + 24 * r + i]));
        m_ext(c);
    }
}

/* hash_elem_slice over a strided view: elem k = src[k*stride].  Overwrite-mode sponge. */
void oracle_p2_hash_strided(fp *out8, const fp *src, size_t count, size_t stride) {
    fp st[24]; memset(st, 0, sizeof st);
    unsigned used = 0;
    for (size_t k = 0; k < count; k++) {
        st[used++] = src[k * stride];
        if (used == 16) { oracle_p2_mix(st); used = 0; }
    }
    if (used != 0 || count == 0) {
        for (unsigned i = used; i < 16; i++) st[i] = 0;
        oracle_p2_mix(st);
    }
    memcpy(out8, st, 8 * sizeof(fp));
}
void oracle_p2_hash_elems(fp *out8, const fp *src, size_t count) { oracle_p2_hash_strided(out8, src, count, 1); }

void oracle_p2_hash_pair(fp *out8, const fp *a8, const fp *b8) {
    fp st[24];
    memcpy(st, a8, 32); memcpy(st + 8, b8, 32); memset(st + 16, 0, 32);
    oracle_p2_mix(st);
    memcpy(out8, st, 32);
}

/* K4 hash_rows: leaf j = sponge over matrix[c*rows + j], c = 0..cols-1 (column-major) */
void oracle_p2_hash_rows(fp *out, const fp *matrix, size_t rows, size_t cols) {
    oracle_p2_init();
#pragma omp parallel for schedule(static)
    for (long j = 0; j < (long)rows; j++) oracle_p2_hash_strided(out + 8 * (size_t)j, matrix + j, cols, rows);
}

/* ---- Poseidon2Rng (Fiat-Shamir) ---- */
void oracle_rng_init(oracle_rng *r) { oracle_p2_init(); memset(r, 0, sizeof *r); }
void oracle_rng_mix(oracle_rng *r, const fp *digest8) {
    for (int i = 0; i < 8; i++) r->cells[i] = fp_add(r->cells[i], digest8[i]);
    oracle_p2_mix(r->cells);
    r->pool_used = 0;
}
fp oracle_rng_elem(oracle_rng *r) {
    if (r->pool_used == 16) { oracle_p2_mix(r->cells); r->pool_used = 0; }
    return r->cells[r->pool_used++];
}
uint32_t oracle_rng_bits(oracle_rng *r, unsigned bits) {
    uint32_t v = fp_to_u32(oracle_rng_elem(r));
    return bits >= 32 ? v : (v & ((1u << bits) - 1));
}
fp4 oracle_rng_ext(oracle_rng *r) {
    fp4 e;
    for (int i = 0; i < 4; i++) e.c[i] = oracle_rng_elem(r);
    return e;
}
