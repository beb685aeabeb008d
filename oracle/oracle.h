/*
 * oracle/oracle.h -- public surface of the CPU oracle (liboracle.so).
 * TEST INFRASTRUCTURE ONLY: see bb.h.  PARITY UNPINNED at seal level (no STARK golden vectors exist in
 * /root/reference; SURVEY.md section 0 finding 5, section 8c).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include "bb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Poseidon2 (poseidon2.c) ---- */
void oracle_p2_init(void);
const uint32_t *oracle_p2_rc_canon(void);      /* 213 canonical round constants */
const uint32_t *oracle_p2_diag_canon(void);    /* 24 canonical internal-diagonal entries */
void oracle_p2_mix(fp *cells24);
void oracle_p2_hash_elems(fp *out8, const fp *src, size_t count);
void oracle_p2_hash_strided(fp *out8, const fp *src, size_t count, size_t stride);
void oracle_p2_hash_pair(fp *out8, const fp *a8, const fp *b8);
void oracle_p2_hash_rows(fp *out, const fp *matrix, size_t rows, size_t cols);      /* K4 */

typedef struct { fp cells[24]; uint32_t pool_used; } oracle_rng;
void oracle_rng_init(oracle_rng *r);
void oracle_rng_mix(oracle_rng *r, const fp *digest8);
fp oracle_rng_elem(oracle_rng *r);
uint32_t oracle_rng_bits(oracle_rng *r, unsigned bits);
fp4 oracle_rng_ext(oracle_rng *r);

/* ---- NTT (ntt.c) ---- */
fp oracle_rou_fwd(unsigned k);
fp oracle_rou_rev(unsigned k);
void oracle_ntt_prepare(unsigned max_n);
void oracle_interpolate_ntt(fp *io, unsigned n);
void oracle_evaluate_ntt(fp *io, unsigned n, unsigned expand_bits);
void oracle_expand(fp *out, const fp *in, unsigned n_in, unsigned e);
void oracle_zk_shift(fp *io, unsigned n);
void oracle_bit_reverse(fp *io, unsigned n);
void oracle_batch_interpolate_ntt(fp *io, unsigned n, size_t count);                /* K1 */
void oracle_batch_zk_shift(fp *io, unsigned n, size_t count);                       /* K2 */
void oracle_batch_evaluate_ntt(fp *io, unsigned n, size_t count);
void oracle_batch_expand_into_evaluate_ntt(fp *out, const fp *in, unsigned n_in, size_t count, unsigned e); /* K3 */

/* ---- Merkle / FRI / DEEP / prover / verifier (stark.c) ---- */
#define ORACLE_QUERIES 50
#define ORACLE_INV_RATE_LOG 2
#define ORACLE_FRI_FOLD 16
#define ORACLE_FRI_MIN_DEGREE 256
#define ORACLE_CHECK_COLS 16
#define ORACLE_GLOBALS 16

typedef struct {
    uint32_t po2;       /* trace rows = 2^po2 (reference default 20: prover/crates/workflow/src/lib.rs:83-84) */
    uint32_t w_code, w_data, w_accum;   /* synthetic widths; segment default 16/208/32 */
    uint32_t kind;      /* 0 segment, 1 lift, 2 join, 3 resolve, 4 union (transcript domain separation) */
} oracle_circuit;

void oracle_merkle_build(fp *nodes /* 2*rows*8 */, const fp *matrix, size_t rows, size_t cols);   /* K4+K5 */
void oracle_fri_fold(fp *out, const fp *in, size_t in_size, fp4 mix);                             /* K6 */
void oracle_batch_evaluate_any(fp4 *out, const fp *coeffs, unsigned n, size_t count, fp4 x);      /* K7 */
/* ---- small HAL operations (halops.c): K8, K9 and element-wise helpers ---- */
void oracle_mix_poly_coeffs(fp4 *out, fp4 mix_start, fp4 mix, const fp *in, const uint32_t *combos, size_t input_size, size_t count);
void oracle_eltwise_sum_extelem(fp *out, const fp4 *in, size_t count, size_t to_add);
void oracle_eltwise_add_elem(fp *out, const fp *a, const fp *b, size_t count);
void oracle_eltwise_copy_elem(fp *out, const fp *in, size_t count);
void oracle_eltwise_zeroize_elem(fp *io, size_t count);
fp4 oracle_poly_divide(fp4 *p, size_t size, fp4 z);                                    /* returns the remainder P(z) */
void oracle_prefix_products(fp4 *io, size_t count);
void oracle_gather_sample(fp *dst, const fp *src, size_t idx, size_t size, size_t stride);
void oracle_scatter(fp *into, const uint32_t *index, size_t n_index, const uint32_t *offsets, const fp *values);
size_t oracle_merkle_open(uint32_t *out, const fp *nodes, const fp *matrix, size_t rows, size_t cols, size_t top_size, size_t idx);
void oracle_commit_group(fp *coeffs, fp *evals, fp *nodes, unsigned n, size_t count);
void oracle_gen_trace(fp *out, uint64_t seed, unsigned po2, size_t cols);   /* witgen stand-in (code+data) */
void oracle_segment_digest(fp *out8, uint64_t seed);
void oracle_seal_digest(fp *out8, const uint32_t *seal, size_t words);
size_t oracle_seal_words(const oracle_circuit *c);
/* trace may be NULL (then generated from seed); returns 0 on success */
int oracle_prove(const oracle_circuit *c, uint64_t seed, const fp *input_digest8, const fp *trace, uint32_t *seal);
/* returns 0 if the seal verifies, else a positive error code identifying the failed check */
int oracle_verify(const uint32_t *seal, size_t words);
int oracle_selftest(void);

#ifdef __cplusplus
}
#endif
#endif
