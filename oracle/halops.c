/*
 * oracle/halops.c -- CPU restatement of the small HAL operations around the three inner loops (K8, K9 and the
 * element-wise helpers of SURVEY.md 8a/8b).  TEST INFRASTRUCTURE ONLY (see bb.h): the product never links this.
 *
 * PARITY UNPINNED: these follow the CPU `Hal` of risc0-zkp 3.0.3 (crates.io dependency, Cargo.lock of
 * /root/reference; sources absent) as restated from recall in SURVEY.md Appendix A / section 8b; the reference
 * tree holds no vectors for them.  Reached in the reference from prover/crates/workflow/src/tasks/prove.rs:44-52
 * (prove_segment) and :96-104 (lift) through risc0_zkp::prove::Prover::finalize.
 * Every loop below is written exactly as the scalar definition reads, so that it can be audited by eye.
 */
#include <string.h>
#include "oracle.h"

/* Hal::mix_poly_coeffs(output, mix_start, mix, input, combos, input_size, count) [RECALL-hi]:
 *   for idx < count { cur = mix_start; for i < input_size { output[combos[i]*count + idx] += cur * input[i*count + idx]; cur *= mix } }
 * output: ExtElem array (AoS) of n_combos*count, accumulated into. */
void oracle_mix_poly_coeffs(fp4 *out, fp4 mix_start, fp4 mix, const fp *in, const uint32_t *combos, size_t input_size,
                            size_t count) {
    for (size_t idx = 0; idx < count; idx++) {
        fp4 cur = mix_start;
        for (size_t i = 0; i < input_size; i++) {
            fp4 *o = &out[(size_t)combos[i] * count + idx];
            *o = fp4_add(*o, fp4_mul_fp(cur, in[i * count + idx]));
            cur = fp4_mul(cur, mix);
        }
    }
}

/* Hal::eltwise_sum_extelem(output: Elem[4*count], input: ExtElem[to_add*count]) [RECALL-hi]:
 *   tot = sum_i input[i*count + idx]; output[j*count + idx] = tot.elems[j]   (AoS in, 4 planes out) */
void oracle_eltwise_sum_extelem(fp *out, const fp4 *in, size_t count, size_t to_add) {
    for (size_t idx = 0; idx < count; idx++) {
        fp4 tot = fp4_zero();
        for (size_t i = 0; i < to_add; i++) tot = fp4_add(tot, in[i * count + idx]);
        for (int j = 0; j < 4; j++) out[(size_t)j * count + idx] = tot.c[j];
    }
}

/* Hal::eltwise_add_elem / eltwise_copy_elem / eltwise_zeroize_elem [RECALL-hi / -hi / -med]:
 * zeroize maps the INVALID marker 0xFFFFFFFF to 0 and leaves every other word alone. */
void oracle_eltwise_add_elem(fp *out, const fp *a, const fp *b, size_t count) {
    for (size_t i = 0; i < count; i++) out[i] = fp_add(a[i], b[i]);
}
void oracle_eltwise_copy_elem(fp *out, const fp *in, size_t count) { memcpy(out, in, count * sizeof(fp)); }
void oracle_eltwise_zeroize_elem(fp *io, size_t count) {
    for (size_t i = 0; i < count; i++) if (io[i] == 0xFFFFFFFFu) io[i] = 0;
}

/* poly_divide (risc0-zkp core/poly.rs; device twin supra_poly_divide) [RECALL-hi]: divide P(x) (natural coefficient order)
 * by (x - z) in place and return the remainder P(z):
 *   cur = 0; for i = size-1 .. 0 { next = z*cur + p[i]; p[i] = cur; cur = next }; return cur */
fp4 oracle_poly_divide(fp4 *p, size_t size, fp4 z) {
    fp4 cur = fp4_zero();
    for (size_t k = size; k-- > 0;) {
        fp4 next = fp4_add(fp4_mul(z, cur), p[k]);
        p[k] = cur;
        cur = next;
    }
    return cur;
}

/* Hal::prefix_products(io: ExtElem[count]) [RECALL-hi]: io[i] = io[0] * ... * io[i] (inclusive) */
void oracle_prefix_products(fp4 *io, size_t count) {
    for (size_t i = 1; i < count; i++) io[i] = fp4_mul(io[i - 1], io[i]);
}

/* Hal::gather_sample(dst, src, idx, size, stride) [RECALL-hi]: dst[g] = src[g*stride + idx], g < size */
void oracle_gather_sample(fp *dst, const fp *src, size_t idx, size_t size, size_t stride) {
    for (size_t g = 0; g < size; g++) dst[g] = src[g * stride + idx];
}

/* Hal::scatter(into, index, offsets, values) [RECALL-med]: row r owns entries index[r] .. index[r+1];
 *   into[offsets[k]] = values[k] for every k in that range (rows = n_index - 1). */
void oracle_scatter(fp *into, const uint32_t *index, size_t n_index, const uint32_t *offsets, const fp *values) {
    for (size_t r = 0; r + 1 < n_index; r++)
        for (uint32_t k = index[r]; k < index[r + 1]; k++) into[offsets[k]] = values[k];
}

/* MerkleTreeProver::prove(idx) (SURVEY Appendix A "Merkle") [RECALL-hi]: the `cols` leaf values of row idx, then the
 * sibling digests from the leaf layer up to (not including) the layer of top_size nodes.  Returns words written. */
size_t oracle_merkle_open(uint32_t *out, const fp *nodes, const fp *matrix, size_t rows, size_t cols, size_t top_size, size_t idx) {
    size_t w = 0;
    for (size_t c = 0; c < cols; c++) out[w++] = matrix[c * rows + idx];
    idx += rows;
    while (idx >= 2 * top_size) {
        memcpy(out + w, nodes + (idx ^ 1) * 8, 32);
        w += 8;
        idx >>= 1;
    }
    return w;
}

/* PolyGroup::new (SURVEY Appendix A "PolyGroup / commit_group") [RECALL-hi]: K1, K2, K3 then the Merkle tree over the
 * evaluations (rows = 4N, cols = count).  coeffs: in = count columns of 2^n evaluations, out = shifted bit-reversed
 * coefficients; evals: count x 2^(n+2); nodes: 2 * 2^(n+2) digests. */
void oracle_commit_group(fp *coeffs, fp *evals, fp *nodes, unsigned n, size_t count) {
    oracle_batch_interpolate_ntt(coeffs, n, count);
    oracle_batch_zk_shift(coeffs, n, count);
    oracle_batch_expand_into_evaluate_ntt(evals, coeffs, n, count, ORACLE_INV_RATE_LOG);
    oracle_merkle_build(nodes, evals, (size_t)1 << (n + ORACLE_INV_RATE_LOG), count);
}
