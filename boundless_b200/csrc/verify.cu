// Seal verification on the device: the `verify_integrity_with_context` step the reference runs after every prove / lift / join
// (/root/reference/prover/crates/workflow/src/tasks/prove.rs:56-58, :81-83, :106-108; tasks/join.rs:41-46, :77-79;
// tasks/union.rs:51-53), for the synthetic protocol of DESIGN.md section 2.
//
// The transcript is replayed with the same device-resident Poseidon2Rng kernels the prover uses (no host round trip); the 50
// queries are checked by 50 CTAs in parallel: one warp per Merkle path (4 trace/check groups + the FRI rounds), then the DEEP
// quotient, the FRI fold chain and the final polynomial.  The result is a single word: 0 = valid, otherwise the code of the
// first failed check in transcript order (same numbering as the CPU oracle's verifier, which is an independent
// implementation used only by the tests).
#include "internal.h"
#include "poseidon2.cuh"

namespace b200 {

// ---- small device helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fp_inv_dev(uint32_t a) { return fp_pow(a, (uint64_t)P - 2); }
// inverse in Fp[X]/(X^4+11) through the quadratic tower: a = A + X B with A, B in Fp[Y]/(Y^2+11), Y = X^2;
// a * (A - X B) = A^2 - Y B^2 =: D in Fp[Y], and D * conj(D) = D0^2 + 11 D1^2 in Fp.
__device__ Fp4 fp4_inv_dev(const Fp4& a) {
    const uint32_t A0 = a.c[0], A1 = a.c[2], B0 = a.c[1], B1 = a.c[3];
    // squares in Fp[Y]: (s0 + s1 Y)^2 = (s0^2 - 11 s1^2) + 2 s0 s1 Y
    const uint32_t AA0 = fp_add(fp_sqr(A0), fp_mul(NBETA_M, fp_sqr(A1))), AA1 = fp_mul(fp_dbl(A0), A1);
    const uint32_t BB0 = fp_add(fp_sqr(B0), fp_mul(NBETA_M, fp_sqr(B1))), BB1 = fp_mul(fp_dbl(B0), B1);
    // Y * BB = -11 BB1 + BB0 Y
    const uint32_t D0 = fp_sub(AA0, fp_mul(NBETA_M, BB1)), D1 = fp_sub(AA1, BB0);
    const uint32_t norm = fp_sub(fp_sqr(D0), fp_mul(NBETA_M, fp_sqr(D1)));
    const uint32_t ni = fp_inv_dev(norm);
    const uint32_t I0 = fp_mul(D0, ni), I1 = fp_neg(fp_mul(D1, ni));          // D^-1 = I0 + I1 Y
    Fp4 r;
    r.c[0] = fp_add(fp_mul(A0, I0), fp_mul(NBETA_M, fp_mul(A1, I1)));
    r.c[2] = fp_add(fp_mul(A0, I1), fp_mul(A1, I0));
    r.c[1] = fp_neg(fp_add(fp_mul(B0, I0), fp_mul(NBETA_M, fp_mul(B1, I1))));
    r.c[3] = fp_neg(fp_add(fp_mul(B0, I1), fp_mul(B1, I0)));
    return r;
}
__device__ __forceinline__ bool fp4_eq(const Fp4& a, const Fp4& b) {
    return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2] && a.c[3] == b.c[3];
}
// multiply by X^e in Fp[X]/(X^4+11)
__device__ __forceinline__ Fp4 fp4_mul_xpow(const Fp4& a, int e) {
    Fp4 b = fp4_zero();
    b.c[e] = R1;
    return fp4_mul(a, b);
}
__device__ __forceinline__ Fp4 ld_fp4_u(const uint32_t* p) { return Fp4{{p[0], p[1], p[2], p[3]}}; }   // unaligned-safe

// block-wide sum of `n` Fp4 values per thread (n <= 4) through shared memory; result valid in thread 0
template <int NV>
__device__ void block_sum_fp4(Fp4 (&v)[NV], uint32_t* red /* blockDim.x * NV * 4 words */) {
    const uint32_t t = threadIdx.x;
#pragma unroll
    for (int k = 0; k < NV; k++)
#pragma unroll
        for (int e = 0; e < 4; e++) red[(t * NV + k) * 4 + e] = v[k].c[e];
    __syncthreads();
    for (uint32_t st = blockDim.x / 2; st >= 1; st >>= 1) {
        if (t < st)
            for (int w = 0; w < NV * 4; w++) red[t * NV * 4 + w] = fp_add(red[t * NV * 4 + w], red[(t + st) * NV * 4 + w]);
        __syncthreads();
    }
    if (t == 0)
#pragma unroll
        for (int k = 0; k < NV; k++)
#pragma unroll
            for (int e = 0; e < 4; e++) v[k].c[e] = red[k * 4 + e];
}

// ---- (1) every word of the seal must be a canonical field element ----------------------------------------------------------
__global__ void k_verify_canonical(uint32_t* __restrict__ ctx, const uint32_t* __restrict__ seal, uint32_t words) {
    bool bad = false;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) bad |= seal[i] >= P;
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicMax(&ctx[VCTX_RC], 106u);
}
// ---- (0) the header must name the circuit the caller expects (shape AND kind); the spare header words must be zero ----------------
__global__ void k_verify_header(uint32_t* __restrict__ ctx, const uint32_t* __restrict__ seal, uint32_t po2, uint32_t w_code,
                                uint32_t w_data, uint32_t w_accum, uint32_t kind) {
    const bool ok = seal[0] == po2 && seal[1] == w_code && seal[2] == w_data && seal[3] == w_accum && seal[4] == kind &&
                    seal[5] == 0 && seal[6] == 0 && seal[7] == 0;
    if (!ok) atomicMax(&ctx[VCTX_RC], 103u);
}
__global__ void k_verify_reset(uint32_t* ctx) {
    for (uint32_t i = threadIdx.x; i < VCTX_WORDS; i += blockDim.x) ctx[i] = 0;
}

// ---- (2) fold a top layer (<= 32 digests) to its root --------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_verify_fold_top(uint32_t* __restrict__ root_out, const uint32_t* __restrict__ top, uint32_t top_size) {
    __shared__ __align__(16) uint32_t nodes[64 * 8];
    for (uint32_t i = threadIdx.x; i < top_size * 8; i += blockDim.x) nodes[top_size * 8 + i] = top[i];
    __syncwarp();
    for (uint32_t sz = top_size / 2; sz >= 1; sz >>= 1) {
        if (threadIdx.x < sz) {
            const uint32_t i = sz + threadIdx.x;
            uint32_t st[24];
#pragma unroll
            for (int k = 0; k < 16; k++) st[k] = nodes[i * 16 + k];
#pragma unroll
            for (int k = 16; k < 24; k++) st[k] = 0;
            p2_permute(st);
#pragma unroll
            for (int k = 0; k < 8; k++) nodes[i * 8 + k] = st[k];
        }
        __syncwarp();
    }
    if (threadIdx.x < 8) root_out[threadIdx.x] = top_size == 1 ? top[threadIdx.x] : nodes[8 + threadIdx.x];
}

// ---- (3) constraint identity at the DEEP point -----------------------------------------------------------------------------
// sum_k poly_mix^k term_k(u) == sum_e X^e sum_q z^rev2(q) check[4e+q](z^4)
__global__ void __launch_bounds__(256) k_verify_constraint(uint32_t* __restrict__ ctx, const uint32_t* __restrict__ u,
                                                           const uint32_t* __restrict__ pm, const uint32_t* __restrict__ z_g, uint32_t w_code,
                                                           uint32_t w_data, uint32_t w_accum) {
    __shared__ uint32_t red[256 * 4];
    const uint32_t W = w_code + w_data + w_accum, nq = W / 4;
    Fp4 acc[1] = {fp4_zero()};
    for (uint32_t k = threadIdx.x; k < nq + w_accum; k += blockDim.x) {
        Fp4 t;
        if (k < nq) {
            t = fp4_mul(fp4_mul(ld_fp4_u(u + 16 * k), ld_fp4_u(u + 16 * k + 4)), fp4_mul(ld_fp4_u(u + 16 * k + 8), ld_fp4_u(u + 16 * k + 12)));
        } else {
            const uint32_t a = k - nq;
            t = fp4_mul(fp4_sub(ld_fp4_u(u + 4 * (w_code + w_data + a)), ld_fp4_u(u + 4 * (W + a))), ld_fp4_u(u + 4 * (a % w_code)));
        }
        acc[0] = fp4_add(acc[0], fp4_mul(ld_fp4_u(pm + 4 * k), t));
    }
    block_sum_fp4<1>(acc, red);
    if (threadIdx.x == 0) {
        const Fp4 z = ld_fp4_u(z_g);
        Fp4 zp[4];
        zp[0] = fp4_one();
        for (int k = 1; k < 4; k++) zp[k] = fp4_mul(zp[k - 1], z);
        const uint32_t* uchk = u + 4 * (W + w_accum);
        Fp4 rhs = fp4_zero();
        for (int e = 0; e < 4; e++) {
            Fp4 inner = fp4_zero();
            for (uint32_t q = 0; q < 4; q++) inner = fp4_add(inner, fp4_mul(zp[bitrev(q, 2)], ld_fp4_u(uchk + 4 * (4 * e + q))));
            rhs = fp4_add(rhs, fp4_mul_xpow(inner, e));
        }
        if (!fp4_eq(acc[0], rhs) && ctx[VCTX_RC] == 0) ctx[VCTX_RC] = 110;
    }
}

// ---- (4) usum[pt] = sum over the taps opened at point pt of mix^t u_t -------------------------------------------------------------
__global__ void __launch_bounds__(256) k_verify_usum(uint32_t* __restrict__ ctx, const uint32_t* __restrict__ u, const uint32_t* __restrict__ mp,
                                                     uint32_t W, uint32_t w_accum, uint32_t T) {
    __shared__ uint32_t red[256 * 12];
    Fp4 s[3] = {fp4_zero(), fp4_zero(), fp4_zero()};
    for (uint32_t t = threadIdx.x; t < T; t += blockDim.x) {
        const Fp4 v = fp4_mul(ld_fp4_u(mp + 4 * t), ld_fp4_u(u + 4 * t));
        const int pt = t < W ? 0 : (t < W + w_accum ? 1 : 2);
        s[pt] = fp4_add(s[pt], v);
    }
    block_sum_fp4<3>(s, red);
    if (threadIdx.x == 0)
        for (int pt = 0; pt < 3; pt++)
            for (int e = 0; e < 4; e++) ctx[VCTX_USUM + 4 * pt + e] = s[pt].c[e];
}

// ---- (5) the queries: one CTA per query --------------------------------------------------------------------------------------
// Merkle path of one opening, by one thread: leaf = sponge(values), then hash_pair up to the committed top layer
__device__ bool verify_path(const uint32_t* __restrict__ vals, uint32_t cols, uint32_t rows, uint32_t top_size, uint32_t idx,
                            const uint32_t* __restrict__ top) {
    uint32_t st[24];
#pragma unroll
    for (int i = 0; i < 24; i++) st[i] = 0;
    uint32_t k = 0;
    for (; k + 16 <= cols; k += 16) {
#pragma unroll
        for (int i = 0; i < 16; i++) st[i] = vals[k + i];
        p2_permute(st);
    }
    const uint32_t rem = cols - k;
    if (rem != 0 || cols == 0) {
#pragma unroll
        for (int i = 0; i < 16; i++) st[i] = (uint32_t)i < rem ? vals[k + i] : 0u;
        p2_permute(st);
    }
    const uint32_t* sib = vals + cols;
    uint32_t node = idx + rows;
    while (node >= 2 * top_size) {
        if (node & 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) { st[8 + i] = st[i]; st[i] = sib[i]; }
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) st[8 + i] = sib[i];
        }
#pragma unroll
        for (int i = 16; i < 24; i++) st[i] = 0;
        p2_permute(st);
        sib += 8;
        node >>= 1;
    }
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 8; i++) ok &= st[i] == top[(size_t)(node - top_size) * 8 + i];
    return ok;
}

__global__ void __launch_bounds__(256) k_verify_queries(uint32_t* __restrict__ ctx, const uint32_t* __restrict__ seal, const VerifyShape sh,
                                                        const uint32_t* __restrict__ mp, const uint32_t* __restrict__ pts_g,
                                                        const uint32_t* __restrict__ fmix_g, const uint32_t* __restrict__ pos_g) {
    __shared__ uint32_t red[256 * 12];
    __shared__ uint32_t path_ok[8];
    const uint32_t q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t D = 4u << sh.po2;
    const uint32_t pos = pos_g[q] & (D - 1);
    const uint32_t* rec = seal + sh.off_queries + (size_t)q * sh.query_words;
    const uint32_t widths[4] = {sh.w_code, sh.w_data, sh.w_accum, (uint32_t)CHECK_COLS};
    const uint32_t ntrees = 4 + sh.rounds;

    // (a) Merkle paths, one warp per tree (lane 0 walks the path: it is a dependent chain of permutations)
    for (uint32_t tr = warp; tr < ntrees; tr += blockDim.x / 32) {
        if (lane == 0) {
            bool ok;
            if (tr < 4) {
                ok = verify_path(rec + sh.q_off_group[tr], widths[tr], D, D < 32 ? D : 32, pos, seal + sh.off_top[tr]);
            } else {
                const uint32_t k = tr - 4, rows = sh.fri_rows[k];
                ok = verify_path(rec + sh.q_off_fri[k], 4 * FRI_FOLD, rows, sh.fri_top[k], pos & (rows - 1), seal + sh.off_fri_top[k]);
            }
            path_ok[tr] = ok ? 1u : 0u;
        }
    }
    // (b) DEEP numerators from the opened rows: num[pt] = sum_t mix^t leaf_t
    Fp4 num[3] = {fp4_zero(), fp4_zero(), fp4_zero()};
    const uint32_t wcd = sh.w_code + sh.w_data;
    for (uint32_t t = tid; t < sh.T; t += blockDim.x) {
        const Fp4 m = ld_fp4_u(mp + 4 * t);
        uint32_t v; int pt;
        if (t < sh.w_code) { v = rec[sh.q_off_group[0] + t]; pt = 0; }
        else if (t < wcd) { v = rec[sh.q_off_group[1] + (t - sh.w_code)]; pt = 0; }
        else if (t < sh.W) { v = rec[sh.q_off_group[2] + (t - wcd)]; pt = 0; }
        else if (t < sh.W + sh.w_accum) { v = rec[sh.q_off_group[2] + (t - sh.W)]; pt = 1; }
        else { v = rec[sh.q_off_group[3] + (t - sh.W - sh.w_accum)]; pt = 2; }
        fp4_fma_fp(num[pt], m, v);
    }
    block_sum_fp4<3>(num, red);
    __syncthreads();
    // (c) the final polynomial at the last folded point: val_e = sum_j final[e][j] x^bitrev(j)
    uint32_t last_rows = D, last_pos = pos;
    for (uint32_t k = 0; k < sh.rounds; k++) { last_rows = sh.fri_rows[k]; last_pos &= last_rows - 1; }
    uint32_t lg_last = 0;
    while ((1u << lg_last) < last_rows) lg_last++;
    const uint32_t xf = fp_pow(sh.rou_fwd[lg_last], last_pos);
    Fp4 fin[1] = {fp4_zero()};
    if (tid < sh.final_size) {
        const uint32_t xp = fp_pow(xf, bitrev(tid, sh.final_lg));
        const uint32_t* fc = seal + sh.off_final;
#pragma unroll
        for (int e = 0; e < 4; e++) fin[0].c[e] = fp_mul(fc[(size_t)e * sh.final_size + tid], xp);
    }
    block_sum_fp4<1>(fin, red);
    // (d) thread 0: DEEP quotient, fold chain, verdict (first failure in the order an in-order verifier meets them)
    if (tid == 0) {
        uint32_t rc = 0;
        for (uint32_t g = 0; g < 4 && !rc; g++) if (!path_ok[g]) rc = 120 + g;
        if (!rc) {
            const uint32_t x = fp_pow(sh.rou_fwd[sh.po2 + 2], pos);
            Fp4 expect = fp4_zero();
            for (int pt = 0; pt < 3; pt++) {
                Fp4 den = fp4_zero();
                den.c[0] = x;
                den = fp4_sub(den, ld_fp4_u(pts_g + 4 * pt));
                const Fp4 us = ld_fp4_u(ctx + VCTX_USUM + 4 * pt);
                expect = fp4_add(expect, fp4_mul(fp4_sub(num[pt], us), fp4_inv_dev(den)));
            }
            uint32_t p = pos, dom = D;
            for (uint32_t k = 0; k < sh.rounds && !rc; k++) {
                const uint32_t rows = sh.fri_rows[k];
                const uint32_t group = p & (rows - 1), quot = p / rows;
                if (!path_ok[4 + k]) { rc = 130 + k; break; }
                const uint32_t* lv = rec + sh.q_off_fri[k];
                Fp4 f[16];
                for (int kk = 0; kk < 16; kk++)
                    for (int e = 0; e < 4; e++) f[kk].c[e] = lv[e * 16 + kk];
                Fp4 at = f[0];
                for (int kk = 1; kk < 16; kk++) if ((uint32_t)kk == quot) at = f[kk];
                if (!fp4_eq(at, expect)) { rc = 140 + k; break; }
                // fold: the 16 values are f(x zeta^kk), zeta = w_16; P_i(x^16) = 1/16 sum_kk (x zeta^kk)^-i f(x zeta^kk)
                uint32_t lgdom = 0;
                while ((1u << lgdom) < dom) lgdom++;
                const uint32_t xg_inv = fp_inv_dev(fp_pow(sh.rou_fwd[lgdom], group));
                const uint32_t zeta_inv = fp_inv_dev(sh.rou_fwd[4]);
                uint32_t base[16], cur[16];
                uint32_t cz = R1;
                for (int kk = 0; kk < 16; kk++) { base[kk] = fp_mul(xg_inv, cz); cz = fp_mul(cz, zeta_inv); cur[kk] = R1; }
                const Fp4 fm = ld_fp4_u(fmix_g + 4 * k);
                Fp4 next = fp4_zero(), mpw = fp4_one();
                for (int i = 0; i < 16; i++) {
                    Fp4 Pi = fp4_zero();
                    for (int kk = 0; kk < 16; kk++) { fp4_fma_fp(Pi, f[kk], cur[kk]); cur[kk] = fp_mul(cur[kk], base[kk]); }
                    Pi = fp4_mul_fp(Pi, sh.inv16);
                    next = fp4_add(next, fp4_mul(mpw, Pi));
                    mpw = fp4_mul(mpw, fm);
                }
                expect = next; p = group; dom = rows;
            }
            if (!rc && !fp4_eq(fin[0], expect)) rc = 150;
        }
        ctx[VCTX_QRC + q] = rc;
    }
}

// ---- (6) verdict ------------------------------------------------------------------------------------------------------
__global__ void k_verify_finish(uint32_t* __restrict__ ctx) {
    uint32_t rc = ctx[VCTX_RC];
    for (uint32_t q = 0; q < (uint32_t)QUERIES && !rc; q++) rc = ctx[VCTX_QRC + q];
    ctx[VCTX_RESULT] = rc;
}

cudaError_t launch_verify_reset(uint32_t* ctx, cudaStream_t s) { B200_LAUNCH(k_verify_reset)<<<1, 128, 0, s>>>(ctx); return cudaGetLastError(); }
cudaError_t launch_verify_header(uint32_t* ctx, const uint32_t* seal, uint32_t po2, uint32_t w_code, uint32_t w_data, uint32_t w_accum,
                                 uint32_t kind, cudaStream_t s) {
    B200_LAUNCH(k_verify_header)<<<1, 1, 0, s>>>(ctx, seal, po2, w_code, w_data, w_accum, kind); return cudaGetLastError();
}
cudaError_t launch_verify_canonical(uint32_t* ctx, const uint32_t* seal, uint32_t words, cudaStream_t s) {
    uint32_t grid = (words + 255) / 256; if (grid > 148) grid = 148;
    B200_LAUNCH(k_verify_canonical)<<<grid, 256, 0, s>>>(ctx, seal, words); return cudaGetLastError();
}
cudaError_t launch_verify_fold_top(uint32_t* root_out, const uint32_t* top, uint32_t top_size, cudaStream_t s) {
    B200_LAUNCH(k_verify_fold_top)<<<1, 32, 0, s>>>(root_out, top, top_size); return cudaGetLastError();
}
cudaError_t launch_verify_constraint(uint32_t* ctx, const uint32_t* u, const uint32_t* pm, const uint32_t* z, uint32_t w_code,
                                     uint32_t w_data, uint32_t w_accum, cudaStream_t s) {
    B200_LAUNCH(k_verify_constraint)<<<1, 256, 0, s>>>(ctx, u, pm, z, w_code, w_data, w_accum); return cudaGetLastError();
}
cudaError_t launch_verify_usum(uint32_t* ctx, const uint32_t* u, const uint32_t* mp, uint32_t W, uint32_t w_accum, uint32_t T, cudaStream_t s) {
    B200_LAUNCH(k_verify_usum)<<<1, 256, 0, s>>>(ctx, u, mp, W, w_accum, T); return cudaGetLastError();
}
cudaError_t launch_verify_queries(uint32_t* ctx, const uint32_t* seal, const VerifyShape& sh, const uint32_t* mp, const uint32_t* pts,
                                  const uint32_t* fmix, const uint32_t* pos, cudaStream_t s) {
    B200_LAUNCH(k_verify_queries)<<<QUERIES, 256, 0, s>>>(ctx, seal, sh, mp, pts, fmix, pos); return cudaGetLastError();
}
cudaError_t launch_verify_finish(uint32_t* ctx, cudaStream_t s) { B200_LAUNCH(k_verify_finish)<<<1, 1, 0, s>>>(ctx); return cudaGetLastError(); }

}  // namespace b200
