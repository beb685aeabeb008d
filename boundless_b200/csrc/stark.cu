// STARK plumbing kernels around the three inner loops: synthetic witness, constraint stand-in, DEEP (K7/K8),
// FRI fold (K6, kernel 3 of the hot path) and query gathers (K9).
//
// Replaces the small HAL kernels of risc0-sys 1.5.0 (fri_fold, batch_evaluate_any, mix_poly_coeffs,
// eltwise_sum_extelem, supra_poly_divide, gather_sample) reached from
// /root/reference/prover/crates/workflow/src/tasks/prove.rs:44-52.  The rv32im witgen / eval_check (SURVEY 8a
// X1/X2; generated circuit code, unavailable) are replaced by a splitmix64 trace and a degree-4 synthetic
// constraint (DESIGN.md "Protocol"), so every kernel below runs on real data of the real shapes.
#include "internal.h"
#include "field.cuh"

namespace b200 {


// ---- witgen stand-in: element idx = splitmix64(seed, idx) mod p, Montgomery form -----------------------
__global__ void k_gen_trace(uint32_t* __restrict__ out, uint64_t seed, uint64_t count, const uint32_t* __restrict__ seed_words) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (seed_words) seed = (uint64_t)seed_words[0] | ((uint64_t)seed_words[1] << 32);   // recursion: seed = child digest words
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    // mont(z mod p) = hi*2^64 + lo*2^32 (mod p) = hi * R3 / R + lo * R2 / R, with R2 = 2^64, R3 = 2^96 (mod p)
    out[i] = fp_add(fp_mul((uint32_t)(z >> 32), 317946875u), fp_mul((uint32_t)z, R2));
}
cudaError_t launch_gen_trace(uint32_t* d_out, uint64_t seed, const uint32_t* d_seed_words, uint64_t count, cudaStream_t s) {
    if (!count) return cudaSuccess;
    B200_LAUNCH(k_gen_trace)<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(d_out, seed, count, d_seed_words);
    return cudaGetLastError();
}

// ---- accumulate stand-in: Fp4 column k (4 base columns, in place) *= mix^(k+1) --------------------------
__global__ void k_accumulate(uint32_t* __restrict__ io, uint32_t rows, uint32_t w_accum, const uint32_t* __restrict__ mix) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rows) return;
    const Fp4 m = ld_fp4(mix);
    Fp4 pw = m;
    for (uint32_t k = 0; k < w_accum / 4; k++) {
        uint32_t* b = io + (size_t)(4 * k) * rows + j;
        Fp4 d{{b[0], b[rows], b[2 * (size_t)rows], b[3 * (size_t)rows]}};
        Fp4 a = fp4_mul(d, pw);
        b[0] = a.c[0]; b[rows] = a.c[1]; b[2 * (size_t)rows] = a.c[2]; b[3 * (size_t)rows] = a.c[3];
        pw = fp4_mul(pw, m);
    }
}
cudaError_t launch_accumulate(uint32_t* d_acc_io, uint32_t rows, uint32_t w_accum, const uint32_t* d_mix, cudaStream_t s) {
    B200_LAUNCH(k_accumulate)<<<(rows + 255) / 256, 256, 0, s>>>(d_acc_io, rows, w_accum, d_mix);
    return cudaGetLastError();
}

// out[k] = base^k, k < count (Fp4).  count is a few hundred: one thread per power, square-and-multiply.
__global__ void k_powers(uint32_t* __restrict__ out, const uint32_t* __restrict__ base, uint32_t count) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    st_fp4(out + 4 * (size_t)k, fp4_pow(ld_fp4(base), k));
}
cudaError_t launch_powers(uint32_t* d_out, const uint32_t* d_base, uint32_t count, cudaStream_t s) {
    if (!count) return cudaSuccess;
    B200_LAUNCH(k_powers)<<<(count + 127) / 128, 128, 0, s>>>(d_out, d_base, count);
    return cudaGetLastError();
}

// ---- eval_check stand-in over the 4N domain -------------------------------------------------------------
// evals: W columns x D (code, data, accum).  planes: 4 x D.
__global__ void __launch_bounds__(256) k_eval_check(uint32_t* __restrict__ planes, const uint32_t* __restrict__ ev, uint32_t D,
                                                    uint32_t w_code, uint32_t w_data, uint32_t w_accum,
                                                    const uint32_t* __restrict__ pmix_g) {
    extern __shared__ uint32_t sm[];
    const uint32_t W = w_code + w_data + w_accum, nterms = W / 4 + w_accum;
    for (uint32_t i = threadIdx.x; i < nterms * 4; i += blockDim.x) sm[i] = pmix_g[i];
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D) return;
    Fp4 tot = fp4_zero();
    const uint32_t* p = ev + i;
    for (uint32_t k = 0; k < W / 4; k++) {
        const uint32_t a = __ldg(p), b = __ldg(p + D), c = __ldg(p + 2 * (size_t)D), d = __ldg(p + 3 * (size_t)D);
        p += 4 * (size_t)D;
        const uint32_t t = fp_mul(fp_mul(a, b), fp_mul(c, d));
        fp4_fma_fp(tot, ld_fp4(sm + 4 * k), t);
    }
    const uint32_t ib = (i + D - 4) & (D - 1);
    const uint32_t* acc = ev + (size_t)(w_code + w_data) * D;
    for (uint32_t a = 0; a < w_accum; a++) {
        const uint32_t t = fp_mul(fp_sub(__ldg(acc + (size_t)a * D + i), __ldg(acc + (size_t)a * D + ib)),
                                  __ldg(ev + (size_t)(a % w_code) * D + i));
        fp4_fma_fp(tot, ld_fp4(sm + 4 * (W / 4 + a)), t);
    }
#pragma unroll
    for (int e = 0; e < 4; e++) planes[(size_t)e * D + i] = tot.c[e];
}
cudaError_t launch_eval_check(uint32_t* d_planes, const uint32_t* d_evals, uint32_t lg_domain, uint32_t w_code, uint32_t w_data,
                              uint32_t w_accum, const uint32_t* d_pmix, cudaStream_t s) {
    const uint32_t D = 1u << lg_domain, nterms = (w_code + w_data + w_accum) / 4 + w_accum;
    B200_LAUNCH(k_eval_check)<<<(D + 255) / 256, 256, nterms * 16, s>>>(d_planes, d_evals, D, w_code, w_data, w_accum, d_pmix);
    return cudaGetLastError();
}

// ---- K6: FRI fold 16 -> 1 -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fri_fold(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, uint32_t in_size,
                                                  const uint32_t* __restrict__ mix_g) {
    __shared__ uint32_t pw[16 * 4];
    if (threadIdx.x < 16) st_fp4(pw + 4 * threadIdx.x, fp4_pow(ld_fp4(mix_g), threadIdx.x));
    __syncthreads();
    const uint32_t cnt = in_size / FRI_FOLD;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cnt) return;
    Fp4 tot = fp4_zero();
#pragma unroll
    for (uint32_t i = 0; i < 16; i++) {
        const size_t j = (size_t)(__brev(i) >> 28) * cnt + idx;
        Fp4 f{{__ldg(in + j), __ldg(in + in_size + j), __ldg(in + 2 * (size_t)in_size + j), __ldg(in + 3 * (size_t)in_size + j)}};
        tot = fp4_add(tot, fp4_mul(ld_fp4(pw + 4 * i), f));
    }
#pragma unroll
    for (int e = 0; e < 4; e++) out[(size_t)e * cnt + idx] = tot.c[e];
}
cudaError_t launch_fri_fold(uint32_t* d_out, const uint32_t* d_in, uint32_t in_size, const uint32_t* d_mix, cudaStream_t s) {
    const uint32_t cnt = in_size / FRI_FOLD;
    if (!cnt) return cudaSuccess;
    B200_LAUNCH(k_fri_fold)<<<(cnt + 255) / 256, 256, 0, s>>>(d_out, d_in, in_size, d_mix);
    return cudaGetLastError();
}

// ---- K7: evaluate bit-reversed coefficient columns at Fp4 points ---------------------------------------
// x^{deg(j)} with deg = bitrev_n(j): split j = (jhi, jlo), jlo = low LG_CH bits.
//   plo[jlo] = x^(bitrev_lc(jlo) << (n - lc)),  phi[jhi] = x^(bitrev_(n-lc)(jhi))
// scratch layout (words): [plo_a 4*CH][phi_a 4*NCH][plo_b 4*CH][phi_b 4*NCH][partial_a 4*count*NCH][partial_b ...]
constexpr uint32_t EV_LG_CH = 12;
static inline uint32_t ev_lc(uint32_t lg_n) { return lg_n < EV_LG_CH ? lg_n : EV_LG_CH; }
size_t evaluate_scratch_words(uint32_t lg_n, uint32_t count) {
    const uint32_t lc = ev_lc(lg_n);
    const size_t CH = (size_t)1 << lc, NCH = (size_t)1 << (lg_n - lc);
    return 2 * (4 * CH + 4 * NCH) + 2 * 4 * (size_t)count * NCH;
}
__global__ void k_ev_tables(uint32_t* __restrict__ plo, uint32_t* __restrict__ phi, const uint32_t* __restrict__ x, uint32_t lg_n, uint32_t lc) {
    const uint32_t CH = 1u << lc, NCH = 1u << (lg_n - lc);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const Fp4 X = ld_fp4(x);
    if (i < CH) st_fp4(plo + 4 * (size_t)i, fp4_pow(X, (uint64_t)bitrev(i, lc) << (lg_n - lc)));
    if (i < NCH) st_fp4(phi + 4 * (size_t)i, fp4_pow(X, bitrev(i, lg_n - lc)));
}
// one CTA per (chunk, column): partial = sum_jlo coeff[j] * plo[jlo]
__global__ void __launch_bounds__(256) k_ev_partial(uint32_t* __restrict__ part_a, uint32_t* __restrict__ part_b, const uint32_t* __restrict__ coeffs,
                                                    uint32_t lg_n, uint32_t lc, const uint32_t* __restrict__ plo_a,
                                                    const uint32_t* __restrict__ plo_b, uint32_t b0, uint32_t b1) {
    const uint32_t CH = 1u << lc, NCH = 1u << (lg_n - lc);
    const uint32_t chunk = blockIdx.x, col = blockIdx.y;
    const bool do_b = plo_b && col >= b0 && col < b1;
    const uint32_t* src = coeffs + ((size_t)col << lg_n) + ((size_t)chunk << lc);
    Fp4 sa = fp4_zero(), sb = fp4_zero();
    for (uint32_t j = threadIdx.x; j < CH; j += blockDim.x) {
        const uint32_t v = __ldg(src + j);
        fp4_fma_fp(sa, ld_fp4(plo_a + 4 * (size_t)j), v);
        if (do_b) fp4_fma_fp(sb, ld_fp4(plo_b + 4 * (size_t)j), v);
    }
    __shared__ uint32_t red[2][256 * 4];
    for (int e = 0; e < 4; e++) { red[0][threadIdx.x * 4 + e] = sa.c[e]; red[1][threadIdx.x * 4 + e] = sb.c[e]; }
    __syncthreads();
    for (uint32_t st = blockDim.x / 2; st >= 1; st >>= 1) {
        if (threadIdx.x < st) {
            for (int w = 0; w < 2; w++)
                for (int e = 0; e < 4; e++)
                    red[w][threadIdx.x * 4 + e] = fp_add(red[w][threadIdx.x * 4 + e], red[w][(threadIdx.x + st) * 4 + e]);
        }
        __syncthreads();
    }
    if (threadIdx.x < 4) {
        part_a[4 * ((size_t)col * NCH + chunk) + threadIdx.x] = red[0][threadIdx.x];
        if (do_b) part_b[4 * ((size_t)(col - b0) * NCH + chunk) + threadIdx.x] = red[1][threadIdx.x];
    }
}
// out[col] = sum_chunk partial[col][chunk] * phi[chunk]; one warp-sized CTA per column
__global__ void __launch_bounds__(128) k_ev_reduce(uint32_t* __restrict__ out, const uint32_t* __restrict__ part, const uint32_t* __restrict__ phi,
                                                   uint32_t NCH) {
    const uint32_t col = blockIdx.x;
    Fp4 s = fp4_zero();
    for (uint32_t c = threadIdx.x; c < NCH; c += blockDim.x)
        s = fp4_add(s, fp4_mul(ld_fp4(part + 4 * ((size_t)col * NCH + c)), ld_fp4(phi + 4 * (size_t)c)));
    __shared__ uint32_t red[128 * 4];
    for (int e = 0; e < 4; e++) red[threadIdx.x * 4 + e] = s.c[e];
    __syncthreads();
    for (uint32_t st = blockDim.x / 2; st >= 1; st >>= 1) {
        if (threadIdx.x < st)
            for (int e = 0; e < 4; e++) red[threadIdx.x * 4 + e] = fp_add(red[threadIdx.x * 4 + e], red[(threadIdx.x + st) * 4 + e]);
        __syncthreads();
    }
    if (threadIdx.x < 4) out[4 * (size_t)col + threadIdx.x] = red[threadIdx.x];
}
cudaError_t launch_evaluate(uint32_t* d_out_a, uint32_t* d_out_b, const uint32_t* d_coeffs, uint32_t lg_n, uint32_t count,
                            const uint32_t* d_x, const uint32_t* d_xb, uint32_t b0, uint32_t b1, uint32_t* d_scratch, cudaStream_t s) {
    if (!count) return cudaSuccess;
    const uint32_t lc = ev_lc(lg_n);
    const uint32_t CH = 1u << lc, NCH = 1u << (lg_n - lc);
    uint32_t* plo_a = d_scratch; uint32_t* phi_a = plo_a + 4 * (size_t)CH;
    uint32_t* plo_b = phi_a + 4 * (size_t)NCH; uint32_t* phi_b = plo_b + 4 * (size_t)CH;
    uint32_t* part_a = phi_b + 4 * (size_t)NCH; uint32_t* part_b = part_a + 4 * (size_t)count * NCH;
    const uint32_t tn = CH > NCH ? CH : NCH;
    B200_LAUNCH(k_ev_tables)<<<(tn + 127) / 128, 128, 0, s>>>(plo_a, phi_a, d_x, lg_n, lc);
    if (d_xb) B200_LAUNCH(k_ev_tables)<<<(tn + 127) / 128, 128, 0, s>>>(plo_b, phi_b, d_xb, lg_n, lc);
    dim3 grid(NCH, count);
    B200_LAUNCH(k_ev_partial)<<<grid, 256, 0, s>>>(part_a, part_b, d_coeffs, lg_n, lc, plo_a, d_xb ? plo_b : nullptr, b0, b1);
    B200_LAUNCH(k_ev_reduce)<<<count, 128, 0, s>>>(d_out_a, part_a, phi_a, NCH);
    if (d_xb && b1 > b0) B200_LAUNCH(k_ev_reduce)<<<b1 - b0, 128, 0, s>>>(d_out_b, part_b, phi_b, NCH);
    return cudaGetLastError();
}

__global__ void k_deep_points(uint32_t* pts, const uint32_t* z, uint32_t rou_rev_n) {
    const Fp4 Z = ld_fp4(z);
    st_fp4(pts, Z);
    st_fp4(pts + 4, fp4_mul_fp(Z, rou_rev_n));
    const Fp4 z2 = fp4_mul(Z, Z);
    st_fp4(pts + 8, fp4_mul(z2, z2));
}
cudaError_t launch_deep_points(uint32_t* d_pts, const uint32_t* d_z, uint32_t rou_rev_n, cudaStream_t s) {
    B200_LAUNCH(k_deep_points)<<<1, 1, 0, s>>>(d_pts, d_z, rou_rev_n);
    return cudaGetLastError();
}

// ---- K8: DEEP combination, division by (x - point), sum --------------------------------------------------
// (1) combos[pt][deg] = sum_t mix^t * coeff_t[bitrev(deg)] - [deg == 0] * sum_t mix^t u_t   (natural degree order, AoS Fp4)
__global__ void __launch_bounds__(256) k_deep_mix(uint32_t* __restrict__ combos, const uint32_t* __restrict__ coeffs,
                                                  const uint32_t* __restrict__ chk, const uint32_t* __restrict__ u,
                                                  const uint32_t* __restrict__ mp_g, uint32_t lg_n, uint32_t W, uint32_t w_accum) {
    extern __shared__ uint32_t mp[];
    const uint32_t T = W + w_accum + CHECK_COLS, N = 1u << lg_n;
    for (uint32_t i = threadIdx.x; i < T * 4; i += blockDim.x) mp[i] = mp_g[i];
    __syncthreads();
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    Fp4 a0 = fp4_zero(), a1 = fp4_zero(), a2 = fp4_zero();
    const uint32_t* p = coeffs + j;
    const uint32_t acc0 = W - w_accum;
    for (uint32_t c = 0; c < acc0; c++) { fp4_fma_fp(a0, ld_fp4(mp + 4 * c), __ldg(p)); p += N; }
    for (uint32_t c = acc0; c < W; c++) {
        const uint32_t v = __ldg(p); p += N;
        fp4_fma_fp(a0, ld_fp4(mp + 4 * c), v);
        fp4_fma_fp(a1, ld_fp4(mp + 4 * (W + c - acc0)), v);
    }
    const uint32_t* q = chk + j;
    for (uint32_t c = 0; c < CHECK_COLS; c++) { fp4_fma_fp(a2, ld_fp4(mp + 4 * (W + w_accum + c)), __ldg(q)); q += N; }
    const uint32_t d = bitrev(j, lg_n);
    if (d == 0) {
        Fp4 s0 = fp4_zero(), s1 = fp4_zero(), s2 = fp4_zero();
        for (uint32_t t = 0; t < W; t++) s0 = fp4_add(s0, fp4_mul(ld_fp4(mp + 4 * t), ld_fp4(u + 4 * (size_t)t)));
        for (uint32_t t = W; t < W + w_accum; t++) s1 = fp4_add(s1, fp4_mul(ld_fp4(mp + 4 * t), ld_fp4(u + 4 * (size_t)t)));
        for (uint32_t t = W + w_accum; t < T; t++) s2 = fp4_add(s2, fp4_mul(ld_fp4(mp + 4 * t), ld_fp4(u + 4 * (size_t)t)));
        a0 = fp4_sub(a0, s0); a1 = fp4_sub(a1, s1); a2 = fp4_sub(a2, s2);
    }
    st_fp4(combos + 4 * ((size_t)0 * N + d), a0);
    st_fp4(combos + 4 * ((size_t)1 * N + d), a1);
    st_fp4(combos + 4 * ((size_t)2 * N + d), a2);
}
// Division of c(x) by (x - a): b_d = sum_{j>d} c_j a^(j-d-1).  Chunks of DV_CH degrees, DV_E per thread.
constexpr uint32_t DV_T = 256, DV_E = 8, DV_CH = DV_T * DV_E;
// (2) chunk value V_k = sum_{d in chunk} c_d a^(d - start)
__global__ void __launch_bounds__(DV_T) k_deep_chunk_vals(uint32_t* __restrict__ vals, const uint32_t* __restrict__ combos,
                                                          const uint32_t* __restrict__ pts, uint32_t N, uint32_t nchunks) {
    const uint32_t pt = blockIdx.y, chunk = blockIdx.x;
    const Fp4 a = ld_fp4(pts + 4 * pt);
    const uint32_t start = chunk * DV_CH;
    const uint32_t cnt = N - start < DV_CH ? N - start : DV_CH;
    const uint32_t t0 = threadIdx.x * DV_E;
    Fp4 v = fp4_zero();
    if (t0 < cnt) {
        const uint32_t* c = combos + 4 * ((size_t)pt * N + start + t0);
        const uint32_t m = cnt - t0 < DV_E ? cnt - t0 : DV_E;
        for (int i = (int)m - 1; i >= 0; i--) v = fp4_add(fp4_mul(v, a), ld_fp4(c + 4 * i));
        v = fp4_mul(v, fp4_pow(a, t0));
    }
    __shared__ uint32_t red[DV_T * 4];
    for (int e = 0; e < 4; e++) red[threadIdx.x * 4 + e] = v.c[e];
    __syncthreads();
    for (uint32_t st = DV_T / 2; st >= 1; st >>= 1) {
        if (threadIdx.x < st)
            for (int e = 0; e < 4; e++) red[threadIdx.x * 4 + e] = fp_add(red[threadIdx.x * 4 + e], red[(threadIdx.x + st) * 4 + e]);
        __syncthreads();
    }
    if (threadIdx.x < 4) vals[4 * ((size_t)pt * nchunks + chunk) + threadIdx.x] = red[threadIdx.x];
}
// (3) carry into chunk k from above: B_k = sum_{m>k} V_m a^(CH*(m-k-1)); one thread per point (nchunks <= ~2048)
__global__ void k_deep_chunk_scan(uint32_t* __restrict__ carry, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ pts,
                                  uint32_t nchunks, uint32_t* __restrict__ rem) {
    const uint32_t pt = blockIdx.x;
    const Fp4 ach = fp4_pow(ld_fp4(pts + 4 * pt), DV_CH);
    Fp4 b = fp4_zero();
    for (int k = (int)nchunks - 1; k >= 0; k--) {
        st_fp4(carry + 4 * ((size_t)pt * nchunks + k), b);
        b = fp4_add(fp4_mul(b, ach), ld_fp4(vals + 4 * ((size_t)pt * nchunks + k)));
    }
    if (rem) st_fp4(rem + 4 * pt, b);      // value of the whole polynomial at the point = remainder of the division
}
// (4) quotients of the three points summed, written as 4 planes in bit-reversed order
__global__ void __launch_bounds__(DV_T) k_deep_divide(uint32_t* __restrict__ planes, const uint32_t* __restrict__ combos,
                                                      const uint32_t* __restrict__ carry, const uint32_t* __restrict__ pts, uint32_t lg_n,
                                                      uint32_t nchunks) {
    const uint32_t N = 1u << lg_n, chunk = blockIdx.x, start = chunk * DV_CH;
    const uint32_t cnt = N - start < DV_CH ? N - start : DV_CH;
    const uint32_t t0 = threadIdx.x * DV_E;
    __shared__ uint32_t sc[DV_T * 4];
    Fp4 q[DV_E];
#pragma unroll
    for (int i = 0; i < (int)DV_E; i++) q[i] = fp4_zero();
    for (uint32_t pt = 0; pt < 3; pt++) {
        const Fp4 a = ld_fp4(pts + 4 * pt);
        const Fp4 aE = fp4_pow(a, DV_E);
        // thread value v_t = sum_i c_{t0+i} a^i
        Fp4 c[DV_E];
        Fp4 v = fp4_zero();
        const uint32_t m = t0 < cnt ? (cnt - t0 < DV_E ? cnt - t0 : DV_E) : 0;
#pragma unroll
        for (int i = (int)DV_E - 1; i >= 0; i--) {
            c[i] = (uint32_t)i < m ? ld_fp4(combos + 4 * ((size_t)pt * N + start + t0 + i)) : fp4_zero();
            v = fp4_add(fp4_mul(v, a), c[i]);
        }
        // suffix scan over threads: S_t = sum_{t' >= t} v_t' a^(E*(t'-t)); Hillis-Steele with multiplier doubling
        __syncthreads();
        for (int e = 0; e < 4; e++) sc[threadIdx.x * 4 + e] = v.c[e];
        __syncthreads();
        Fp4 mult = aE;
        for (uint32_t off = 1; off < DV_T; off <<= 1) {
            Fp4 add = fp4_zero();
            const bool has = threadIdx.x + off < DV_T;
            if (has) add = fp4_mul(ld_fp4(sc + 4 * (threadIdx.x + off)), mult);
            __syncthreads();
            if (has) { v = fp4_add(v, add); for (int e = 0; e < 4; e++) sc[threadIdx.x * 4 + e] = v.c[e]; }
            __syncthreads();
            mult = fp4_mul(mult, mult);
        }
        // carry into this thread's run from everything above it:
        //   above-in-chunk = S_{t+1};  above-chunk = B_chunk * a^(E*(T-1-t))   (a^(distance from run end to chunk end))
        Fp4 b = (threadIdx.x + 1 < DV_T) ? ld_fp4(sc + 4 * (threadIdx.x + 1)) : fp4_zero();
        const Fp4 B = ld_fp4(carry + 4 * ((size_t)pt * nchunks + chunk));
        b = fp4_add(b, fp4_mul(B, fp4_pow(aE, DV_T - 1 - threadIdx.x)));
        // walk down the run: quotient coefficient at degree d is b (the carry above d), then b = c_d + a*b
#pragma unroll
        for (int i = (int)DV_E - 1; i >= 0; i--) {
            q[i] = fp4_add(q[i], b);
            b = fp4_add(c[i], fp4_mul(b, a));
        }
    }
#pragma unroll
    for (int i = 0; i < (int)DV_E; i++) {
        const uint32_t d = start + t0 + i;
        if (d < N) {
            const uint32_t j = bitrev(d, lg_n);
#pragma unroll
            for (int e = 0; e < 4; e++) planes[(size_t)e * N + j] = q[i].c[e];
        }
    }
}
cudaError_t launch_deep(const DeepArgs& a, cudaStream_t s) {
    const uint32_t N = 1u << a.lg_n, T = a.W + a.w_accum + CHECK_COLS;
    const uint32_t nchunks = (N + DV_CH - 1) / DV_CH;
    B200_LAUNCH(k_deep_mix)<<<(N + 255) / 256, 256, T * 16, s>>>(a.combos, a.coeffs, a.check_coeffs, a.u, a.mix_pows, a.lg_n, a.W, a.w_accum);
    dim3 g2(nchunks, 3);
    B200_LAUNCH(k_deep_chunk_vals)<<<g2, DV_T, 0, s>>>(a.chunk_vals, a.combos, a.pts, N, nchunks);
    B200_LAUNCH(k_deep_chunk_scan)<<<3, 1, 0, s>>>(a.chunk_carry, a.chunk_vals, a.pts, nchunks, nullptr);
    B200_LAUNCH(k_deep_divide)<<<nchunks, DV_T, 0, s>>>(a.f_planes, a.combos, a.chunk_carry, a.pts, a.lg_n, nchunks);
    return cudaGetLastError();
}

// ---- supra_poly_divide as a standalone operation: one Fp4 polynomial (AoS, natural order) divided in place by (x - z) --------
__global__ void __launch_bounds__(DV_T) k_poly_divide_apply(uint32_t* __restrict__ poly, const uint32_t* __restrict__ carry,
                                                            const uint32_t* __restrict__ zp, uint32_t N) {
    const uint32_t chunk = blockIdx.x, start = chunk * DV_CH;
    const uint32_t cnt = N - start < DV_CH ? N - start : DV_CH;
    const uint32_t t0 = threadIdx.x * DV_E;
    __shared__ uint32_t sc[DV_T * 4];
    const Fp4 a = ld_fp4(zp);
    const Fp4 aE = fp4_pow(a, DV_E);
    Fp4 c[DV_E];
    Fp4 v = fp4_zero();
    const uint32_t m = t0 < cnt ? (cnt - t0 < DV_E ? cnt - t0 : DV_E) : 0;
#pragma unroll
    for (int i = (int)DV_E - 1; i >= 0; i--) {
        c[i] = (uint32_t)i < m ? ld_fp4(poly + 4 * ((size_t)start + t0 + i)) : fp4_zero();
        v = fp4_add(fp4_mul(v, a), c[i]);
    }
    for (int e = 0; e < 4; e++) sc[threadIdx.x * 4 + e] = v.c[e];
    __syncthreads();
    Fp4 mult = aE;
    for (uint32_t off = 1; off < DV_T; off <<= 1) {      // suffix scan S_t = sum_{t' >= t} v_t' a^(E (t' - t))
        Fp4 add = fp4_zero();
        const bool has = threadIdx.x + off < DV_T;
        if (has) add = fp4_mul(ld_fp4(sc + 4 * (threadIdx.x + off)), mult);
        __syncthreads();
        if (has) { v = fp4_add(v, add); for (int e = 0; e < 4; e++) sc[threadIdx.x * 4 + e] = v.c[e]; }
        __syncthreads();
        mult = fp4_mul(mult, mult);
    }
    Fp4 b = (threadIdx.x + 1 < DV_T) ? ld_fp4(sc + 4 * (threadIdx.x + 1)) : fp4_zero();
    b = fp4_add(b, fp4_mul(ld_fp4(carry + 4 * (size_t)chunk), fp4_pow(aE, DV_T - 1 - threadIdx.x)));
#pragma unroll
    for (int i = (int)DV_E - 1; i >= 0; i--) {
        if ((uint32_t)i < m) st_fp4(poly + 4 * ((size_t)start + t0 + i), b);
        b = fp4_add(c[i], fp4_mul(b, a));
    }
}
size_t poly_divide_scratch_words(uint32_t size) { return (size_t)8 * ((size + DV_CH - 1) / DV_CH) + 8; }
cudaError_t launch_poly_divide(uint32_t* d_poly, uint32_t size, uint32_t* d_remainder, const uint32_t* d_pow, uint32_t* d_scratch,
                               cudaStream_t s) {
    if (size == 0) return cudaMemsetAsync(d_remainder, 0, 16, s);
    const uint32_t nchunks = (size + DV_CH - 1) / DV_CH;
    uint32_t* vals = d_scratch;
    uint32_t* carry = d_scratch + (size_t)4 * nchunks;
    B200_LAUNCH(k_deep_chunk_vals)<<<dim3(nchunks, 1), DV_T, 0, s>>>(vals, d_poly, d_pow, size, nchunks);
    B200_LAUNCH(k_deep_chunk_scan)<<<1, 1, 0, s>>>(carry, vals, d_pow, nchunks, d_remainder);
    B200_LAUNCH(k_poly_divide_apply)<<<nchunks, DV_T, 0, s>>>(d_poly, carry, d_pow, size);
    return cudaGetLastError();
}

// ---- K9: query openings -----------------------------------------------------------------------------------
// one CTA per (tree, query): leaf values then sibling digests up to (excluding) the top layer
__global__ void __launch_bounds__(128) k_gather(uint32_t* __restrict__ seal, uint32_t query_base, uint32_t query_words,
                                                const uint32_t* __restrict__ pos_g, const GatherTree* __restrict__ trees) {
    const GatherTree t = trees[blockIdx.x];
    const uint32_t q = blockIdx.y;
    // position for this tree: trace groups use pos; FRI round r uses pos reduced modulo each earlier round's rows chain
    uint32_t pos = pos_g[q];
    if (t.pos_shift_mod) pos &= (t.pos_shift_mod - 1);
    uint32_t* dst = seal + query_base + (size_t)q * query_words + t.seal_off;
    for (uint32_t c = threadIdx.x; c < t.cols; c += blockDim.x) dst[c] = t.matrix[(size_t)c * t.rows + pos];
    dst += t.cols;
    uint32_t idx = pos + t.rows, level = 0;
    while (idx >= 2 * t.top_size) {
        if (threadIdx.x < 8) dst[level * 8 + threadIdx.x] = t.nodes[(size_t)(idx ^ 1) * 8 + threadIdx.x];
        idx >>= 1; level++;
    }
}
cudaError_t launch_gather_queries(uint32_t* d_seal, uint32_t query_base, uint32_t query_words, const uint32_t* d_pos,
                                  const GatherTree* d_trees, uint32_t n_trees, cudaStream_t s) {
    dim3 grid(n_trees, QUERIES);
    B200_LAUNCH(k_gather)<<<grid, 128, 0, s>>>(d_seal, query_base, query_words, d_pos, d_trees);
    return cudaGetLastError();
}

}  // namespace b200
