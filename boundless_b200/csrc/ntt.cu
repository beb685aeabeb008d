// Batched NTT / iNTT / expand+NTT / zk_shift over BabyBear for sm_100a (kernel 1 of the hot path).
//
// Replaces risc0-sys 1.5.0's sppark_batch_iNTT / sppark_batch_NTT / sppark_batch_expand / sppark_batch_zk_shift
// (un-vendored), reached from
// /root/reference/prover/crates/workflow/src/tasks/prove.rs:44-52 (prove_segment) and :96-104 (lift).
// Conventions are the reference's: iNTT maps natural-order evaluations to BIT-REVERSED coefficients scaled
// by 1/n; the forward transform maps bit-reversed coefficients to natural-order evaluations; `expand` is the
// x4 replicate that equals zero padding in bit-reversed order, so the first lg_blowup levels are skipped.
//
// Design (B200): a size-2^m transform is two HBM passes (six-step): N = N1*N2 (three for 2^25 / 2^26).
//   iNTT : pass A  strided tile [N1 rows][8 cols], DIF along rows, in place;
//          pass B  contiguous rows of N2: the inter-pass twiddle w^-(i0*bitrev(r)) (with the 1/N scale folded in) on the loads, DIF,
//                  optionally the zk_shift factor on the stores, in place.  (Sizes without a per-element table and rows > 1024 keep
//                  the twiddle in pass A's epilogue, as in round 1.)
//   fwd  : pass 1  contiguous row of N2 coefficients -> replicated x4 -> DIT (levels above the blow-up) -> inter-pass twiddle ->
//                  written as a row of 4*N2;
//          pass 2  strided tile, DIT along rows, in place.
// Every pass moves each element HBM -> registers / shared memory -> HBM once in >= 32-byte runs.  The hot shapes (1024-row strided
// tiles, 1024-value rows) are radix-32 register kernels: two five-level stages with one exchange through shared memory; the other
// shapes are radix-16 stages (4 levels per shared-memory round trip).  Stage twiddles come from one compact per-level table
// (TMA-staged into shared memory in the strided passes, read-only path elsewhere); inter-pass twiddles and zk_shift factors from
// per-element tables in data layout (one coalesced load + one multiply; tables.cpp get_full_table) or, above their size cap, from a
// two-table decomposition w^e = lo[e & mask] * hi[e >> h].
// Kernel families, fastest first (the host dispatch falls through to the next when a shape is not covered):
//   k_ntt_strided_r32d / k_ntt_invb_r32 / k_ntt_fwd1 / k_ntt_invb   the segment's shapes
//   k_ntt_strided_r32 / k_ntt_strided_pf                            cp.async double-buffered persistent strided passes
//   k_ntt_strided_c / k_ntt_contig_c                                compile-time sizes, one tile per CTA
//   k_ntt_strided   / k_ntt_contig                                  runtime sizes (any 2^k)
#include "internal.h"
#include "field.cuh"
#include <cstdlib>

namespace b200 {

// Address functors: operator()(pos) is the general form; base(pos) + off(delta) is the split form the compile-time stages
// use: for the stage gathers (delta = j*q with the group offset o < q) phys(base + delta) == base(base) + off(delta), and
// off(delta) is an immediate, so a gather/scatter costs no address arithmetic per element.
struct AddrStrided {
    uint32_t lgTW, t;
    __device__ __forceinline__ uint32_t operator()(uint32_t pos) const { return (pos << lgTW) + t; }
    __device__ __forceinline__ uint32_t base(uint32_t pos) const { return (pos << lgTW) + t; }
    __device__ __forceinline__ uint32_t off(uint32_t d) const { return d << lgTW; }
};
struct AddrContig {
    uint32_t base;
    // one pad word per 16 keeps the radix-16 gathers (stride q < 32) free of bank conflicts
    __device__ __forceinline__ uint32_t operator()(uint32_t pos) const { return base + pos + (pos >> 4); }
    __device__ __forceinline__ uint32_t base_of(uint32_t pos) const { return base + pos + (pos >> 4); }
    __device__ __forceinline__ static constexpr uint32_t off(uint32_t d) { return d + (d >> 4); }
};

// Twiddles are table CONSTANTS, so they are stored as Shoup pairs {w, floor(w * 2^32 / p)} with w the plain (non-Montgomery)
// value: x~ * w mod p keeps x~ in Montgomery form and costs IMAD.HI + 2 IMAD (8 multiplier-pipe cycles) instead of
// IMAD.WIDE + IMAD + IMAD.HI (10).  x may be any u32, the result is canonical.
typedef uint2 tw_t;
__device__ __forceinline__ uint32_t mul_tw(uint32_t x, const tw_t w) {
    const uint32_t q = __umulhi(x, w.y);
    const uint32_t r = x * w.x - q * P;          // in [0, 2p)
    return addmin(r, 0u - P, r);
}
// Butterfly adds for the transforms.  ptxas splits plain adds between IADD3 (ALU pipe) and IMAD.IADD (the multiplier pipe, which the
// twiddle products already load more than the ALU); B200_NTT_ADD_V bit 0 / bit 1 write the add / the subtract as VIADDMNMX, which only
// the ALU executes (the same device B200_P2_ZALL uses in poseidon2.cuh).  Measured: profiles/ntt_addv_r02.txt.
#ifndef B200_NTT_ADD_V
#define B200_NTT_ADD_V 0
#endif
__device__ __forceinline__ uint32_t nt_add(uint32_t a, uint32_t b) {
    const uint32_t s = (B200_NTT_ADD_V & 1) ? addmin(a, b, 0xffffffffu) : a + b;
    return addmin(s, 0u - P, s);
}
__device__ __forceinline__ uint32_t nt_sub(uint32_t a, uint32_t b) {
    const uint32_t d = (B200_NTT_ADD_V & 2) ? addmin(a, 0u - b, 0xffffffffu) : a - b;       // VIADDMNMX takes -b as an operand modifier
    return addmin(d, P, d);
}
// a - b + p in (0, 2p) for a, b < p: enough for mul_tw, which takes any u32 (one three-input add instead of a reduced subtract)
__device__ __forceinline__ uint32_t nt_subp(uint32_t a, uint32_t b) { return a - b + P; }
// v * lo * hi with both factors from the two-table decomposition w^e = lo[e & mask] * hi[e >> h]
__device__ __forceinline__ uint32_t pow_apply(uint32_t v, const tw_t lo, const tw_t hi) { return mul_tw(mul_tw(v, lo), hi); }

// One radix-2^K register stage: gathers the 2^K elements base + j*q of the block of size 2^lb that contains them,
// runs K butterfly levels in registers and scatters them back.  tw is the compact table tw[2^(l-1)+i] = w_{2^l}^i.
template <int K, bool DIF, typename ADDR>
__device__ __forceinline__ void ntt_stage(uint32_t* s, const tw_t* __restrict__ tw, uint32_t lb, uint32_t gidx, ADDR addr) {
    constexpr int R = 1 << K;
    const uint32_t q = (1u << lb) >> K;
    const uint32_t b = gidx >> (lb - K), o = gidx & (q - 1);
    const uint32_t base = (b << lb) + o;
    uint32_t x[R];
#pragma unroll
    for (int j = 0; j < R; j++) x[j] = s[addr(base + j * q)];
#pragma unroll
    for (int ll = 0; ll < K; ll++) {
        const int l = DIF ? ll : (K - 1 - ll);
        const int half = R >> (l + 1);
        const uint32_t twbase = ((1u << lb) >> (l + 1)) + o;
#pragma unroll
        for (int j = 0; j < R; j++) {
            if ((j & half) == 0) {
                const tw_t w = tw[twbase + (j & (half - 1)) * q];
                const uint32_t a = x[j], bb = x[j + half];
                if (DIF) {
                    x[j] = nt_add(a, bb);
                    x[j + half] = mul_tw(nt_subp(a, bb), w);
                } else {
                    const uint32_t t = mul_tw(bb, w);
                    x[j] = nt_add(a, t);
                    x[j + half] = nt_sub(a, t);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < R; j++) s[addr(base + j * q)] = x[j];
}

template <bool DIF, typename ADDR>
__device__ __forceinline__ void ntt_stage_k(int K, uint32_t* s, const tw_t* tw, uint32_t lb, uint32_t gidx, ADDR addr) {
    switch (K) {
        case 4: ntt_stage<4, DIF>(s, tw, lb, gidx, addr); break;
        case 3: ntt_stage<3, DIF>(s, tw, lb, gidx, addr); break;
        case 2: ntt_stage<2, DIF>(s, tw, lb, gidx, addr); break;
        default: ntt_stage<1, DIF>(s, tw, lb, gidx, addr); break;
    }
}

struct PowTab {
    const tw_t* lo; const tw_t* hi; uint32_t h, mask;
    __device__ __forceinline__ uint32_t apply(uint32_t v, uint32_t e) const { return pow_apply(v, lo[e & mask], hi[e >> h]); }
};

// ---- strided pass: tile [L][TW], FFT along rows -------------------------------------------------------
// data element (row r, col c) of polynomial p lives at data[p*poly_stride + r*row_stride + c].
template <bool DIF>
__global__ void __launch_bounds__(512) k_ntt_strided(uint32_t* __restrict__ data, uint32_t logL, uint32_t lgTW, uint32_t row_stride,
                                                     uint32_t tiles_per_poly, size_t poly_stride,
                                                     const tw_t* __restrict__ tw_g, const tw_t* __restrict__ pow_g,
                                                     uint32_t lg_m) {
    extern __shared__ uint32_t smem[];
    const uint32_t L = 1u << logL, TW = 1u << lgTW;
    uint32_t* tile = smem;                  // L * TW
    tw_t* tw = reinterpret_cast<tw_t*>(tile + (L << lgTW));      // L pairs
    tw_t* pw = tw + L;                      // 2^h + 2^(m-h) pairs when pow_g
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    const uint32_t poly = blockIdx.x / tiles_per_poly, tcol = blockIdx.x % tiles_per_poly;
    uint32_t* g = data + (size_t)poly * poly_stride + ((size_t)tcol << lgTW);

    for (uint32_t i = tid; i < L; i += nth) tw[i] = tw_g[i];
    const uint32_t h = (lg_m + 1) / 2;
    if (pow_g) { const uint32_t n = (1u << h) + (1u << (lg_m - h)); for (uint32_t i = tid; i < n; i += nth) pw[i] = pow_g[i]; }
    for (uint32_t i = tid; i < (L << lgTW); i += nth) tile[i] = g[(size_t)(i >> lgTW) * row_stride + (i & (TW - 1))];
    __syncthreads();

    if (DIF) {
        uint32_t lb = logL;
        while (lb > 0) {
            const int K = lb >= 4 ? 4 : (int)lb;
            const uint32_t total = (L >> K) << lgTW;
            for (uint32_t w = tid; w < total; w += nth) ntt_stage_k<true>(K, tile, tw, lb, w >> lgTW, AddrStrided{lgTW, w & (TW - 1)});
            __syncthreads();
            lb -= K;
        }
    } else {
        uint32_t done = 0;
        while (done < logL) {
            const int K = (logL - done) >= 4 ? 4 : (int)(logL - done);
            const uint32_t lb = done + K;
            const uint32_t total = (L >> K) << lgTW;
            for (uint32_t w = tid; w < total; w += nth) ntt_stage_k<false>(K, tile, tw, lb, w >> lgTW, AddrStrided{lgTW, w & (TW - 1)});
            __syncthreads();
            done += K;
        }
    }
    if (pow_g) {
        PowTab pt{pw, pw + (1u << h), h, (1u << h) - 1};
        const uint32_t mmask = (lg_m >= 32) ? 0xffffffffu : ((1u << lg_m) - 1);
        for (uint32_t i = tid; i < (L << lgTW); i += nth) {
            const uint32_t r = i >> lgTW, c = (tcol << lgTW) + (i & (TW - 1));
            const uint32_t e = (c * bitrev(r, logL)) & mmask;
            g[(size_t)r * row_stride + (i & (TW - 1))] = pt.apply(tile[i], e);
        }
    } else {
        for (uint32_t i = tid; i < (L << lgTW); i += nth) g[(size_t)(i >> lgTW) * row_stride + (i & (TW - 1))] = tile[i];
    }
}

// ---- contiguous pass: rows of Lc elements, FFT along the row -----------------------------------------
// input row (length Lc >> lg_e) is replicated x 2^lg_e into shared memory (forward/expand) or copied (lg_e = 0).
template <bool DIF>
__global__ void __launch_bounds__(256) k_ntt_contig(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, uint32_t logLc,
                                                    uint32_t lg_e, uint32_t rows_per_cta, uint32_t rows_per_poly, uint32_t total_rows,
                                                    size_t in_poly_stride, size_t out_poly_stride,
                                                    const tw_t* __restrict__ tw_g, const tw_t* __restrict__ pow_g,
                                                    uint32_t lg_m, uint32_t lg_rows, uint32_t scale) {
    extern __shared__ uint32_t smem[];
    const uint32_t Lc = 1u << logLc, Lin = Lc >> lg_e;
    const uint32_t rowpad = Lc + (Lc >> 4);
    uint32_t* tile = smem;                              // rows_per_cta * rowpad
    tw_t* tw = reinterpret_cast<tw_t*>(tile + ((rows_per_cta * rowpad + 1) & ~1u));        // Lc pairs
    tw_t* pw = tw + Lc;
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    const uint32_t row0 = blockIdx.x * rows_per_cta;

    for (uint32_t i = tid; i < Lc; i += nth) tw[i] = tw_g[i];
    const uint32_t h = (lg_m + 1) / 2;
    if (pow_g) { const uint32_t n = (1u << h) + (1u << (lg_m - h)); for (uint32_t i = tid; i < n; i += nth) pw[i] = pow_g[i]; }
    for (uint32_t i = tid; i < rows_per_cta * Lin; i += nth) {
        const uint32_t rr = i / Lin, k = i - rr * Lin, R = row0 + rr;
        if (R < total_rows) {
            const uint32_t poly = R / rows_per_poly, rho = R - poly * rows_per_poly;
            const uint32_t v = in[(size_t)poly * in_poly_stride + (size_t)rho * Lin + k];
            AddrContig ad{rr * rowpad};
            for (uint32_t e = 0; e < (1u << lg_e); e++) tile[ad((k << lg_e) + e)] = v;
        }
    }
    __syncthreads();

    const uint32_t nlev = logLc - lg_e;
    if (DIF) {
        uint32_t lb = logLc;        // lg_e == 0 for DIF
        while (lb > 0) {
            const int K = lb >= 4 ? 4 : (int)lb;
            const uint32_t per_row = Lc >> K, total = rows_per_cta * per_row;
            for (uint32_t w = tid; w < total; w += nth) { const uint32_t rr = w / per_row; ntt_stage_k<true>(K, tile, tw, lb, w - rr * per_row, AddrContig{rr * rowpad}); }
            __syncthreads();
            lb -= K;
        }
    } else {
        uint32_t done = 0;
        while (done < nlev) {
            const int K = (nlev - done) >= 4 ? 4 : (int)(nlev - done);
            const uint32_t lb = lg_e + done + K;
            const uint32_t per_row = Lc >> K, total = rows_per_cta * per_row;
            for (uint32_t w = tid; w < total; w += nth) { const uint32_t rr = w / per_row; ntt_stage_k<false>(K, tile, tw, lb, w - rr * per_row, AddrContig{rr * rowpad}); }
            __syncthreads();
            done += K;
        }
    }
    PowTab pt{pw, pw + (1u << h), h, (1u << h) - 1};
    const uint32_t mmask = (1u << lg_m) - 1;
    for (uint32_t i = tid; i < rows_per_cta * Lc; i += nth) {
        const uint32_t rr = i >> logLc, k = i & (Lc - 1), R = row0 + rr;
        if (R < total_rows) {
            const uint32_t poly = R / rows_per_poly, rho = R - poly * rows_per_poly;
            uint32_t v = tile[AddrContig{rr * rowpad}(k)];
            if (pow_g) v = pt.apply(v, (k * bitrev(rho, lg_rows)) & mmask);
            else if (scale) v = fp_mul(v, scale);
            out[(size_t)poly * out_poly_stride + (size_t)rho * Lc + k] = v;
        }
    }
}


// =========================================================================================================
// Specialised (compile-time size) passes: all stage strides, twiddle offsets and loop counts are immediates,
// tables are read through the read-only path (L1-resident, no per-CTA staging), so a CTA only stages its data tile.
// =========================================================================================================
template <typename A> __device__ __forceinline__ uint32_t addr_base(const A& a, uint32_t pos) { return a.base(pos); }
__device__ __forceinline__ uint32_t addr_base(const AddrContig& a, uint32_t pos) { return a.base_of(pos); }

template <int K, int LB, bool DIF, bool TWS = false, typename ADDR>
__device__ __forceinline__ void ntt_stage_c(uint32_t* s, const tw_t* __restrict__ tw, uint32_t gidx, ADDR addr) {
    constexpr int R = 1 << K;
    constexpr uint32_t q = (1u << LB) >> K;
    const uint32_t b = gidx >> (LB - K), o = gidx & (q - 1);
    const uint32_t base = (b << LB) + o;
    uint32_t x[R];
#pragma unroll
    const uint32_t a0 = addr_base(addr, base);
#pragma unroll
    for (int j = 0; j < R; j++) x[j] = s[a0 + addr.off(j * q)];
    const tw_t* twp = tw + o;
#pragma unroll
    for (int ll = 0; ll < K; ll++) {
        const int l = DIF ? ll : (K - 1 - ll);
        const int half = R >> (l + 1);
#pragma unroll
        for (int j = 0; j < R; j++) {
            if ((j & half) == 0) {
                const tw_t* wp = twp + ((1u << LB) >> (l + 1)) + (j & (half - 1)) * q;
                const tw_t w = TWS ? *wp : __ldg(wp);      // TWS: table staged in shared memory by TMA
                const uint32_t a = x[j], bb = x[j + half];
                if (DIF) {
                    x[j] = nt_add(a, bb);
                    x[j + half] = mul_tw(nt_subp(a, bb), w);
                } else {
                    const uint32_t t = mul_tw(bb, w);
                    x[j] = nt_add(a, t);
                    x[j + half] = nt_sub(a, t);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < R; j++) s[a0 + addr.off(j * q)] = x[j];
}

// runs all stages of a length-2^LOGL transform held in shared memory; LOW = number of already-done low levels (DIT expand)
template <int LOGL, int LOW, bool DIF, int DONE = 0, bool TWS = false>
struct NttStages {
    template <typename MK>
    static __device__ __forceinline__ void run(uint32_t* s, const tw_t* tw, uint32_t tid, uint32_t nth, uint32_t nbatch_shift, MK mk) {
        constexpr int REM = LOGL - LOW - DONE;
        if constexpr (REM > 0) {
            // DIT takes the remainder stage first (next to the expand), DIF takes it last
            constexpr int K = DIF ? (REM >= 4 ? 4 : REM) : ((REM % 4) ? (REM % 4) : 4);
            constexpr int LB = DIF ? (LOGL - DONE) : (LOW + DONE + K);
            const uint32_t total = ((1u << LOGL) >> K) << nbatch_shift;
            for (uint32_t w = tid; w < total; w += nth) ntt_stage_c<K, LB, DIF, TWS>(s, tw, mk.gidx(w), mk.addr(w));
            __syncthreads();
            NttStages<LOGL, LOW, DIF, DONE + K, TWS>::run(s, tw, tid, nth, nbatch_shift, mk);
        }
    }
};

struct AddrStridedPad8 {   // 8-column tile, one padding row per 16 rows: groups of a warp that sit 16 rows apart hit different banks
    uint32_t t;
    __device__ __forceinline__ uint32_t operator()(uint32_t pos) const { return ((pos + (pos >> 4)) << 3) + t; }
    __device__ __forceinline__ uint32_t base(uint32_t pos) const { return ((pos + (pos >> 4)) << 3) + t; }
    __device__ __forceinline__ static constexpr uint32_t off(uint32_t d) { return (d + (d >> 4)) << 3; }
};
struct MkStridedPad8 {
    __device__ __forceinline__ uint32_t gidx(uint32_t w) const { return w >> 3; }
    __device__ __forceinline__ AddrStridedPad8 addr(uint32_t w) const { return AddrStridedPad8{w & 7u}; }
};
struct MkStrided {      // work item w -> (group index, column t)
    uint32_t lgTW;
    __device__ __forceinline__ uint32_t gidx(uint32_t w) const { return w >> lgTW; }
    __device__ __forceinline__ AddrStrided addr(uint32_t w) const { return AddrStrided{lgTW, w & ((1u << lgTW) - 1)}; }
};
template <int LOGL, bool DIF>
__global__ void __launch_bounds__(512) k_ntt_strided_c(uint32_t* __restrict__ data, uint32_t lgTW, uint32_t row_stride, uint32_t tiles_per_poly,
                                                       size_t poly_stride, const tw_t* __restrict__ tw_g,
                                                       const tw_t* __restrict__ pow_g, uint32_t lg_m) {
    extern __shared__ uint32_t smem[];
    constexpr uint32_t L = 1u << LOGL;
    const uint32_t TW = 1u << lgTW;
    uint32_t* tile = smem;
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    const uint32_t poly = blockIdx.x / tiles_per_poly, tcol = blockIdx.x % tiles_per_poly;
    uint32_t* g = data + (size_t)poly * poly_stride + ((size_t)tcol << lgTW);
    const uint32_t c = tid & (TW - 1), r0 = tid >> lgTW, rstep = nth >> lgTW;
    {
        const uint32_t* gp = g + (size_t)r0 * row_stride + c;
        const size_t gstep = (size_t)rstep * row_stride;
        uint32_t* tp = tile + (r0 << lgTW) + c;
#pragma unroll 16
        for (uint32_t r = r0; r < L; r += rstep) { *tp = __ldg(gp); gp += gstep; tp += rstep << lgTW; }
    }
    __syncthreads();
    NttStages<LOGL, 0, DIF>::run(tile, tw_g, tid, nth, lgTW, MkStrided{lgTW});
    if (pow_g) {
        const uint32_t h = (lg_m + 1) / 2;
        PowTab pt{pow_g, pow_g + (1u << h), h, (1u << h) - 1};
        const uint32_t mmask = (1u << lg_m) - 1, col = (tcol << lgTW) + c;
        uint32_t* gp = g + (size_t)r0 * row_stride + c;
        const size_t gstep = (size_t)rstep * row_stride;
        const uint32_t* tp = tile + (r0 << lgTW) + c;
        for (uint32_t r = r0; r < L; r += rstep) {
            const uint32_t e = (col * bitrev(r, LOGL)) & mmask;
            *gp = pow_apply(*tp, __ldg(pt.lo + (e & pt.mask)), __ldg(pt.hi + (e >> h)));
            gp += gstep; tp += rstep << lgTW;
        }
    } else {
        uint32_t* gp = g + (size_t)r0 * row_stride + c;
        const size_t gstep = (size_t)rstep * row_stride;
        const uint32_t* tp = tile + (r0 << lgTW) + c;
#pragma unroll 8
        for (uint32_t r = r0; r < L; r += rstep) { *gp = *tp; gp += gstep; tp += rstep << lgTW; }
    }
}

// Persistent, double-buffered strided pass: 8-column tiles (32-byte sectors), LDGSTS (cp.async) stages the NEXT tile into
// the second shared-memory buffer while the current one is transformed, so the HBM latency of a tile load is hidden
// behind the butterflies of the previous tile instead of stalling the CTA.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// TMA (cp.async.bulk, SASS UBLKCP) staging of a contiguous global table into shared memory, completion on an mbarrier.
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(mbar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
    const unsigned m = (unsigned)__cvta_generic_to_shared(mbar), d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gsrc), "r"(bytes), "r"(m) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    const unsigned m = (unsigned)__cvta_generic_to_shared(mbar);
    asm volatile("{\n\t.reg .pred p;\n\tB200_MBAR_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra B200_MBAR_DONE;\n\tbra B200_MBAR_WAIT;\n\tB200_MBAR_DONE:\n\t}" ::"r"(m), "r"(parity) : "memory");
}

template <int LOGL, bool DIF>
__global__ void __launch_bounds__(512) k_ntt_strided_p(uint32_t* __restrict__ data, uint32_t row_stride, uint32_t tiles_per_poly,
                                                       uint32_t num_tiles, size_t poly_stride, const tw_t* __restrict__ tw_g,
                                                       const tw_t* __restrict__ pow_g, uint32_t lg_m) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr uint32_t L = 1u << LOGL, TILE = (L + (L >> 4)) * 8;
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    const uint32_t h = (lg_m + 1) / 2, lmask = (1u << h) - 1, mmask = (1u << lg_m) - 1;
    const tw_t* plo = pow_g; const tw_t* phi = pow_g + (1u << h);
    // stage twiddles (L words) live behind the two tile buffers; one TMA bulk copy per CTA, awaited before the first stage
    tw_t* tw_s = reinterpret_cast<tw_t*>(smem + 2 * TILE);
    __shared__ __align__(8) uint64_t tw_bar;
    if (tid == 0) mbar_init(&tw_bar, 1);
    __syncthreads();
    if (tid == 0) tma_load_1d(tw_s, tw_g, L * 8, &tw_bar);
    auto tile_ptr = [&](uint32_t tile) { return data + (size_t)(tile / tiles_per_poly) * poly_stride + (size_t)(tile % tiles_per_poly) * 8; };
    auto issue_load = [&](uint32_t tile, uint32_t* buf) {
        const uint32_t* g = tile_ptr(tile);
        for (uint32_t ch = tid; ch < 2 * L; ch += nth) { const uint32_t r = ch >> 1; cp_async16(buf + (r + (r >> 4)) * 8 + (ch & 1) * 4, g + (size_t)r * row_stride + (ch & 1) * 4); }
        cp_async_commit();
    };
    uint32_t cur = 0;
    uint32_t tile = blockIdx.x;
    if (tile < num_tiles) issue_load(tile, smem);
    mbar_wait(&tw_bar, 0);
    for (; tile < num_tiles; tile += gridDim.x) {
        uint32_t* buf = smem + cur * TILE;
        const uint32_t next = tile + gridDim.x;
        if (next < num_tiles) { issue_load(next, smem + (cur ^ 1) * TILE); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        NttStages<LOGL, 0, DIF, 0, true>::run(buf, tw_s, tid, nth, 3, MkStridedPad8{});
        uint32_t* g = data + (size_t)(tile / tiles_per_poly) * poly_stride + (size_t)(tile % tiles_per_poly) * 8;
        const uint32_t col0 = (tile % tiles_per_poly) * 8;
        for (uint32_t ch = tid; ch < 2 * L; ch += nth) {
            const uint32_t r = ch >> 1, c4 = (ch & 1) * 4;
            uint4 v = *reinterpret_cast<const uint4*>(buf + (r + (r >> 4)) * 8 + c4);
            if (pow_g) {
                const uint32_t d1 = bitrev(r, LOGL);
                uint32_t e = ((col0 + c4) * d1) & mmask;
                v.x = pow_apply(v.x, __ldg(plo + (e & lmask)), __ldg(phi + (e >> h))); e = (e + d1) & mmask;
                v.y = pow_apply(v.y, __ldg(plo + (e & lmask)), __ldg(phi + (e >> h))); e = (e + d1) & mmask;
                v.z = pow_apply(v.z, __ldg(plo + (e & lmask)), __ldg(phi + (e >> h))); e = (e + d1) & mmask;
                v.w = pow_apply(v.w, __ldg(plo + (e & lmask)), __ldg(phi + (e >> h)));
            }
            *reinterpret_cast<uint4*>(g + (size_t)r * row_stride + c4) = v;
        }
        __syncthreads();       // everyone done with buf before the next iteration's prefetch overwrites it
        cur ^= 1;
    }
}

// contiguous rows (rows_per_poly is a power of two: poly = R >> lg_rpp)
template <int LOGLC, int LGE, bool DIF>
__global__ void __launch_bounds__(256) k_ntt_contig_c(uint32_t* out, const uint32_t* in, uint32_t rows_per_cta, uint32_t lg_rpp,
                                                      uint32_t total_rows, size_t in_poly_stride, size_t out_poly_stride,
                                                      const tw_t* __restrict__ tw_g, const tw_t* __restrict__ pow_g,
                                                      uint32_t lg_m, uint32_t lg_rows, uint32_t scale) {
    extern __shared__ uint32_t smem[];
    constexpr uint32_t Lc = 1u << LOGLC, Lin = Lc >> LGE, rowpad = Lc + (Lc >> 4);
    uint32_t* tile = smem;
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    const uint32_t row0 = blockIdx.x * rows_per_cta;
    const uint32_t rpp_mask = (1u << lg_rpp) - 1;
    for (uint32_t rr = 0; rr < rows_per_cta; rr++) {
        const uint32_t R = row0 + rr;
        if (R >= total_rows) break;
        const uint32_t* src = in + (size_t)(R >> lg_rpp) * in_poly_stride + (size_t)(R & rpp_mask) * Lin;
        uint32_t* trow = tile + rr * rowpad;
        for (uint32_t k = tid; k < Lin; k += nth) {
            const uint32_t v = src[k], pp = k << LGE, ph = pp + (pp >> 4);     // replicas stay inside one 16-word pad group
#pragma unroll
            for (uint32_t e = 0; e < (1u << LGE); e++) trow[ph + e] = v;
        }
    }
    __syncthreads();
    constexpr int NLEV = LOGLC - LGE;
    if constexpr (DIF) {
        constexpr int K1 = NLEV >= 4 ? 4 : NLEV, K2 = (NLEV - K1) >= 4 ? 4 : (NLEV - K1), K3 = (NLEV - K1 - K2) >= 4 ? 4 : (NLEV - K1 - K2),
                      K4 = NLEV - K1 - K2 - K3;
        static_assert(K4 <= 4, "row too long");
        if constexpr (K1 > 0) { for (uint32_t w = tid; w < rows_per_cta << (LOGLC - K1); w += nth) ntt_stage_c<K1, LOGLC, true>(tile, tw_g, w & ((1u << (LOGLC - K1)) - 1), AddrContig{(w >> (LOGLC - K1)) * rowpad}); __syncthreads(); }
        if constexpr (K2 > 0) { for (uint32_t w = tid; w < rows_per_cta << (LOGLC - K2); w += nth) ntt_stage_c<K2, LOGLC - K1, true>(tile, tw_g, w & ((1u << (LOGLC - K2)) - 1), AddrContig{(w >> (LOGLC - K2)) * rowpad}); __syncthreads(); }
        if constexpr (K3 > 0) { for (uint32_t w = tid; w < rows_per_cta << (LOGLC - K3); w += nth) ntt_stage_c<K3, LOGLC - K1 - K2, true>(tile, tw_g, w & ((1u << (LOGLC - K3)) - 1), AddrContig{(w >> (LOGLC - K3)) * rowpad}); __syncthreads(); }
        if constexpr (K4 > 0) { for (uint32_t w = tid; w < rows_per_cta << (LOGLC - K4); w += nth) ntt_stage_c<K4, LOGLC - K1 - K2 - K3, true>(tile, tw_g, w & ((1u << (LOGLC - K4)) - 1), AddrContig{(w >> (LOGLC - K4)) * rowpad}); __syncthreads(); }
    } else {
        constexpr int K1 = (NLEV % 4) ? (NLEV % 4) : (NLEV >= 4 ? 4 : 0), K2 = (NLEV - K1) >= 4 ? 4 : 0, K3 = (NLEV - K1 - K2) >= 4 ? 4 : 0,
                      K4 = NLEV - K1 - K2 - K3;
        static_assert(K4 == 0 || K4 == 4, "row too long");
        if constexpr (K1 > 0) { for (uint32_t w = tid; w < rows_per_cta << (LOGLC - K1); w += nth) ntt_stage_c<K1, LGE + K1, false>(tile, tw_g, w & ((1u << (LOGLC - K1)) - 1), AddrContig{(w >> (LOGLC - K1)) * rowpad}); __syncthreads(); }
        if constexpr (K2 > 0) { for (uint32_t w = tid; w < rows_per_cta << (LOGLC - K2); w += nth) ntt_stage_c<K2, LGE + K1 + K2, false>(tile, tw_g, w & ((1u << (LOGLC - K2)) - 1), AddrContig{(w >> (LOGLC - K2)) * rowpad}); __syncthreads(); }
        if constexpr (K3 > 0) { for (uint32_t w = tid; w < rows_per_cta << (LOGLC - K3); w += nth) ntt_stage_c<K3, LGE + K1 + K2 + K3, false>(tile, tw_g, w & ((1u << (LOGLC - K3)) - 1), AddrContig{(w >> (LOGLC - K3)) * rowpad}); __syncthreads(); }
        if constexpr (K4 > 0) { for (uint32_t w = tid; w < rows_per_cta << (LOGLC - K4); w += nth) ntt_stage_c<K4, LGE + K1 + K2 + K3 + K4, false>(tile, tw_g, w & ((1u << (LOGLC - K4)) - 1), AddrContig{(w >> (LOGLC - K4)) * rowpad}); __syncthreads(); }
    }
    const uint32_t h = (lg_m + 1) / 2;
    const tw_t* plo = pow_g; const tw_t* phi = pow_g + (1u << h);
    const uint32_t lmask = (1u << h) - 1, mmask = (1u << lg_m) - 1;
    for (uint32_t rr = 0; rr < rows_per_cta; rr++) {
        const uint32_t R = row0 + rr;
        if (R >= total_rows) break;
        const uint32_t rho = R & rpp_mask;
        uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(R >> lg_rpp) * out_poly_stride + (size_t)rho * Lc);
        const uint32_t* trow = tile + rr * rowpad;
        const uint32_t d1 = bitrev(rho, lg_rows);
        for (uint32_t k4 = tid; k4 < Lc / 4; k4 += nth) {
            const uint32_t k = k4 * 4, ph = k + (k >> 4);
            uint32_t v[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                v[i] = trow[ph + i];
                if (pow_g) { const uint32_t e = ((k + i) * d1) & mmask; v[i] = pow_apply(v[i], __ldg(plo + (e & lmask)), __ldg(phi + (e >> h))); }
                else if (scale) v[i] = fp_mul(v[i], scale);
            }
            dst[k4] = make_uint4(v[0], v[1], v[2], v[3]);
        }
    }
}


// =========================================================================================================
// Fused contiguous passes: the first transform stage reads straight from global memory (for the forward pass the x4
// `expand` and the two lowest remaining levels are done in registers from a 16-byte load of 4 coefficients), the last
// stage writes straight to global memory (with the inter-pass twiddle), so a row makes ONE shared-memory round trip
// per middle stage instead of load + every stage + store.
// =========================================================================================================
template <int K, int LB, bool DIF, bool TWS = false, typename LD, typename ST>
__device__ __forceinline__ void ntt_stage_io(const tw_t* __restrict__ tw, uint32_t gidx, LD ld, ST st) {
    constexpr int R = 1 << K;
    constexpr uint32_t q = (1u << LB) >> K;
    const uint32_t b = gidx >> (LB - K), o = gidx & (q - 1);
    const uint32_t base = (b << LB) + o;
    uint32_t x[R];
#pragma unroll
    for (int j = 0; j < R; j++) x[j] = ld(base, (uint32_t)(j * q));
    const tw_t* twp = tw + o;
#pragma unroll
    for (int ll = 0; ll < K; ll++) {
        const int l = DIF ? ll : (K - 1 - ll);
        const int half = R >> (l + 1);
#pragma unroll
        for (int j = 0; j < R; j++) {
            if ((j & half) == 0) {
                const tw_t* wp = twp + ((1u << LB) >> (l + 1)) + (j & (half - 1)) * q;
                const tw_t w = TWS ? *wp : __ldg(wp);
                const uint32_t a = x[j], bb = x[j + half];
                if (DIF) { x[j] = nt_add(a, bb); x[j + half] = mul_tw(nt_subp(a, bb), w); }
                else { const uint32_t t = mul_tw(bb, w); x[j] = nt_add(a, t); x[j + half] = nt_sub(a, t); }
            }
        }
    }
    st(base, x);
}

// Persistent strided pass, final stage fused with the global store: radix-16 stages run shared -> shared, the last
// (remainder) stage goes shared -> registers -> global (with the inter-pass twiddle for the inverse transform), so a tile
// makes one shared-memory round trip less and needs one barrier less than k_ntt_strided_p.
template <int LOGL, bool DIF>
__global__ void __launch_bounds__(512) k_ntt_strided_pf(uint32_t* __restrict__ data, uint32_t row_stride, uint32_t tiles_per_poly,
                                                        uint32_t num_tiles, size_t poly_stride, const tw_t* __restrict__ tw_g,
                                                        const tw_t* __restrict__ pow_g, uint32_t lg_m) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr uint32_t L = 1u << LOGL, TILE = (L + (L >> 4)) * 8;
    constexpr int KF = (LOGL % 4) ? (LOGL % 4) : 4, NFULL = (LOGL - KF) / 4;
    static_assert(NFULL >= 1 && NFULL <= 2, "unsupported tile height");
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    const uint32_t h = (lg_m + 1) / 2, lmask = (1u << h) - 1, mmask = (1u << lg_m) - 1;
    const tw_t* plo = pow_g; const tw_t* phi = pow_g + (1u << h);
    tw_t* tw_s = reinterpret_cast<tw_t*>(smem + 2 * TILE);
    __shared__ __align__(8) uint64_t tw_bar;
    if (tid == 0) mbar_init(&tw_bar, 1);
    __syncthreads();
    if (tid == 0) tma_load_1d(tw_s, tw_g, L * 8, &tw_bar);
    auto issue_load = [&](uint32_t tile, uint32_t* buf) {
        const uint32_t* g = data + (size_t)(tile / tiles_per_poly) * poly_stride + (size_t)(tile % tiles_per_poly) * 8;
        for (uint32_t ch = tid; ch < 2 * L; ch += nth) { const uint32_t r = ch >> 1; cp_async16(buf + (r + (r >> 4)) * 8 + (ch & 1) * 4, g + (size_t)r * row_stride + (ch & 1) * 4); }
        cp_async_commit();
    };
    uint32_t cur = 0, tile = blockIdx.x;
    if (tile < num_tiles) issue_load(tile, smem);
    mbar_wait(&tw_bar, 0);
    for (; tile < num_tiles; tile += gridDim.x) {
        uint32_t* buf = smem + cur * TILE;
        const uint32_t next = tile + gridDim.x;
        if (next < num_tiles) { issue_load(next, smem + (cur ^ 1) * TILE); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        {
            constexpr int LB1 = DIF ? LOGL : 4;
            for (uint32_t w = tid; w < (L >> 4) * 8; w += nth) ntt_stage_c<4, LB1, DIF, true>(buf, tw_s, w >> 3, AddrStridedPad8{w & 7u});
            __syncthreads();
        }
        if constexpr (NFULL >= 2) {
            constexpr int LB2 = DIF ? LOGL - 4 : 8;
            for (uint32_t w = tid; w < (L >> 4) * 8; w += nth) ntt_stage_c<4, LB2, DIF, true>(buf, tw_s, w >> 3, AddrStridedPad8{w & 7u});
            __syncthreads();
        }
        uint32_t* g = data + (size_t)(tile / tiles_per_poly) * poly_stride + (size_t)(tile % tiles_per_poly) * 8;
        const uint32_t col0 = (tile % tiles_per_poly) * 8;
        constexpr int LBF = DIF ? KF : LOGL;
        constexpr uint32_t QF = (1u << LBF) >> KF;
        for (uint32_t w = tid; w < (L >> KF) * 8; w += nth) {
            const uint32_t t = w & 7u;
            ntt_stage_io<KF, LBF, DIF, true>(tw_s, w >> 3,
                [&](uint32_t b0, uint32_t d) { return buf[((b0 + (b0 >> 4)) << 3) + t + ((d + (d >> 4)) << 3)]; },
                [&](uint32_t base, const uint32_t (&x)[1 << KF]) {
#pragma unroll
                    for (int j = 0; j < (1 << KF); j++) {
                        const uint32_t row = base + j * QF;
                        uint32_t v = x[j];
                        if (pow_g) { const uint32_t e = ((col0 + t) * bitrev(row, LOGL)) & mmask; v = pow_apply(v, __ldg(plo + (e & lmask)), __ldg(phi + (e >> h))); }
                        g[(size_t)row * row_stride + t] = v;
                    }
                });
        }
        __syncthreads();
        cur ^= 1;
    }
}

// =========================================================================================================
// Radix-32 register passes for the 1024-row strided tile (the hot shape: both 2^20 and 2^22 transforms split as 2^10 x rest).
// A 10-level transform is TWO register stages of 5 levels with ONE exchange through shared memory:
//   contiguous stage: a thread owns rows 32b..32b+31 of one column; its twiddles w_32^k are compile-time constants
//                     (immediate operands, and the 31 of 80 multiplications by w^0 = 1 are not emitted at all);
//   strided stage:    a thread owns rows t, t+32, ..., t+992; its 31 distinct twiddles come from the TMA-staged table.
// 256 threads = 8 columns x 32 row groups; one padding row per 32 rows makes both access patterns bank-conflict free
// (a warp touches 8 consecutive words at 4 row offsets that land 8 banks apart).  Against the radix-16 kernel
// (k_ntt_strided_pf: 4+4+2 levels) a tile makes one shared-memory round trip and one block barrier less and issues
// about a third fewer instructions.
// =========================================================================================================
__host__ __device__ constexpr uint32_t c_mulmod(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b % P); }
__host__ __device__ constexpr uint32_t c_powmod(uint32_t a, uint64_t e) {
    uint32_t r = 1;
    while (e) { if (e & 1) r = c_mulmod(r, a); a = c_mulmod(a, a); e >>= 1; }
    return r;
}
constexpr uint32_t W32_FWD = c_powmod(137u, 1ull << 22);      // ROU_FWD[5] = 137^(2^(27-5)), plain value
constexpr uint32_t W32_INV = c_powmod(W32_FWD, 31);
static_assert(c_powmod(W32_FWD, 16) == P - 1 && c_mulmod(W32_FWD, W32_INV) == 1, "w_32");
// w_32^k (k < 16) of the forward / inverse root as a Shoup pair, usable as immediates after unrolling
template <bool INV>
__host__ __device__ constexpr tw_t w32_pair(int k) {
    const uint32_t w = c_powmod(INV ? W32_INV : W32_FWD, (uint64_t)k);
    return tw_t{w, (uint32_t)(((uint64_t)w << 32) / P)};
}

// five butterfly levels on x[0..32): x[j] is the element at position o + j*q of its block.
// CONST: q = 1 (levels 1..5, twiddle of pair index i at sub-level m is w_32^(i * 2^(5-m)));
// otherwise q = 32 (levels 6..10, twiddle tws[(16 << m) + t + 32*i] from the compact per-level table).
template <bool DIF, bool CONST>
__device__ __forceinline__ void radix32_levels(uint32_t (&x)[32], const tw_t* __restrict__ tws, uint32_t t) {
#pragma unroll
    for (int mm = 0; mm < 5; mm++) {
        const int m = DIF ? 5 - mm : mm + 1;        // sub-level 1..5; DIF runs from the widest butterflies down
        const int half = 1 << (m - 1);
#pragma unroll
        for (int j = 0; j < 32; j++) {
            if ((j & half) == 0) {
                const int i = j & (half - 1);
                const bool one = CONST && i == 0;
                tw_t w;
                if (CONST) w = w32_pair<DIF>(i << (5 - m)); else w = tws[(16u << m) + t + 32u * i];
                const uint32_t a = x[j], b = x[j + half];
                if (DIF) {
                    x[j] = nt_add(a, b);
                    x[j + half] = one ? nt_sub(a, b) : mul_tw(nt_subp(a, b), w);     // a - b + p in (0, 2p): mul_tw takes any u32
                } else {
                    const uint32_t tt = one ? b : mul_tw(b, w);
                    x[j] = nt_add(a, tt);
                    x[j + half] = nt_sub(a, tt);
                }
            }
        }
    }
}

// LGRS >= 0: row_stride == 2^LGRS is a compile-time constant, so every global address of a tile is base + immediate.
// POW: the inverse inter-pass twiddle in the epilogue (only when pass B is not the fused kernel that applies it on load).
template <bool DIF, int MINB, int LGRS, bool POW>
__global__ void __launch_bounds__(256, MINB) k_ntt_strided_r32(uint32_t* __restrict__ data, uint32_t row_stride_rt, uint32_t tiles_per_poly,
                                                                uint32_t num_tiles, size_t poly_stride, const tw_t* __restrict__ tw_g,
                                                                const tw_t* __restrict__ pow_g, uint32_t lg_m) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr uint32_t L = 1024, TILE = (L + (L >> 5)) * 8;
    const uint32_t tid = threadIdx.x;
    const size_t row_stride = LGRS >= 0 ? ((size_t)1 << (LGRS >= 0 ? LGRS : 0)) : (size_t)row_stride_rt;
    const uint32_t h = (lg_m + 1) / 2, lmask = (1u << h) - 1, mmask = (1u << lg_m) - 1;
    const tw_t* plo = pow_g; const tw_t* phi = pow_g + (1u << h);
    tw_t* tw_s = reinterpret_cast<tw_t*>(smem + 2 * TILE);
    __shared__ __align__(8) uint64_t tw_bar;
    if (tid == 0) mbar_init(&tw_bar, 1);
    __syncthreads();
    if (tid == 0) tma_load_1d(tw_s, tw_g, L * 8, &tw_bar);
    // chunk i of thread tid: row (tid >> 1) + 128 i, half (tid & 1): 128 rows = 132 padded rows apart in shared memory
    const uint32_t ld_r = tid >> 1, ld_soff = (ld_r + (ld_r >> 5)) * 8 + (tid & 1) * 4;
    auto issue_load = [&](uint32_t tile, uint32_t* buf) {
        const uint32_t* g = data + (size_t)(tile / tiles_per_poly) * poly_stride + (size_t)(tile % tiles_per_poly) * 8 + ld_r * row_stride + (tid & 1) * 4;
#pragma unroll
        for (int i = 0; i < 8; i++) cp_async16(buf + ld_soff + i * 132 * 8, g + (size_t)(128 * i) * row_stride);
        cp_async_commit();
    };
    const uint32_t c = tid & 7u, grp = tid >> 3;                 // column of the tile, row group 0..31
    // contiguous stage: rows 32*grp + j  -> word (33*grp + j)*8 + c;   strided stage: rows grp + 32*k -> word (grp + 33*k)*8 + c
    const uint32_t off_contig = (33u * grp) * 8u + c, off_strided = grp * 8u + c;
    uint32_t cur = 0, tile = blockIdx.x;
    if (tile < num_tiles) issue_load(tile, smem);
    mbar_wait(&tw_bar, 0);
    for (; tile < num_tiles; tile += gridDim.x) {
        uint32_t* buf = smem + cur * TILE;
        const uint32_t next = tile + gridDim.x;
        if (next < num_tiles) { issue_load(next, smem + (cur ^ 1) * TILE); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        uint32_t* g = data + (size_t)(tile / tiles_per_poly) * poly_stride + (size_t)(tile % tiles_per_poly) * 8 + c +
                      (size_t)(DIF ? 32u * grp : grp) * row_stride;      // first row this thread stores
        uint32_t x[32];
        if constexpr (!DIF) {
            // levels 1..5 on 32 consecutive rows (constant twiddles), back to shared memory
            uint32_t* p1 = buf + off_contig;
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] = p1[j * 8];
            radix32_levels<false, true>(x, nullptr, 0u);
#pragma unroll
            for (int j = 0; j < 32; j++) p1[j * 8] = x[j];
            __syncthreads();
            // levels 6..10 on rows grp + 32k, straight to global memory
            const uint32_t* p2 = buf + off_strided;
#pragma unroll
            for (int k = 0; k < 32; k++) x[k] = p2[k * 33 * 8];
            radix32_levels<false, false>(x, tw_s, grp);
#pragma unroll
            for (int k = 0; k < 32; k++) g[(size_t)(32 * k) * row_stride] = x[k];
        } else {
            // levels 10..6 on rows grp + 32k, back to shared memory
            uint32_t* p2 = buf + off_strided;
#pragma unroll
            for (int k = 0; k < 32; k++) x[k] = p2[k * 33 * 8];
            radix32_levels<true, false>(x, tw_s, grp);
#pragma unroll
            for (int k = 0; k < 32; k++) p2[k * 33 * 8] = x[k];
            __syncthreads();
            // levels 5..1 on 32 consecutive rows (constant twiddles), inter-pass twiddle, straight to global memory
            const uint32_t* p1 = buf + off_contig;
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] = p1[j * 8];
            radix32_levels<true, true>(x, nullptr, 0u);
            // exponent col * bitrev10(32 grp + j) = col * bitrev5(grp) + (col << 5) * bitrev5(j): two per-tile values and an immediate
            const uint32_t col = (tile % tiles_per_poly) * 8 + c;
            const uint32_t ea = col * bitrev(grp, 5), eb = col << 5;
            if constexpr (POW) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    constexpr uint32_t BR5[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                                  1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};
                    const uint32_t e = (ea + eb * BR5[j]) & mmask;
                    g[(size_t)j * row_stride] = pow_apply(x[j], __ldg(plo + (e & lmask)), __ldg(phi + (e >> h)));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) g[(size_t)j * row_stride] = x[j];
            }
        }
        __syncthreads();       // every thread is done with buf before the next iteration's prefetch overwrites it
        cur ^= 1;
    }
}

// The same two radix-32 stages with the tile loaded straight into registers (no cp.async staging, no double buffer): the only shared
// memory is the exchange buffer between the stages (33 KB) and the stage table, so 4 CTAs fit an SM (against 3 with the staged tile)
// and the loads of one CTA are hidden by the arithmetic of the other three.  A warp's 32-bit accesses cover 4 rows x 32 bytes.
template <bool DIF, int LGRS, int MINB>
__global__ void __launch_bounds__(256, MINB) k_ntt_strided_r32d(uint32_t* __restrict__ data, uint32_t row_stride_rt, uint32_t tiles_per_poly,
                                                             uint32_t num_tiles, size_t poly_stride, const tw_t* __restrict__ tw_g) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr uint32_t L = 1024, TILE = (L + (L >> 5)) * 8;
    const uint32_t tid = threadIdx.x;
    const size_t row_stride = LGRS >= 0 ? ((size_t)1 << (LGRS >= 0 ? LGRS : 0)) : (size_t)row_stride_rt;
    tw_t* tw_s = reinterpret_cast<tw_t*>(smem + TILE);
    __shared__ __align__(8) uint64_t tw_bar;
    if (tid == 0) mbar_init(&tw_bar, 1);
    __syncthreads();
    if (tid == 0) tma_load_1d(tw_s, tw_g, L * 8, &tw_bar);
    const uint32_t c = tid & 7u, grp = tid >> 3;
    uint32_t* p1 = smem + (33u * grp) * 8u + c;        // rows 32 grp + j
    uint32_t* p2 = smem + grp * 8u + c;                // rows grp + 32 k
    bool first = true;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        uint32_t* g0 = data + (size_t)(tile / tiles_per_poly) * poly_stride + (size_t)(tile % tiles_per_poly) * 8 + c;
        uint32_t* gc = g0 + (size_t)(32u * grp) * row_stride;
        uint32_t* gs = g0 + (size_t)grp * row_stride;
        uint32_t x[32];
        if constexpr (!DIF) {
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] = gc[(size_t)j * row_stride];
            radix32_levels<false, true>(x, nullptr, 0u);
#pragma unroll
            for (int j = 0; j < 32; j++) p1[j * 8] = x[j];
            if (first) { mbar_wait(&tw_bar, 0); first = false; }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 32; k++) x[k] = p2[k * 33 * 8];
            radix32_levels<false, false>(x, tw_s, grp);
#pragma unroll
            for (int k = 0; k < 32; k++) gs[(size_t)(32 * k) * row_stride] = x[k];
        } else {
#pragma unroll
            for (int k = 0; k < 32; k++) x[k] = gs[(size_t)(32 * k) * row_stride];
            if (first) { mbar_wait(&tw_bar, 0); first = false; }
            radix32_levels<true, false>(x, tw_s, grp);
#pragma unroll
            for (int k = 0; k < 32; k++) p2[k * 33 * 8] = x[k];
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] = p1[j * 8];
            radix32_levels<true, true>(x, nullptr, 0u);
#pragma unroll
            for (int j = 0; j < 32; j++) gc[(size_t)j * row_stride] = x[j];
        }
        __syncthreads();       // the exchange buffer is free again
    }
}

// forward pass 1: rows of Lin = Lc/4 bit-reversed coefficients -> Lc values (levels 3..LOGLC of the size-Lc DIT), times
// w_M^(k * d1).  One work item of the head = 4 coefficients -> 16 consecutive positions (levels 3 and 4 in registers).
// MINB: CTAs per SM the register allocation aims at.  The expanding form fits 40 registers (6 CTAs) without spills and is 6 % faster
// there than at its natural 48 (profiles/ntt_fwd1_minb_r02.txt); the plain form needs 59 (4 CTAs).
template <int LOGLC, int LGE, int MINB = (LGE == 2 ? 6 : 4)>
__global__ void __launch_bounds__(256, MINB) k_ntt_fwd1(uint32_t* out, const uint32_t* in, uint32_t rows_per_cta, uint32_t lg_rpp,
                                                  uint32_t total_rows, size_t in_poly_stride, size_t out_poly_stride,
                                                  const tw_t* __restrict__ tw_g, const tw_t* __restrict__ pow_g, uint32_t lg_m,
                                                  uint32_t lg_rows, const tw_t* __restrict__ tw_full) {
    extern __shared__ uint32_t smem[];
    static_assert(LGE == 0 || LGE == 2, "blow-up 1 or 4");
    constexpr uint32_t Lc = 1u << LOGLC, Lin = Lc >> LGE, rowpad = Lc + (Lc >> 4);
    constexpr int NREM = LOGLC - 4;                 // levels 5..LOGLC
    static_assert(NREM >= 4, "row too short for the fused kernel");
    uint32_t* tile = smem;
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    const uint32_t row0 = blockIdx.x * rows_per_cta, rpp_mask = (1u << lg_rpp) - 1;
    uint32_t nrows = total_rows - row0; if (nrows > rows_per_cta) nrows = rows_per_cta;
    // ---- head: levels 1..4 of 16 consecutive positions, in registers ----
    if constexpr (LGE == 0) {
        for (uint32_t w = tid; w < nrows << (LOGLC - 4); w += nth) {
            const uint32_t rr = w >> (LOGLC - 4), t = w & ((Lc >> 4) - 1), R = row0 + rr;
            const uint4* src = reinterpret_cast<const uint4*>(in + (size_t)(R >> lg_rpp) * in_poly_stride + (size_t)(R & rpp_mask) * Lin + 16 * t);
            const uint4 c0 = src[0], c1 = src[1], c2 = src[2], c3 = src[3];
            const uint32_t c[16] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w, c3.x, c3.y, c3.z, c3.w};
            uint32_t* dst = tile + rr * rowpad + 17 * t;
            ntt_stage_io<4, 4, false>(tw_g, 0u, [&](uint32_t, uint32_t d) { return c[d]; },
                                      [&](uint32_t, const uint32_t (&x)[16]) {
#pragma unroll
                                          for (int j = 0; j < 16; j++) dst[j] = x[j];
                                      });
        }
    } else {   // expand x4 + levels 3,4
        const tw_t w8_1 = __ldg(tw_g + 5), w8_2 = __ldg(tw_g + 6), w8_3 = __ldg(tw_g + 7);       // w_8^i  at tw[4+i]
        tw_t w16[8];
#pragma unroll
        for (int i = 0; i < 8; i++) w16[i] = __ldg(tw_g + 8 + i);                                       // w_16^i at tw[8+i]
        for (uint32_t w = tid; w < nrows << (LOGLC - 4); w += nth) {
            const uint32_t rr = w >> (LOGLC - 4), t = w & ((Lc >> 4) - 1), R = row0 + rr;
            const uint4 c = *reinterpret_cast<const uint4*>(in + (size_t)(R >> lg_rpp) * in_poly_stride + (size_t)(R & rpp_mask) * Lin + 4 * t);
            uint32_t y[16];
            {   // level 3 on (c.x, c.y) -> y[0..7], on (c.z, c.w) -> y[8..15]
                const uint32_t b1 = mul_tw(c.y, w8_1), b2 = mul_tw(c.y, w8_2), b3 = mul_tw(c.y, w8_3);
                y[0] = nt_add(c.x, c.y); y[4] = nt_sub(c.x, c.y);
                y[1] = nt_add(c.x, b1);  y[5] = nt_sub(c.x, b1);
                y[2] = nt_add(c.x, b2);  y[6] = nt_sub(c.x, b2);
                y[3] = nt_add(c.x, b3);  y[7] = nt_sub(c.x, b3);
                const uint32_t d1 = mul_tw(c.w, w8_1), d2 = mul_tw(c.w, w8_2), d3 = mul_tw(c.w, w8_3);
                y[8] = nt_add(c.z, c.w);  y[12] = nt_sub(c.z, c.w);
                y[9] = nt_add(c.z, d1);   y[13] = nt_sub(c.z, d1);
                y[10] = nt_add(c.z, d2);  y[14] = nt_sub(c.z, d2);
                y[11] = nt_add(c.z, d3);  y[15] = nt_sub(c.z, d3);
            }
            uint32_t* dst = tile + rr * rowpad + 17 * t;        // phys(16t + j) = 16t + j + t
#pragma unroll
            for (int i = 0; i < 8; i++) {                        // level 4
                const uint32_t bb = i == 0 ? y[8] : mul_tw(y[8 + i], w16[i]);
                dst[i] = nt_add(y[i], bb);
                dst[8 + i] = nt_sub(y[i], bb);
            }
        }
    }
    __syncthreads();
    // ---- middle stages (shared -> shared): remainder first, then radix-16 stages, leaving one radix-16 stage for the output ----
    constexpr int KA = NREM % 4;                                  // 0..3
    constexpr int NMID4 = (NREM - KA) / 4 - 1;                    // number of full middle stages
    if constexpr (KA > 0) {
        for (uint32_t w = tid; w < nrows << (LOGLC - KA); w += nth)
            ntt_stage_c<KA, 4 + KA, false>(tile, tw_g, w & ((1u << (LOGLC - KA)) - 1), AddrContig{(w >> (LOGLC - KA)) * rowpad});
        __syncthreads();
    }
    if constexpr (NMID4 >= 1) {
        for (uint32_t w = tid; w < nrows << (LOGLC - 4); w += nth)
            ntt_stage_c<4, 4 + KA + 4, false>(tile, tw_g, w & ((1u << (LOGLC - 4)) - 1), AddrContig{(w >> (LOGLC - 4)) * rowpad});
        __syncthreads();
    }
    static_assert(NMID4 <= 1, "row too long for the fused kernel");
    // ---- last stage (levels LOGLC-3..LOGLC) straight to global, times the inter-pass twiddle ----
    const uint32_t h = (lg_m + 1) / 2, lmask = (1u << h) - 1, mmask = (1u << lg_m) - 1;
    const tw_t* plo = pow_g; const tw_t* phi = pow_g + (1u << h);
    for (uint32_t w = tid; w < nrows << (LOGLC - 4); w += nth) {
        const uint32_t rr = w >> (LOGLC - 4), R = row0 + rr, rho = R & rpp_mask;
        const uint32_t* trow = tile + rr * rowpad;
        uint32_t* orow = out + (size_t)(R >> lg_rpp) * out_poly_stride + (size_t)rho * Lc;
        const uint32_t d1 = pow_g ? bitrev(rho, lg_rows) : 0u;
        ntt_stage_io<4, LOGLC, false>(tw_g, w & ((1u << (LOGLC - 4)) - 1),
            [&](uint32_t b0, uint32_t d) { return trow[(b0 + (b0 >> 4)) + (d + (d >> 4))]; },
            [&](uint32_t base, const uint32_t (&x)[16]) {
                constexpr uint32_t q = Lc >> 4;
                if (tw_full) {      // inter-pass twiddle from the table in data layout (get_full_table)
                    const tw_t* tf = tw_full + (size_t)rho * Lc + base;
#pragma unroll
                    for (int j = 0; j < 16; j++) orow[base + j * q] = mul_tw(x[j], __ldg(tf + j * q));
                    return;
                }
                uint32_t e = (base * d1) & mmask;
                const uint32_t estep = (q * d1) & mmask;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    uint32_t v = x[j];
                    if (pow_g) { v = pow_apply(v, __ldg(plo + (e & lmask)), __ldg(phi + (e >> h))); e = (e + estep) & mmask; }
                    orow[base + j * q] = v;
                }
            });
    }
}

// inverse pass B: contiguous rows of Lc values, DIF levels LOGLC..1, in place (out == in allowed), optional scale.
// pow_g != nullptr: the six-step inter-pass twiddle w_N^-(pos * bitrev(row)) (times the 1/N folded into its table) is applied HERE, to
// the values as they are loaded, instead of in the epilogue of the strided pass A: there it costs the radix-32 kernel 64 table loads in
// flight per thread (125 registers, 2 CTAs per SM); here it is 2 multiplies on the way into a kernel that runs at 70 % pipe use.
template <int LOGLC>
__global__ void __launch_bounds__(256) k_ntt_invb(uint32_t* out, const uint32_t* in, uint32_t rows_per_cta, uint32_t lg_rpp, uint32_t total_rows,
                                                  size_t in_poly_stride, size_t out_poly_stride, const tw_t* __restrict__ tw_g,
                                                  uint32_t scale, const tw_t* __restrict__ p3lo, const tw_t* __restrict__ p3hi,
                                                  const tw_t* __restrict__ pow_g, uint32_t lg_m,
                                                  const tw_t* __restrict__ tw_full, const tw_t* __restrict__ zk_full) {
    extern __shared__ uint32_t smem[];
    constexpr uint32_t Lc = 1u << LOGLC, rowpad = Lc + (Lc >> 4);
    constexpr int REM = LOGLC - 4;                   // levels after the first radix-16 stage
    constexpr int KF = (REM % 4) ? (REM % 4) : 4;    // final stage radix (1..4), written as vectors of 2^KF consecutive words
    constexpr int NMID4 = (REM - KF) / 4;
    static_assert(REM >= 1 && NMID4 <= 2, "unsupported row length");
    uint32_t* tile = smem;
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    const uint32_t row0 = blockIdx.x * rows_per_cta, rpp_mask = (1u << lg_rpp) - 1;
    uint32_t nrows = total_rows - row0; if (nrows > rows_per_cta) nrows = rows_per_cta;
    // ---- first stage: global -> registers -> shared ----
    const uint32_t hh = (lg_m + 1) / 2, lmask = (1u << hh) - 1, mmask = (1u << lg_m) - 1;
    const tw_t* plo = pow_g; const tw_t* phi = pow_g + (1u << hh);
    for (uint32_t w = tid; w < nrows << (LOGLC - 4); w += nth) {
        const uint32_t rr = w >> (LOGLC - 4), R = row0 + rr;
        const uint32_t* irow = in + (size_t)(R >> lg_rpp) * in_poly_stride + (size_t)(R & rpp_mask) * Lc;
        uint32_t* trow = tile + rr * rowpad;
        const uint32_t d1 = pow_g ? bitrev(R & rpp_mask, lg_rpp) : 0u;
        const tw_t* tf = tw_full ? tw_full + (size_t)(R & rpp_mask) * Lc : nullptr;
        ntt_stage_io<4, LOGLC, true>(tw_g, w & ((1u << (LOGLC - 4)) - 1),
            [&](uint32_t b0, uint32_t d) {
                uint32_t v = irow[b0 + d];
                if (tw_full) v = mul_tw(v, __ldg(tf + b0 + d));       // per-element table in data layout (get_full_table)
                else if (pow_g) {
                    const uint32_t e = ((b0 + d) * d1) & mmask;
                    v = pow_apply(v, __ldg(plo + (e & lmask)), __ldg(phi + (e >> hh)));
                }
                return v;
            },
            [&](uint32_t base, const uint32_t (&x)[16]) {
                constexpr uint32_t q = Lc >> 4;
                const uint32_t pb = base + (base >> 4);
#pragma unroll
                for (int j = 0; j < 16; j++) trow[pb + AddrContig::off(j * q)] = x[j];
            });
    }
    __syncthreads();
    if constexpr (NMID4 >= 1) {
        for (uint32_t w = tid; w < nrows << (LOGLC - 4); w += nth)
            ntt_stage_c<4, LOGLC - 4, true>(tile, tw_g, w & ((1u << (LOGLC - 4)) - 1), AddrContig{(w >> (LOGLC - 4)) * rowpad});
        __syncthreads();
    }
    if constexpr (NMID4 >= 2) {
        for (uint32_t w = tid; w < nrows << (LOGLC - 4); w += nth)
            ntt_stage_c<4, LOGLC - 8, true>(tile, tw_g, w & ((1u << (LOGLC - 4)) - 1), AddrContig{(w >> (LOGLC - 4)) * rowpad});
        __syncthreads();
    }
    // ---- final stage: shared -> registers -> global, 2^KF consecutive words per work item ----
    for (uint32_t w = tid; w < nrows << (LOGLC - KF); w += nth) {
        const uint32_t rr = w >> (LOGLC - KF), R = row0 + rr;
        const uint32_t* trow = tile + rr * rowpad;
        uint32_t* orow = out + (size_t)(R >> lg_rpp) * out_poly_stride + (size_t)(R & rpp_mask) * Lc;
        ntt_stage_io<KF, KF, true>(tw_g, w & ((1u << (LOGLC - KF)) - 1),
            [&](uint32_t b0, uint32_t d) { return trow[(b0 + (b0 >> 4)) + (d + (d >> 4))]; },
            [&](uint32_t base, const uint32_t (&x)[1 << KF]) {
                uint32_t v[1 << KF];
#pragma unroll
                for (int j = 0; j < (1 << KF); j++) {
                    v[j] = scale ? fp_mul(x[j], scale) : x[j];
                    if (p3lo && zk_full) v[j] = mul_tw(v[j], __ldg(zk_full + (size_t)(R & rpp_mask) * Lc + base + j));
                    else if (p3lo) {   // fused zk_shift (K2): slot holds degree d = bitrev(row) + rows_per_poly * bitrev(pos); multiply by 3^d
                        const uint32_t d = bitrev(R & rpp_mask, lg_rpp) + (bitrev(base + j, LOGLC) << lg_rpp);
                        v[j] = pow_apply(v[j], __ldg(p3lo + (d & 4095)), __ldg(p3hi + (d >> 12)));
                    }
                }
                if constexpr (KF == 1) *reinterpret_cast<uint2*>(orow + base) = make_uint2(v[0], v[1]);
                else {
#pragma unroll
                    for (int j = 0; j < (1 << KF); j += 4) *reinterpret_cast<uint4*>(orow + base + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            });
    }
}

// inverse pass B for rows of 1024 values (the 2^20 transform of the segment's columns) in the radix-32 register structure: ONE WARP per
// row.  Lane t loads elements t + 32k (coalesced, with the inter-pass twiddle), runs levels 10..6, the 32 x 32 exchange goes through the
// warp's own 1056 words of shared memory (element i at i + i/32: both access patterns conflict-free), levels 5..1 run on elements
// 32t..32t+31 with immediate twiddles, and the result goes back through the same words so that the stores are 16-byte and coalesced.
// No block barrier anywhere: the warps of a CTA only share the TMA-staged stage table.
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_ntt_invb_r32(uint32_t* out, const uint32_t* in, uint32_t lg_rpp, uint32_t total_rows,
                                                         size_t in_poly_stride, size_t out_poly_stride, const tw_t* __restrict__ tw_g,
                                                         uint32_t scale, const tw_t* __restrict__ p3lo, const tw_t* __restrict__ p3hi,
                                                         const tw_t* __restrict__ pow_g, uint32_t lg_m,
                                                         const tw_t* __restrict__ tw_full, const tw_t* __restrict__ zk_full) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr uint32_t Lc = 1024, ROWW = Lc + (Lc >> 5);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    tw_t* tw_s = reinterpret_cast<tw_t*>(smem + 8 * ROWW);
    __shared__ __align__(8) uint64_t tw_bar;
    if (tid == 0) mbar_init(&tw_bar, 1);
    __syncthreads();
    if (tid == 0) tma_load_1d(tw_s, tw_g, Lc * 8, &tw_bar);
    uint32_t* buf = smem + warp * ROWW;
    const uint32_t hh = (lg_m + 1) / 2, lmask = (1u << hh) - 1, mmask = (1u << lg_m) - 1, rpp_mask = (1u << lg_rpp) - 1;
    const tw_t* plo = pow_g; const tw_t* phi = pow_g + (1u << hh);
    const uint32_t brl = bitrev(lane, 5);
    bool first = true;
    for (uint32_t R = blockIdx.x * 8 + warp; R < total_rows; R += gridDim.x * 8) {
        const uint32_t rho = R & rpp_mask, d1 = bitrev(rho, lg_rpp);
        const uint32_t* irow = in + (size_t)(R >> lg_rpp) * in_poly_stride + (size_t)rho * Lc + lane;
        uint32_t x[32];
#pragma unroll
        for (int k = 0; k < 32; k++) x[k] = irow[32 * k];
        if (tw_full) {        // inter-pass twiddle from the table in data layout (get_full_table): one coalesced load, one multiply
            const tw_t* tf = tw_full + (size_t)rho * Lc + lane;
#pragma unroll
            for (int k = 0; k < 32; k++) x[k] = mul_tw(x[k], __ldg(tf + 32 * k));
        } else if (pow_g) {
            uint32_t e = (lane * d1) & mmask;
            const uint32_t estep = (32u * d1) & mmask;
#pragma unroll
            for (int k = 0; k < 32; k++) { x[k] = pow_apply(x[k], __ldg(plo + (e & lmask)), __ldg(phi + (e >> hh))); e = (e + estep) & mmask; }
        }
        if (first) { mbar_wait(&tw_bar, 0); first = false; }
        radix32_levels<true, false>(x, tw_s, lane);
#pragma unroll
        for (int k = 0; k < 32; k++) buf[lane + 33 * k] = x[k];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j++) x[j] = buf[33 * lane + j];
        radix32_levels<true, true>(x, nullptr, 0u);
        if (scale) {
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] = fp_mul(x[j], scale);
        }
        if (p3lo && !zk_full) {   // fused zk_shift: slot 32 t + j holds degree bitrev(row) + rows_per_poly * bitrev10(32 t + j); multiply by 3^d
            constexpr uint32_t BR5[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                          1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};
            const uint32_t da = d1 + (brl << lg_rpp);
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const uint32_t d = da + ((BR5[j] << 5) << lg_rpp);
                x[j] = pow_apply(x[j], __ldg(p3lo + (d & 4095)), __ldg(p3hi + (d >> 12)));
            }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j++) buf[33 * lane + j] = x[j];
        __syncwarp();
        uint32_t* orow = out + (size_t)(R >> lg_rpp) * out_poly_stride + (size_t)rho * Lc;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t idx = i * 128 + lane * 4, pa = idx + (idx >> 5);
            uint4 v = make_uint4(buf[pa], buf[pa + 1], buf[pa + 2], buf[pa + 3]);
            if (p3lo && zk_full) {      // fused zk_shift from the table in data layout: four factors = two 16-byte loads
                const uint4* zf = reinterpret_cast<const uint4*>(zk_full + (size_t)rho * Lc + idx);
                const uint4 z0 = __ldg(zf), z1 = __ldg(zf + 1);
                v.x = mul_tw(v.x, make_uint2(z0.x, z0.y)); v.y = mul_tw(v.y, make_uint2(z0.z, z0.w));
                v.z = mul_tw(v.z, make_uint2(z1.x, z1.y)); v.w = mul_tw(v.w, make_uint2(z1.z, z1.w));
            }
            *reinterpret_cast<uint4*>(orow + idx) = v;
        }
        __syncwarp();
    }
}

__global__ void k_zk_shift(uint32_t* __restrict__ io, uint32_t lg_n, size_t total, const tw_t* __restrict__ p3lo,
                           const tw_t* __restrict__ p3hi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t d = bitrev((uint32_t)(i & ((1u << lg_n) - 1)), lg_n);
    io[i] = pow_apply(io[i], p3lo[d & 4095], p3hi[d >> 12]);
}

// Inter-pass twiddle of the three-pass forward transform (sizes 2^25, 2^26): element k of row R (rho = R mod 2^lg_rows) is multiplied
// by w_M^(k * bitrev(rho)).  In the two-pass route this multiply is fused into the contiguous pass (k_ntt_fwd1); here the contiguous
// "pass" is itself a two-pass transform, so it is a kernel of its own (one extra trip through HBM, only for these two sizes).
__global__ void k_ntt_row_twiddle(uint32_t* __restrict__ io, uint32_t lg_row, uint32_t lg_rows, size_t total, const tw_t* __restrict__ pow_g,
                                  uint32_t lg_m) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= total) return;
    const uint32_t h = (lg_m + 1) / 2, lmask = (1u << h) - 1, mmask = (1u << lg_m) - 1;
    const tw_t* plo = pow_g; const tw_t* phi = pow_g + (1u << h);
    const uint32_t k = (uint32_t)(i & (((size_t)1 << lg_row) - 1));
    const uint32_t rho = (uint32_t)((i >> lg_row) & ((1u << lg_rows) - 1));
    const uint32_t d1 = bitrev(rho, lg_rows);
    uint4 v = *reinterpret_cast<uint4*>(io + i);
    uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t e = (uint32_t)(((uint64_t)(k + j) * d1) & mmask);
        x[j] = pow_apply(x[j], __ldg(plo + (e & lmask)), __ldg(phi + (e >> h)));
    }
    *reinterpret_cast<uint4*>(io + i) = make_uint4(x[0], x[1], x[2], x[3]);
}

__global__ void k_bit_reverse(uint32_t* __restrict__ io, uint32_t lg_n, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t j = (uint32_t)(i & ((1u << lg_n) - 1));
    const uint32_t r = bitrev(j, lg_n);
    if (j < r) { size_t b = i - j; uint32_t t = io[b + j]; io[b + j] = io[b + r]; io[b + r] = t; }
}

// ---- host launchers ------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
static uint32_t strided_lgTW(uint32_t logL) {
    const int ov = env_int("B200_NTT_LGTW", -1);        // tuning override (tools/time_ntt.py)
    if (ov >= 2 && ov <= 5) return (uint32_t)ov;
    // 32 KB tiles (4-6 CTAs of 256 threads per SM hide the tile-load latency; measured best on B200 with
    // tools/time_ntt.py), at least 8 columns (32-byte sectors), at most 32
    uint32_t lg = logL >= 13 ? 3 : (13 - logL);
    if (lg > 5) lg = 5;
    if (lg < 3) lg = 3;
    return lg;
}
// N = N1 * N2 with the strided pass over N1 rows.  2^10 rows is the shape the radix-32 register kernel covers, so it is preferred
// whenever the contiguous pass then has rows of 2^8 .. 2^13 (sizes 2^18 .. 2^23); otherwise the balanced split.
static uint32_t split_n1(uint32_t lg, uint32_t lg_e = 0) {
    if (lg >= 18 && lg + lg_e <= 23 && env_int("B200_NTT_SPLIT10", 1)) return 10;
    uint32_t n1 = lg / 2;
    return n1 > 11 ? 11 : n1;
}

template <bool DIF>
static cudaError_t run_strided(const DeviceTables* T, uint32_t* d, uint32_t logL, uint32_t row_stride, uint32_t ncols, uint32_t count,
                               size_t poly_stride, const tw_t* pow_g, uint32_t lg_m, cudaStream_t s) {
    uint32_t lgTW = strided_lgTW(logL);
    while ((1u << lgTW) > ncols) lgTW--;
    const uint32_t tiles_per_poly = ncols >> lgTW;
    const uint32_t h = (lg_m + 1) / 2;
    size_t smem = ((size_t)(1u << logL) << lgTW) * 4 + ((size_t)8 << logL) + (pow_g ? ((size_t)8 << h) + ((size_t)8 << (lg_m - h)) : 0);
    const tw_t* twt = DIF ? T->tw_inv : T->tw_fwd;
    if (logL >= 6 && logL <= 11 && ncols % 8 == 0 && row_stride % 4 == 0 && poly_stride % 4 == 0 && ((uintptr_t)d & 15) == 0 &&
        env_int("B200_NTT_PERSISTENT", 1)) {
        // persistent double-buffered kernel: 2 x (L x 8 words) of shared memory per CTA
        const size_t sm = (size_t)2 * (((size_t)8 << logL) + ((size_t)8 << logL) / 16) * 4 + ((size_t)8 << logL);
        const uint32_t tpp = ncols / 8, num_tiles = tpp * count;
        uint32_t per_sm = (uint32_t)(226 * 1024 / (sm + 1024)); if (per_sm > 4) per_sm = 4; if (per_sm < 1) per_sm = 1;
        uint32_t grid = (uint32_t)T->sm_count * per_sm; if (grid > num_tiles) grid = num_tiles;
        if (logL == 10 && env_int("B200_NTT_R32", 1)) {
            // radix-32 register kernel: 256 threads, 2 x 1056 x 8 words of tile + the 1024-entry stage table
            const size_t sm32 = (size_t)2 * (1024 + 32) * 8 * 4 + 1024 * 8;
            // B200_NTT_R32_DIRECT: bit 0 = inverse (DIF) pass, bit 1 = forward (DIT) pass
            if (!(DIF && pow_g != nullptr) && (env_int("B200_NTT_R32_DIRECT", 3) & (DIF ? 1 : 2))) {
                const size_t smd = (size_t)(1024 + 32) * 8 * 4 + 1024 * 8;
                const int mb = env_int("B200_NTT_R32D_MINB", 4) == 5 ? 5 : 4;     // CTAs per SM (60 / 48 registers); 5 measured 12 % slower (profiles/ntt_minb_r02.txt)
                uint32_t gd = (uint32_t)T->sm_count * (uint32_t)mb; if (gd > num_tiles) gd = num_tiles;
                const int lg = row_stride == 1024 ? 10 : (row_stride == 4096 ? 12 : -1);
#define B200_R32D_LAUNCH(RS) { auto kp = mb == 5 ? k_ntt_strided_r32d<DIF, RS, 5> : k_ntt_strided_r32d<DIF, RS, 4>; \
                    cudaError_t e = cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smd); if (e != cudaSuccess) return e; \
                    B200_LAUNCH(kp)<<<gd, 256, smd, s>>>(d, row_stride, tpp, num_tiles, poly_stride, twt); return cudaGetLastError(); }
                if (lg == 10) B200_R32D_LAUNCH(10) else if (lg == 12) B200_R32D_LAUNCH(12)
                else { auto kp = k_ntt_strided_r32d<DIF, -1, 4>; gd = (uint32_t)T->sm_count * 4u; if (gd > num_tiles) gd = num_tiles;
                    cudaError_t e = cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smd); if (e != cudaSuccess) return e;
                    B200_LAUNCH(kp)<<<gd, 256, smd, s>>>(d, row_stride, tpp, num_tiles, poly_stride, twt); return cudaGetLastError(); }
#undef B200_R32D_LAUNCH
            }
            // CTAs per SM: measured best per direction (tools/time_ntt2.py); the inverse pass runs the same at 2 and 3 with or without
            // the twiddle epilogue (profiles/ntt_tw_in_b_r02.txt)
            const bool pw = DIF && pow_g != nullptr;
            const int nb = env_int("B200_NTT_R32_MINB", DIF ? 2 : 3) == 3 ? 3 : 2;
            uint32_t g32 = (uint32_t)T->sm_count * (uint32_t)nb; if (g32 > num_tiles) g32 = num_tiles;
            const int lgrs = row_stride == 1024 ? 10 : (row_stride == 4096 ? 12 : -1);
#define B200_R32_LAUNCH(NB, RS) { auto kp = pw ? k_ntt_strided_r32<DIF, NB, RS, DIF> : k_ntt_strided_r32<DIF, NB, RS, false>; \
                cudaError_t e = cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm32); if (e != cudaSuccess) return e; \
                B200_LAUNCH(kp)<<<g32, 256, sm32, s>>>(d, row_stride, tpp, num_tiles, poly_stride, twt, pow_g, lg_m); return cudaGetLastError(); }
            if (nb == 3) { if (lgrs == 10) B200_R32_LAUNCH(3, 10) else if (lgrs == 12) B200_R32_LAUNCH(3, 12) else B200_R32_LAUNCH(3, -1) }
            else { if (lgrs == 10) B200_R32_LAUNCH(2, 10) else if (lgrs == 12) B200_R32_LAUNCH(2, 12) else B200_R32_LAUNCH(2, -1) }
#undef B200_R32_LAUNCH
        }
#define B200_STRIDED_PF_CASE(LL) case LL: { auto kp = k_ntt_strided_pf<LL, DIF>; \
            cudaError_t e = cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); if (e != cudaSuccess) return e; \
            B200_LAUNCH(kp)<<<grid, env_int("B200_NTT_THREADS", 512), sm, s>>>(d, row_stride, tpp, num_tiles, poly_stride, twt, pow_g, lg_m); return cudaGetLastError(); }
        if (env_int("B200_NTT_PF", 1)) switch (logL) { B200_STRIDED_PF_CASE(6) B200_STRIDED_PF_CASE(7) B200_STRIDED_PF_CASE(8) B200_STRIDED_PF_CASE(9) B200_STRIDED_PF_CASE(10) B200_STRIDED_PF_CASE(11) default: break; }
#undef B200_STRIDED_PF_CASE
#define B200_STRIDED_P_CASE(LL) case LL: { auto kp = k_ntt_strided_p<LL, DIF>; \
            cudaError_t e = cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); if (e != cudaSuccess) return e; \
            B200_LAUNCH(kp)<<<grid, env_int("B200_NTT_THREADS", 512), sm, s>>>(d, row_stride, tpp, num_tiles, poly_stride, twt, pow_g, lg_m); return cudaGetLastError(); }
        switch (logL) { B200_STRIDED_P_CASE(6) B200_STRIDED_P_CASE(7) B200_STRIDED_P_CASE(8) B200_STRIDED_P_CASE(9) B200_STRIDED_P_CASE(10) B200_STRIDED_P_CASE(11) default: break; }
#undef B200_STRIDED_P_CASE
    }
    if (logL >= 6 && logL <= 11) {
        const size_t sm = ((size_t)(1u << logL) << lgTW) * 4;
#define B200_STRIDED_CASE(LL) case LL: { auto kc = k_ntt_strided_c<LL, DIF>; \
            cudaError_t e = cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); if (e != cudaSuccess) return e; \
            B200_LAUNCH(kc)<<<tiles_per_poly * count, env_int("B200_NTT_THREADS", 256), sm, s>>>(d, lgTW, row_stride, tiles_per_poly, poly_stride, twt, pow_g, lg_m); return cudaGetLastError(); }
        switch (logL) { B200_STRIDED_CASE(6) B200_STRIDED_CASE(7) B200_STRIDED_CASE(8) B200_STRIDED_CASE(9) B200_STRIDED_CASE(10) B200_STRIDED_CASE(11) default: break; }
#undef B200_STRIDED_CASE
    }
    auto kern = k_ntt_strided<DIF>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    B200_LAUNCH(kern)<<<tiles_per_poly * count, 512, smem, s>>>(d, logL, lgTW, row_stride, tiles_per_poly, poly_stride, twt, pow_g, lg_m);
    return cudaGetLastError();
}

template <bool DIF>
static cudaError_t run_contig(const DeviceTables* T, uint32_t* out, const uint32_t* in, uint32_t logLc, uint32_t lg_e, uint32_t rows_per_poly,
                              uint32_t count, size_t in_stride, size_t out_stride, const tw_t* pow_g, uint32_t lg_m,
                              uint32_t lg_rows, uint32_t scale, cudaStream_t s, const tw_t* p3lo = nullptr, const tw_t* p3hi = nullptr,
                              bool* shift_done = nullptr) {
    // per-element tables in data layout for the kernels that can use them (nullptr: the two-table decomposition)
    const tw_t* tw_full = nullptr; const tw_t* zk_full = nullptr;
    uint32_t rpc = logLc >= 12 ? 1 : (1u << (12 - logLc));
    const uint64_t total_rows = (uint64_t)rows_per_poly * count;
    if (rpc > total_rows) rpc = (uint32_t)total_rows;
    const uint32_t h = (lg_m + 1) / 2;
    const uint32_t Lc = 1u << logLc;
    size_t smem = ((size_t)rpc * (Lc + (Lc >> 4)) + 2) * 4 + (size_t)Lc * 8 + (pow_g ? ((size_t)8 << h) + ((size_t)8 << (lg_m - h)) : 0);
    const uint32_t grid = (uint32_t)((total_rows + rpc - 1) / rpc);
    const tw_t* twt = DIF ? T->tw_inv : T->tw_fwd;
    uint32_t lg_rpp = 0; while ((1u << lg_rpp) < rows_per_poly) lg_rpp++;
    if (logLc >= 8 && logLc <= 13 && (DIF ? lg_e == 0 : (lg_e == 2 || lg_e == 0)) && (1u << lg_rpp) == rows_per_poly && env_int("B200_NTT_FUSED", 1) &&
        (((uintptr_t)in | (uintptr_t)out) & 15) == 0 && in_stride % 4 == 0 && out_stride % 4 == 0) {
        const size_t sm = (size_t)rpc * (Lc + (Lc >> 4)) * 4;
        if (DIF && logLc == 10 && env_int("B200_NTT_INVB_R32", 1)) {
            if (pow_g && lg_m == logLc + lg_rpp) tw_full = get_full_table(T, FULL_INV, lg_m, lg_rpp);
            if (p3lo) zk_full = get_full_table(T, FULL_ZK, logLc + lg_rpp, lg_rpp);
            const size_t smr = (size_t)8 * (1024 + 32) * 4 + 1024 * 8;
            const int mb = env_int("B200_NTT_INVB_R32_MINB", 3) == 4 ? 4 : 3;     // CTAs per SM (80 / 64 registers): no measurable difference
            auto kr = mb == 4 ? k_ntt_invb_r32<4> : k_ntt_invb_r32<3>;
            cudaError_t e = cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr); if (e != cudaSuccess) return e;
            // one CTA per 8 rows by default; B200_NTT_INVB_R32_WAVES = w > 0 caps the grid at w resident waves (persistent loop)
            uint32_t gr = (uint32_t)((total_rows + 7) / 8);
            const uint32_t waves = (uint32_t)env_int("B200_NTT_INVB_R32_WAVES", 0), gmax = (uint32_t)T->sm_count * (uint32_t)mb * waves;
            if (waves && gr > gmax) gr = gmax;
            B200_LAUNCH(kr)<<<gr, 256, smr, s>>>(out, in, lg_rpp, (uint32_t)total_rows, in_stride, out_stride, twt, scale, p3lo, p3hi, pow_g, lg_m, tw_full, zk_full);
            if (shift_done) *shift_done = p3lo != nullptr;
            return cudaGetLastError();
        }
        if (!DIF && pow_g && lg_m == logLc + lg_rpp && lg_rows == lg_rpp) tw_full = get_full_table(T, FULL_FWD, lg_m, lg_rpp);
        if (DIF && pow_g && lg_m == logLc + lg_rpp) tw_full = get_full_table(T, FULL_INV, lg_m, lg_rpp);
        if (DIF && p3lo) zk_full = get_full_table(T, FULL_ZK, logLc + lg_rpp, lg_rpp);
#define B200_FUSED_CASE(LL) case LL: { cudaError_t e; \
            if (DIF) { auto kf = k_ntt_invb<LL>; e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); if (e != cudaSuccess) return e; \
                B200_LAUNCH(kf)<<<grid, 256, sm, s>>>(out, in, rpc, lg_rpp, (uint32_t)total_rows, in_stride, out_stride, twt, scale, p3lo, p3hi, pow_g, lg_m, tw_full, zk_full); \
                if (shift_done) *shift_done = p3lo != nullptr; } \
            else if (lg_e == 2) { auto kf = env_int("B200_NTT_FWD1_MINB", (tw_full || !pow_g) ? 6 : 5) == 5 ? k_ntt_fwd1<LL, 2, 5> : k_ntt_fwd1<LL, 2>; /* two-table twiddle: 5 CTAs leave the gathers their L1 */ e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); if (e != cudaSuccess) return e; \
                if (scale) break; \
                B200_LAUNCH(kf)<<<grid, 256, sm, s>>>(out, in, rpc, lg_rpp, (uint32_t)total_rows, in_stride, out_stride, twt, pow_g, lg_m, lg_rows, tw_full); } \
            else { auto kf = k_ntt_fwd1<LL, 0>; e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); if (e != cudaSuccess) return e; \
                if (scale) break; \
                B200_LAUNCH(kf)<<<grid, 256, sm, s>>>(out, in, rpc, lg_rpp, (uint32_t)total_rows, in_stride, out_stride, twt, pow_g, lg_m, lg_rows, tw_full); } \
            return cudaGetLastError(); }
        switch (logLc) { B200_FUSED_CASE(8) B200_FUSED_CASE(9) B200_FUSED_CASE(10) B200_FUSED_CASE(11) B200_FUSED_CASE(12) B200_FUSED_CASE(13) default: break; }
#undef B200_FUSED_CASE
    }
    // only the fused inverse kernels apply an inter-pass twiddle on their loads; twiddle_in_pass_b() asks for it under exactly their
    // conditions, so reaching this point with one is a dispatch bug: fail instead of returning an untwiddled transform
    if (DIF && pow_g) return cudaErrorInvalidValue;
    if (logLc >= 8 && logLc <= 13 && (DIF ? lg_e == 0 : lg_e == 2) && (1u << lg_rpp) == rows_per_poly) {
        const size_t sm = (size_t)rpc * (Lc + (Lc >> 4)) * 4;
#define B200_CONTIG_CASE(LL) case LL: { auto kc = k_ntt_contig_c<LL, DIF ? 0 : 2, DIF>; \
            cudaError_t e = cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); if (e != cudaSuccess) return e; \
            B200_LAUNCH(kc)<<<grid, 256, sm, s>>>(out, in, rpc, lg_rpp, (uint32_t)total_rows, in_stride, out_stride, twt, pow_g, lg_m, lg_rows, scale); \
            return cudaGetLastError(); }
        switch (logLc) { B200_CONTIG_CASE(8) B200_CONTIG_CASE(9) B200_CONTIG_CASE(10) B200_CONTIG_CASE(11) B200_CONTIG_CASE(12) B200_CONTIG_CASE(13) default: break; }
#undef B200_CONTIG_CASE
    }
    auto kern = k_ntt_contig<DIF>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    B200_LAUNCH(kern)<<<grid, 256, smem, s>>>(out, in, logLc, lg_e, rpc, rows_per_poly, (uint32_t)total_rows, in_stride, out_stride,
                                 twt, pow_g, lg_m, lg_rows, scale);
    return cudaGetLastError();
}

// true when pass B of a two-pass inverse transform runs in the fused kernel k_ntt_invb (run_contig<true>'s first branch), which can
// apply the inter-pass twiddle on load; B200_NTT_TW_IN_B=0 keeps it in pass A's epilogue (round 1's arrangement) for A/B timing
// Without a per-element table (sizes above B200_NTT_FULL_MAX_LG) the twiddle is two gathers from tables of up to 48 KB: that only pays
// in pass B for rows of <= 1024 values -- the longer rows' kernels leave too little L1 beside their shared memory (2^23 / 2^24: 33 %
// slower, profiles/ntt_sweep_r02_v7.jsonl against v8) -- so those sizes keep the twiddle in pass A's epilogue.
static bool twiddle_in_pass_b(const DeviceTables* T, const uint32_t* d_io, uint32_t lg_n, uint32_t n1, size_t N) {
    const uint32_t n2 = lg_n - n1;
    if (!(n2 >= 8 && n2 <= 13 && env_int("B200_NTT_FUSED", 1) && env_int("B200_NTT_TW_IN_B", 1) && ((uintptr_t)d_io & 15) == 0 && N % 4 == 0))
        return false;
    return n2 <= 10 || get_full_table(T, FULL_INV, lg_n, n1) != nullptr;
}

cudaError_t launch_batch_intt(const DeviceTables* T, uint32_t* d_io, uint32_t lg_n, uint32_t count, cudaStream_t s) {
    if (lg_n == 0 || count == 0) return cudaSuccess;
    if (lg_n > MAX_LG) return cudaErrorInvalidValue;
    const size_t N = (size_t)1 << lg_n;
    if (lg_n <= 12) {
        const uint32_t scale = h_inv(h_to_mont((uint32_t)N));
        return run_contig<true>(T, d_io, d_io, lg_n, 0, 1, count, N, N, nullptr, 0, 0, scale, s);
    }
    if (lg_n > MAX_LG_2PASS) {
        // three passes: the outer strided DIF over 2^BIG_N1 rows (its twiddle table carries 1/2^BIG_N1), then every contiguous row of
        // 2^(lg_n - BIG_N1) values is an inverse transform of its own (two passes, carrying the rest of the normalisation)
        const uint32_t n2 = lg_n - BIG_N1;
        if ((uint64_t)count << BIG_N1 > 0xffffffffull) return cudaErrorInvalidValue;
        cudaError_t e = run_strided<true>(T, d_io, BIG_N1, 1u << n2, 1u << n2, count, N, T->pow_inv[lg_n], lg_n, s);
        if (e != cudaSuccess) return e;
        return launch_batch_intt(T, d_io, n2, count << BIG_N1, s);
    }
    const uint32_t n1 = split_n1(lg_n), n2 = lg_n - n1;
    const bool tw_b = twiddle_in_pass_b(T, d_io, lg_n, n1, N);
    // pass A: DIF over i1 (rows of stride N2); the inter-pass twiddle (1/N folded into its table) here or on pass B's loads
    cudaError_t e = run_strided<true>(T, d_io, n1, 1u << n2, 1u << n2, count, N, tw_b ? nullptr : T->pow_inv[lg_n], lg_n, s);
    if (e != cudaSuccess) return e;
    // pass B: DIF over each contiguous row of N2
    return run_contig<true>(T, d_io, d_io, n2, 0, 1u << n1, count, N, N, tw_b ? T->pow_inv[lg_n] : nullptr, lg_n, 0, 0, s);
}

// K1 + K2 in one go: iNTT whose last pass multiplies the coefficient of x^d by 3^d (falls back to two launches when the
// fused kernel does not cover the shape).
cudaError_t launch_batch_intt_shift(const DeviceTables* T, uint32_t* d_io, uint32_t lg_n, uint32_t count, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    if (lg_n == 0 || lg_n > MAX_LG) return lg_n == 0 ? cudaSuccess : cudaErrorInvalidValue;     // 3^0 = 1 for size-1 polynomials
    const size_t N = (size_t)1 << lg_n;
    bool done = false;
    cudaError_t e;
    if (lg_n > MAX_LG_2PASS) {          // three-pass sizes: the coset shift is a pass of its own
        e = launch_batch_intt(T, d_io, lg_n, count, s);
        return e != cudaSuccess ? e : launch_zk_shift(T, d_io, lg_n, count, s);
    }
    if (lg_n <= 12) {
        const uint32_t scale = h_inv(h_to_mont((uint32_t)N));
        e = run_contig<true>(T, d_io, d_io, lg_n, 0, 1, count, N, N, nullptr, 0, 0, scale, s, T->p3lo, T->p3hi, &done);
    } else {
        const uint32_t n1 = split_n1(lg_n), n2 = lg_n - n1;
        const bool tw_b = twiddle_in_pass_b(T, d_io, lg_n, n1, N);
        e = run_strided<true>(T, d_io, n1, 1u << n2, 1u << n2, count, N, tw_b ? nullptr : T->pow_inv[lg_n], lg_n, s);
        if (e != cudaSuccess) return e;
        e = run_contig<true>(T, d_io, d_io, n2, 0, 1u << n1, count, N, N, tw_b ? T->pow_inv[lg_n] : nullptr, lg_n, 0, 0, s, T->p3lo, T->p3hi, &done);
    }
    if (e != cudaSuccess) return e;
    return done ? cudaSuccess : launch_zk_shift(T, d_io, lg_n, count, s);
}

cudaError_t launch_batch_expand_ntt(const DeviceTables* T, uint32_t* d_out, const uint32_t* d_in, uint32_t lg_n, uint32_t lg_e,
                                    uint32_t count, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    const uint32_t lg_m = lg_n + lg_e;
    if (lg_m == 0)        // size-1 transforms are the identity
        return d_out == d_in ? cudaSuccess : cudaMemcpyAsync(d_out, d_in, (size_t)count * 4, cudaMemcpyDeviceToDevice, s);
    const size_t N = (size_t)1 << lg_n, M = (size_t)1 << lg_m;
    if (lg_m <= 13) return run_contig<false>(T, d_out, d_in, lg_m, lg_e, 1, count, N, M, nullptr, 0, 0, 0, s);
    if (lg_m > MAX_LG) return cudaErrorInvalidValue;
    if (lg_m > MAX_LG_2PASS) {
        // three passes: every row of 2^(lg_n - BIG_N1) coefficients is a forward transform of its own (expand included), then the
        // inter-pass twiddle, then the strided DIT over the 2^BIG_N1 rows
        if (lg_n <= (uint32_t)BIG_N1 || (uint64_t)count << BIG_N1 > 0xffffffffull) return cudaErrorInvalidValue;
        const uint32_t n2 = lg_n - BIG_N1, lg_row = n2 + lg_e;
        cudaError_t e = launch_batch_expand_ntt(T, d_out, d_in, n2, lg_e, count << BIG_N1, s);
        if (e != cudaSuccess) return e;
        const size_t total = (size_t)count << lg_m;
        B200_LAUNCH(k_ntt_row_twiddle)<<<(unsigned)((total / 4 + 255) / 256), 256, 0, s>>>(d_out, lg_row, BIG_N1, total, T->pow_fwd[lg_m], lg_m);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        return run_strided<false>(T, d_out, BIG_N1, 1u << lg_row, 1u << lg_row, count, M, nullptr, 0, s);
    }
    // Column batching: the 2^lg_m-word intermediate of every column goes pass 1 -> pass 2; with `batch` columns per pair of launches it
    // is batch * 4 * M bytes, which for small batches stays in the 126 MB L2 instead of making a round trip through HBM
    // (B200_NTT_COLBATCH; 0 = all columns in one pair of launches).
    const uint32_t batch = (uint32_t)env_int("B200_NTT_COLBATCH", 0);
    if (batch > 0 && count > batch && d_out != d_in) {
        for (uint32_t c0 = 0; c0 < count; c0 += batch) {
            const uint32_t nb = count - c0 < batch ? count - c0 : batch;
            cudaError_t e = launch_batch_expand_ntt(T, d_out + (size_t)c0 * M, d_in + (size_t)c0 * N, lg_n, lg_e, nb, s);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    const uint32_t n1 = split_n1(lg_n, lg_e), n2 = lg_n - n1;
    // pass 1: row rho (N2 coefficients) -> 2^lg_e * N2 values, times w_M^(i0 * bitrev(rho))
    cudaError_t e = run_contig<false>(T, d_out, d_in, n2 + lg_e, lg_e, 1u << n1, count, N, M, T->pow_fwd[lg_m], lg_m, n1, 0, s);
    if (e != cudaSuccess) return e;
    // pass 2: DIT over rho (rows of stride 2^lg_e * N2)
    return run_strided<false>(T, d_out, n1, 1u << (n2 + lg_e), 1u << (n2 + lg_e), count, M, nullptr, 0, s);
}

cudaError_t launch_batch_ntt(const DeviceTables* T, uint32_t* d_io, uint32_t lg_n, uint32_t count, cudaStream_t s) {
    return launch_batch_expand_ntt(T, d_io, d_io, lg_n, 0, count, s);
}

cudaError_t launch_zk_shift(const DeviceTables* T, uint32_t* d_io, uint32_t lg_n, uint32_t count, cudaStream_t s) {
    const size_t total = (size_t)count << lg_n;
    if (total == 0) return cudaSuccess;
    B200_LAUNCH(k_zk_shift)<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(d_io, lg_n, total, T->p3lo, T->p3hi);
    return cudaGetLastError();
}

cudaError_t launch_bit_reverse(uint32_t* d_io, uint32_t lg_n, uint32_t count, cudaStream_t s) {
    const size_t total = (size_t)count << lg_n;
    if (total == 0) return cudaSuccess;
    B200_LAUNCH(k_bit_reverse)<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(d_io, lg_n, total);
    return cudaGetLastError();
}

}  // namespace b200
