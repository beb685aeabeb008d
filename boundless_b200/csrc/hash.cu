// Poseidon2 sponge / Merkle tree / Fiat-Shamir transcript kernels for sm_100a (kernel 2 of the hot path).
//
// Replaces risc0-sys 1.5.0's sppark_poseidon2_rows / sppark_poseidon2_fold and the
// host-side Poseidon2Rng of risc0-zkp 3.0.3 (un-vendored), reached from
// /root/reference/prover/crates/workflow/src/tasks/prove.rs:44-52.  Semantics follow SURVEY.md Appendix A:
// overwrite-mode sponge (rate 16) over a column-major matrix, hash_pair for interior nodes, nodes[1] = root.
//
// B200 notes: leaf hashing is INT32-pipe bound (1356 Montgomery multiplies per permutation), not HBM bound;
// one thread owns one row so that a warp reads 128 contiguous bytes per column (DRAM traffic = 1.001x the algorithmic
// bytes, profiles/traffic_r01.json); 16 loads are in flight per thread before each permutation.  The transcript lives in
// device memory so a whole proof is enqueued without a host round trip (CUDA-graph friendly).
#include "internal.h"
#include "poseidon2.cuh"
#include <cstdlib>

namespace b200 {

// K4 ---------------------------------------------------------------------------------------------------
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_p2_rows(uint32_t* __restrict__ out, const uint32_t* __restrict__ m, uint32_t rows,
                                                           uint32_t cols, size_t col_stride) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rows) return;
    uint32_t st[24];
#pragma unroll
    for (int i = 0; i < 24; i++) st[i] = 0;
    const uint32_t* src = m + j;
    uint32_t c = 0;
    for (; c + 16 <= cols; c += 16) {
#pragma unroll
        for (int i = 0; i < 16; i++) st[i] = __ldg(src + (size_t)(c + i) * col_stride);
        p2_permute(st);
    }
    const uint32_t rem = cols - c;
    if (rem != 0 || cols == 0) {
#pragma unroll
        for (int i = 0; i < 16; i++) st[i] = (uint32_t)i < rem ? __ldg(src + (size_t)(c + i) * col_stride) : 0u;
        p2_permute(st);
    }
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)j * 8);
    o[0] = make_uint4(st[0], st[1], st[2], st[3]);
    o[1] = make_uint4(st[4], st[5], st[6], st[7]);
}

// K5: one layer, out[i] = hash_pair(in[2i], in[2i+1]) -------------------------------------------------------
__device__ __forceinline__ void load_pair(uint32_t (&st)[24], const uint32_t* in) {
    const uint4* p = reinterpret_cast<const uint4*>(in);
    uint4 a = p[0], b = p[1], c = p[2], d = p[3];
    st[0] = a.x; st[1] = a.y; st[2] = a.z; st[3] = a.w; st[4] = b.x; st[5] = b.y; st[6] = b.z; st[7] = b.w;
    st[8] = c.x; st[9] = c.y; st[10] = c.z; st[11] = c.w; st[12] = d.x; st[13] = d.y; st[14] = d.z; st[15] = d.w;
#pragma unroll
    for (int i = 16; i < 24; i++) st[i] = 0;
}
__global__ void __launch_bounds__(256, 2) k_p2_fold(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, uint32_t n_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    uint32_t st[24];
    load_pair(st, in + (size_t)i * 16);
    p2_permute(st);
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)i * 8);
    o[0] = make_uint4(st[0], st[1], st[2], st[3]);
    o[1] = make_uint4(st[4], st[5], st[6], st[7]);
}
// top of the tree: layers with <= 1024 nodes handled by one CTA.  Wide layers (>= 64 nodes) use one thread per node; the last six
// layers (32 ... 1 nodes), where a one-thread permutation would leave the CTA waiting ~23 us per layer, use one WARP per node.
__global__ void __launch_bounds__(1024, 1) k_p2_fold_top(uint32_t* nodes, uint32_t top_nodes) {
    for (uint32_t sz = top_nodes; sz >= 64; sz >>= 1) {
        const uint32_t j = threadIdx.x;
        if (j < sz) {
            const uint32_t i = sz + j;
            uint32_t st[24];
            load_pair(st, nodes + (size_t)i * 16);
            p2_permute(st);
            uint4* o = reinterpret_cast<uint4*>(nodes + (size_t)i * 8);
            o[0] = make_uint4(st[0], st[1], st[2], st[3]);
            o[1] = make_uint4(st[4], st[5], st[6], st[7]);
        }
        __syncthreads();
    }
}
// one WARP per parent, any number of CTAs: the form for the narrow layers of a tree (<= 2^12 parents), where the one-thread form leaves
// most of the chip idle and every layer costs a full single-thread permutation latency (~23 us); in warp form a layer is ~5 us.
__global__ void __launch_bounds__(256) k_p2_fold_w(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, uint32_t n_out) {
    P2Warp w; w.init();
    const uint32_t i = blockIdx.x * 8u + (threadIdx.x >> 5);
    if (i >= n_out) return;                                   // warp-uniform
    uint32_t x = w.lane < 16 ? in[(size_t)i * 16 + w.lane] : 0u;
    x = w.permute(x);
    if (w.lane < 8) out[(size_t)i * 8 + w.lane] = x;
}
__global__ void __launch_bounds__(1024, 1) k_p2_fold_top_w(uint32_t* nodes, uint32_t top_nodes) {
    P2Warp w; w.init();
    const uint32_t wid = threadIdx.x >> 5;
    for (uint32_t sz = top_nodes; sz >= 1; sz >>= 1) {
        if (wid < sz) {
            const uint32_t i = sz + wid;
            uint32_t x = w.lane < 16 ? nodes[(size_t)i * 16 + w.lane] : 0u;
            x = w.permute(x);
            if (w.lane < 8) nodes[(size_t)i * 8 + w.lane] = x;
        }
        __syncthreads();
    }
}

cudaError_t launch_poseidon2_rows(uint32_t* d_out, const uint32_t* d_matrix, uint32_t rows, uint32_t cols, size_t col_stride,
                                  cudaStream_t s) {
    if (rows == 0) return cudaSuccess;
    // launch-shape A/B switch (tools/time_p2_variants.py).  Default <256, 4>: 48 registers -> 5 CTAs = 40 warps per SM, measured best for
    // the round-2 arithmetic (profiles/p2_variants_r02.txt: 23.5 ms for 2^22 x 208 against 23.9 ms with <256, 2>; 40-register shapes lose)
    static const int cfg = getenv("B200_P2_CFG") ? atoi(getenv("B200_P2_CFG")) : 1;
    switch (cfg) {
        case 0: B200_LAUNCH(k_p2_rows<256, 2>)<<<(rows + 255) / 256, 256, 0, s>>>(d_out, d_matrix, rows, cols, col_stride); break;
        case 2: B200_LAUNCH(k_p2_rows<512, 2>)<<<(rows + 511) / 512, 512, 0, s>>>(d_out, d_matrix, rows, cols, col_stride); break;
        case 3: B200_LAUNCH(k_p2_rows<1024, 1>)<<<(rows + 1023) / 1024, 1024, 0, s>>>(d_out, d_matrix, rows, cols, col_stride); break;
        case 4: B200_LAUNCH(k_p2_rows<128, 4>)<<<(rows + 127) / 128, 128, 0, s>>>(d_out, d_matrix, rows, cols, col_stride); break;
        case 5: B200_LAUNCH(k_p2_rows<256, 6>)<<<(rows + 255) / 256, 256, 0, s>>>(d_out, d_matrix, rows, cols, col_stride); break;
        case 6: B200_LAUNCH(k_p2_rows<128, 12>)<<<(rows + 127) / 128, 128, 0, s>>>(d_out, d_matrix, rows, cols, col_stride); break;
        case 7: B200_LAUNCH(k_p2_rows<256, 5>)<<<(rows + 255) / 256, 256, 0, s>>>(d_out, d_matrix, rows, cols, col_stride); break;
        default: B200_LAUNCH(k_p2_rows<256, 4>)<<<(rows + 255) / 256, 256, 0, s>>>(d_out, d_matrix, rows, cols, col_stride); break;
    }
    return cudaGetLastError();
}
cudaError_t launch_poseidon2_fold(uint32_t* d_out, const uint32_t* d_in, uint32_t n_out, cudaStream_t s) {
    if (n_out == 0) return cudaSuccess;
    B200_LAUNCH(k_p2_fold)<<<(n_out + 255) / 256, 256, 0, s>>>(d_out, d_in, n_out);
    return cudaGetLastError();
}
cudaError_t launch_poseidon2_fold_tree(uint32_t* d_nodes, uint32_t lg_rows, cudaStream_t s) {
    if (lg_rows == 0) return cudaSuccess;
    // wide layers: one thread per parent (throughput form).  Layers of <= 2^12 parents: one warp per parent over as many CTAs as it
    // takes (latency form; measured cross-over, profiles/ncu_brief_r02_round2_kernels.txt: 17 us against ~25 us at 4096 parents, but
    // 26 / 47 us at 8192 / 16384; B200_FOLD_WARP_MAX overrides it for A/B timing).  The last six layers: one CTA, no relaunch.
    static const uint32_t warp_max = getenv("B200_FOLD_WARP_MAX") ? (uint32_t)atoi(getenv("B200_FOLD_WARP_MAX")) : (1u << 12);
    uint32_t sz = 1u << (lg_rows - 1);
    while (sz > 1024 && sz > warp_max) {
        cudaError_t e = launch_poseidon2_fold(d_nodes + (size_t)sz * 8, d_nodes + (size_t)sz * 16, sz, s);
        if (e != cudaSuccess) return e;
        sz >>= 1;
    }
    if (warp_max >= 64) {
        while (sz > 32) {
            B200_LAUNCH(k_p2_fold_w)<<<(sz + 7) / 8, 256, 0, s>>>(d_nodes + (size_t)sz * 8, d_nodes + (size_t)sz * 16, sz);
            sz >>= 1;
        }
    } else if (sz >= 64) {
        B200_LAUNCH(k_p2_fold_top)<<<1, 1024, 0, s>>>(d_nodes, sz);        // round 1's one-CTA thread form (kept for the A/B)
    }
    B200_LAUNCH(k_p2_fold_top_w)<<<1, 1024, 0, s>>>(d_nodes, sz < 32 ? sz : 32u);
    return cudaGetLastError();
}

// Transcript (Poseidon2Rng) -----------------------------------------------------------
// The transcript kernels run as ONE WARP (P2Warp, poseidon2.cuh): they are single permutation chains on the proof's critical path.
__device__ __forceinline__ uint32_t tr_load_w(const P2Warp& w, const Transcript* t) { return w.lane < 24 ? t->cells[w.lane] : 0u; }
__device__ __forceinline__ void tr_store_w(const P2Warp& w, Transcript* t, uint32_t x, uint32_t used) {
    if (w.lane < 24) t->cells[w.lane] = x;
    if (w.lane == 0) t->pool_used = used;
}
// overwrite-mode sponge over `count` elements; returns this lane's cell (digest = lanes 0..7)
__device__ __forceinline__ uint32_t sponge_elems_w(const P2Warp& w, const uint32_t* __restrict__ e, uint32_t count) {
    uint32_t x = 0, k = 0;
    for (; k + 16 <= count; k += 16) {
        if (w.lane < 16) x = e[k + w.lane];
        x = w.permute(x);
    }
    const uint32_t rem = count - k;
    if (rem != 0 || count == 0) {
        if (w.lane < 16) x = w.lane < rem ? e[k + w.lane] : 0u;
        x = w.permute(x);
    }
    return x;
}
__global__ void k_iop_init(Transcript* t) {
    for (int i = 0; i < 24; i++) t->cells[i] = 0;
    t->pool_used = 0;
}
__global__ void __launch_bounds__(32) k_iop_commit(Transcript* t, const uint32_t* __restrict__ digest) {
    P2Warp w; w.init();
    uint32_t x = tr_load_w(w, t);
    if (w.lane < 8) x = fp_add(x, digest[w.lane]);
    tr_store_w(w, t, w.permute(x), 0);
}
__global__ void __launch_bounds__(32) k_iop_commit_elems(Transcript* t, const uint32_t* __restrict__ elems, uint32_t count, uint32_t* digest_out) {
    P2Warp w; w.init();
    const uint32_t d = sponge_elems_w(w, elems, count);
    if (digest_out && w.lane < 8) digest_out[w.lane] = d;
    uint32_t x = tr_load_w(w, t);
    if (w.lane < 8) x = fp_add(x, d);
    tr_store_w(w, t, w.permute(x), 0);
}
__global__ void __launch_bounds__(32) k_hash_elems(uint32_t* digest_out, const uint32_t* __restrict__ elems, uint32_t count) {
    P2Warp w; w.init();
    const uint32_t d = sponge_elems_w(w, elems, count);
    if (w.lane < 8) digest_out[w.lane] = d;
}
__global__ void __launch_bounds__(32) k_hash_pair_one(uint32_t* out, const uint32_t* a, const uint32_t* b) {
    P2Warp w; w.init();
    uint32_t x = w.lane < 8 ? a[w.lane] : (w.lane < 16 ? b[w.lane - 8] : 0u);
    x = w.permute(x);
    if (w.lane < 8) out[w.lane] = x;
}
// draws n elems; bits == 0 -> raw Montgomery elems, else as_u32() & mask
__global__ void __launch_bounds__(32) k_iop_draw(Transcript* t, uint32_t* out, uint32_t n, uint32_t bits) {
    P2Warp w; w.init();
    uint32_t x = tr_load_w(w, t);
    uint32_t used = t->pool_used;
    for (uint32_t k = 0; k < n; k++) {
        if (used == 16) { x = w.permute(x); used = 0; }
        uint32_t v = __shfl_sync(0xffffffffu, x, used);
        used++;
        if (bits) { v = fp_from_mont(v); if (bits < 32) v &= (1u << bits) - 1; }
        if (w.lane == 0) out[k] = v;
    }
    tr_store_w(w, t, x, used);
}

// seal[0..8) = circuit header; seal[8..16) = digest of the segment seed (segments) or left for the caller (recursion)
__global__ void k_set_globals(uint32_t* seal, uint32_t po2, uint32_t w_code, uint32_t w_data, uint32_t w_accum, uint32_t kind,
                              uint64_t seed, int hash_seed) {
    seal[0] = po2; seal[1] = w_code; seal[2] = w_data; seal[3] = w_accum; seal[4] = kind; seal[5] = 0; seal[6] = 0; seal[7] = 0;
    if (hash_seed) {
        uint32_t st[24];
#pragma unroll
        for (int i = 0; i < 24; i++) st[i] = 0;
        st[0] = (uint32_t)(seed & 0x3FFFFFFF); st[1] = (uint32_t)((seed >> 30) & 0x3FFFFFFF); st[2] = (uint32_t)(seed >> 60);
        p2_permute(st);
        for (int i = 0; i < 8; i++) seal[8 + i] = st[i];
    }
}
cudaError_t launch_set_globals(uint32_t* d_seal, uint32_t po2, uint32_t w_code, uint32_t w_data, uint32_t w_accum, uint32_t kind,
                               uint64_t seed, int hash_seed, cudaStream_t s) {
    B200_LAUNCH(k_set_globals)<<<1, 1, 0, s>>>(d_seal, po2, w_code, w_data, w_accum, kind, seed, hash_seed);
    return cudaGetLastError();
}

cudaError_t launch_iop_init(Transcript* t, cudaStream_t s) { B200_LAUNCH(k_iop_init)<<<1, 1, 0, s>>>(t); return cudaGetLastError(); }
cudaError_t launch_iop_commit(Transcript* t, const uint32_t* d, cudaStream_t s) { B200_LAUNCH(k_iop_commit)<<<1, 32, 0, s>>>(t, d); return cudaGetLastError(); }
cudaError_t launch_iop_commit_elems(Transcript* t, const uint32_t* e, uint32_t count, uint32_t* dout, cudaStream_t s) {
    B200_LAUNCH(k_iop_commit_elems)<<<1, 32, 0, s>>>(t, e, count, dout); return cudaGetLastError();
}
cudaError_t launch_iop_draw_ext(Transcript* t, uint32_t* out, uint32_t n_ext, cudaStream_t s) {
    B200_LAUNCH(k_iop_draw)<<<1, 32, 0, s>>>(t, out, n_ext * 4, 0); return cudaGetLastError();
}
cudaError_t launch_iop_draw_bits(Transcript* t, uint32_t* out, uint32_t n, uint32_t bits, cudaStream_t s) {
    B200_LAUNCH(k_iop_draw)<<<1, 32, 0, s>>>(t, out, n, bits); return cudaGetLastError();
}
cudaError_t launch_hash_elems(uint32_t* dout, const uint32_t* e, uint32_t count, cudaStream_t s) {
    B200_LAUNCH(k_hash_elems)<<<1, 32, 0, s>>>(dout, e, count); return cudaGetLastError();
}
cudaError_t launch_hash_pair_one(uint32_t* out, const uint32_t* a, const uint32_t* b, cudaStream_t s) {
    B200_LAUNCH(k_hash_pair_one)<<<1, 32, 0, s>>>(out, a, b); return cudaGetLastError();
}

}  // namespace b200
