// The small HAL operations around the three inner loops, as standalone C-ABI entry points: K8 (mix_poly_coeffs,
// eltwise_sum_extelem, poly_divide), K9 (gather_sample, Merkle opening), prefix_products, the element-wise helpers and the
// composite commit_group.
//
// Replaces the C wrappers of risc0-sys 1.5.0 that risc0-zkp's CUDA `Hal` calls (un-vendored crates.io dependency of
// /root/reference, Cargo.lock; SURVEY.md 8b "Kernel-level API"): mix_poly_coeffs, eltwise_sum_extelem, eltwise_add_elem,
// eltwise_copy_elem, eltwise_zeroize_elem, supra_poly_divide, prefix_products, gather_sample, scatter.  Reached from
// /root/reference/prover/crates/workflow/src/tasks/prove.rs:44-52 (prove_segment) and :96-104 (lift).
// Layouts are the reference's: Elem buffers are column-major u32 Montgomery words, ExtElem buffers are arrays of 4 words.
// All of these are HBM-streaming kernels (a handful of multiplies per word moved): coalesced 16-byte accesses, grids sized in
// multiples of the SM count, no shared-memory staging beyond the per-CTA power table.
#include "../../include/b200zkp.h"
#include "internal.h"
#include "field.cuh"

namespace b200 {

static int sm_count_of_current() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}
static uint32_t grid_for(size_t work_items, uint32_t threads, uint32_t per_sm) {
    size_t need = (work_items + threads - 1) / threads;
    size_t cap = (size_t)sm_count_of_current() * per_sm;
    if (need < 1) need = 1;
    return (uint32_t)(need < cap ? need : cap);
}

// ---- K8a: mix_poly_coeffs -------------------------------------------------------------------------------------------
// out[combos[i]*count + idx] += mix_start * mix^i * in[i*count + idx].  Each CTA first builds the power table
// pw[i] = mix_start * mix^i in shared memory (one fp4_pow per entry, in parallel), then streams its share of idx:
// every input word is read once, every output ExtElem read and written once.
__global__ void __launch_bounds__(256) k_mix_poly_coeffs(uint32_t* __restrict__ out, const uint32_t* __restrict__ mix_start,
                                                         const uint32_t* __restrict__ mix, const uint32_t* __restrict__ in,
                                                         const uint32_t* __restrict__ combos, uint32_t input_size, uint32_t count,
                                                         uint32_t n_combos) {
    extern __shared__ __align__(16) uint32_t sm_mix[];
    uint32_t* pw = sm_mix;                       // input_size x Fp4
    uint32_t* cb = sm_mix + 4 * (size_t)input_size;
    const Fp4 m = ld_fp4(mix), m0 = ld_fp4(mix_start);
    for (uint32_t i = threadIdx.x; i < input_size; i += blockDim.x) {
        st_fp4(pw + 4 * i, fp4_mul(m0, fp4_pow(m, i)));
        cb[i] = combos[i];
    }
    __syncthreads();
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += (size_t)gridDim.x * blockDim.x) {
        for (uint32_t c = 0; c < n_combos; c++) {
            uint32_t* o = out + 4 * ((size_t)c * count + idx);
            Fp4 acc = ld_fp4(o);
            bool any = false;
            for (uint32_t i = 0; i < input_size; i++) {
                if (cb[i] != c) continue;          // uniform across the CTA
                fp4_fma_fp(acc, ld_fp4(pw + 4 * i), __ldg(in + (size_t)i * count + idx));
                any = true;
            }
            if (any) st_fp4(o, acc);
        }
    }
}

// ---- K8b: eltwise_sum_extelem: AoS ExtElem rows summed, written as 4 planes -------------------------------------------
__global__ void __launch_bounds__(256) k_eltwise_sum_extelem(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, uint32_t count,
                                                             uint32_t to_add) {
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += (size_t)gridDim.x * blockDim.x) {
        Fp4 tot = fp4_zero();
        for (uint32_t i = 0; i < to_add; i++) tot = fp4_add(tot, ld_fp4(in + 4 * ((size_t)i * count + idx)));
#pragma unroll
        for (int j = 0; j < 4; j++) out[(size_t)j * count + idx] = tot.c[j];
    }
}

__global__ void __launch_bounds__(256) k_eltwise_add_elem(uint32_t* __restrict__ out, const uint32_t* __restrict__ a,
                                                          const uint32_t* __restrict__ b, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        out[i] = fp_add(a[i], b[i]);
}
__global__ void __launch_bounds__(256) k_eltwise_zeroize_elem(uint32_t* __restrict__ io, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        if (io[i] == 0xFFFFFFFFu) io[i] = 0u;
}

// ---- prefix_products: inclusive running product of an ExtElem array, chunked three-phase scan -----------------------------
constexpr uint32_t PP_T = 256, PP_E = 8, PP_CH = PP_T * PP_E;
// (1) product of each chunk
__global__ void __launch_bounds__(PP_T) k_pp_chunk_prod(uint32_t* __restrict__ prods, const uint32_t* __restrict__ io, uint32_t count) {
    const size_t start = (size_t)blockIdx.x * PP_CH + (size_t)threadIdx.x * PP_E;
    Fp4 v = fp4_one();
#pragma unroll
    for (uint32_t i = 0; i < PP_E; i++)
        if (start + i < count) v = fp4_mul(v, ld_fp4(io + 4 * (start + i)));
    __shared__ __align__(16) uint32_t red[PP_T * 4];
    st_fp4(red + 4 * threadIdx.x, v);
    __syncthreads();
    for (uint32_t st = PP_T / 2; st >= 1; st >>= 1) {
        if (threadIdx.x < st) st_fp4(red + 4 * threadIdx.x, fp4_mul(ld_fp4(red + 4 * threadIdx.x), ld_fp4(red + 4 * (threadIdx.x + st))));
        __syncthreads();
    }
    if (threadIdx.x == 0) st_fp4(prods + 4 * (size_t)blockIdx.x, ld_fp4(red));
}
// (2) exclusive scan of the chunk products: one warp, 32 chunks per step (shuffle-free: shared memory Hillis-Steele)
__global__ void __launch_bounds__(256) k_pp_scan(uint32_t* __restrict__ carry, const uint32_t* __restrict__ prods, uint32_t nchunks) {
    __shared__ __align__(16) uint32_t sc[256 * 4];
    Fp4 running = fp4_one();
    for (uint32_t base = 0; base < nchunks; base += 256) {
        const uint32_t k = base + threadIdx.x;
        Fp4 v = k < nchunks ? ld_fp4(prods + 4 * (size_t)k) : fp4_one();
        st_fp4(sc + 4 * threadIdx.x, v);
        __syncthreads();
        for (uint32_t off = 1; off < 256; off <<= 1) {     // inclusive scan of 256 products
            Fp4 lhs = fp4_one();
            const bool has = threadIdx.x >= off;
            if (has) lhs = ld_fp4(sc + 4 * (threadIdx.x - off));
            __syncthreads();
            if (has) { v = fp4_mul(lhs, v); st_fp4(sc + 4 * threadIdx.x, v); }
            __syncthreads();
        }
        const Fp4 excl = threadIdx.x ? ld_fp4(sc + 4 * (threadIdx.x - 1)) : fp4_one();
        if (k < nchunks) st_fp4(carry + 4 * (size_t)k, fp4_mul(running, excl));
        running = fp4_mul(running, ld_fp4(sc + 4 * 255));
        __syncthreads();
    }
}
// (3) in-chunk scan, multiplied by the carry of the chunk
__global__ void __launch_bounds__(PP_T) k_pp_apply(uint32_t* __restrict__ io, const uint32_t* __restrict__ carry, uint32_t count) {
    const size_t start = (size_t)blockIdx.x * PP_CH + (size_t)threadIdx.x * PP_E;
    Fp4 x[PP_E];
    Fp4 v = fp4_one();
#pragma unroll
    for (uint32_t i = 0; i < PP_E; i++) {
        if (start + i < count) v = fp4_mul(v, ld_fp4(io + 4 * (start + i)));
        x[i] = v;                                  // inclusive product inside the thread's run
    }
    __shared__ __align__(16) uint32_t sc[PP_T * 4];
    st_fp4(sc + 4 * threadIdx.x, v);
    __syncthreads();
    for (uint32_t off = 1; off < PP_T; off <<= 1) {
        Fp4 lhs = fp4_one();
        const bool has = threadIdx.x >= off;
        if (has) lhs = ld_fp4(sc + 4 * (threadIdx.x - off));
        __syncthreads();
        if (has) { v = fp4_mul(lhs, v); st_fp4(sc + 4 * threadIdx.x, v); }
        __syncthreads();
    }
    Fp4 pre = ld_fp4(carry + 4 * (size_t)blockIdx.x);
    if (threadIdx.x) pre = fp4_mul(pre, ld_fp4(sc + 4 * (threadIdx.x - 1)));
#pragma unroll
    for (uint32_t i = 0; i < PP_E; i++)
        if (start + i < count) st_fp4(io + 4 * (start + i), fp4_mul(pre, x[i]));
}

// ---- K9: gather_sample / scatter / Merkle opening ---------------------------------------------------------------------
__global__ void k_gather_sample(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, size_t idx, uint32_t size, size_t stride) {
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < size; g += gridDim.x * blockDim.x) dst[g] = src[(size_t)g * stride + idx];
}
__global__ void k_scatter(uint32_t* __restrict__ into, const uint32_t* __restrict__ index, uint32_t n_index,
                          const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ values) {
    const uint32_t lo = index[0], hi = index[n_index - 1];
    for (size_t k = (size_t)lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < hi; k += (size_t)gridDim.x * blockDim.x)
        into[offsets[k]] = values[k];
}
__global__ void __launch_bounds__(128) k_merkle_open(uint32_t* __restrict__ out, const uint32_t* __restrict__ nodes,
                                                     const uint32_t* __restrict__ matrix, uint32_t rows, uint32_t cols, uint32_t top_size,
                                                     uint32_t idx) {
    for (uint32_t c = threadIdx.x; c < cols; c += blockDim.x) out[c] = matrix[(size_t)c * rows + idx];
    out += cols;
    uint32_t node = idx + rows, level = 0;
    while (node >= 2 * top_size) {
        if (threadIdx.x < 8) out[level * 8 + threadIdx.x] = nodes[(size_t)(node ^ 1) * 8 + threadIdx.x];
        node >>= 1; level++;
    }
}

}  // namespace b200

using namespace b200;

#define FAIL(...) do { set_error(__VA_ARGS__); return last_error(); } while (0)
#define DONE(what) do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) FAIL("b200: %s: %s", what, cudaGetErrorString(e__)); return nullptr; } while (0)
static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
static bool have_device() { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess && n > 0; }
#define NEED_GPU() do { if (!have_device()) FAIL("b200: no CUDA device available (this library has no CPU path)"); } while (0)

extern "C" {

const char* b200_shutdown(void) {
    compat_release();
    free_tables();
    return nullptr;
}

const char* b200_mix_poly_coeffs(uint32_t* d_out, const uint32_t* d_mix_start, const uint32_t* d_mix, const uint32_t* d_in,
                                 const uint32_t* d_combos, uint32_t input_size, uint32_t count, uint32_t n_combos, void* stream) {
    NEED_GPU();
    if (input_size == 0 || count == 0 || n_combos == 0) return nullptr;
    if (!aligned16(d_out) || !aligned16(d_mix_start) || !aligned16(d_mix)) FAIL("b200_mix_poly_coeffs: ExtElem buffers must be 16-byte aligned");
    const size_t smem = (size_t)input_size * 20;
    if (smem > 200 * 1024) FAIL("b200_mix_poly_coeffs: input_size %u exceeds the 10240 polynomials one call can mix", input_size);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_mix_poly_coeffs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        FAIL("b200_mix_poly_coeffs: cannot reserve %zu bytes of shared memory", smem);
    B200_LAUNCH(k_mix_poly_coeffs)<<<grid_for(count, 256, 4), 256, smem, (cudaStream_t)stream>>>(d_out, d_mix_start, d_mix, d_in, d_combos,
                                                                                              input_size, count, n_combos);
    DONE("mix_poly_coeffs");
}

const char* b200_eltwise_sum_extelem(uint32_t* d_out, const uint32_t* d_in, uint32_t count, uint32_t to_add, void* stream) {
    NEED_GPU();
    if (count == 0) return nullptr;
    if (!aligned16(d_in)) FAIL("b200_eltwise_sum_extelem: ExtElem buffer must be 16-byte aligned");
    B200_LAUNCH(k_eltwise_sum_extelem)<<<grid_for(count, 256, 8), 256, 0, (cudaStream_t)stream>>>(d_out, d_in, count, to_add);
    DONE("eltwise_sum_extelem");
}

const char* b200_eltwise_add_elem(uint32_t* d_out, const uint32_t* d_a, const uint32_t* d_b, size_t count, void* stream) {
    NEED_GPU();
    if (count == 0) return nullptr;
    B200_LAUNCH(k_eltwise_add_elem)<<<grid_for(count, 256, 8), 256, 0, (cudaStream_t)stream>>>(d_out, d_a, d_b, count);
    DONE("eltwise_add_elem");
}

const char* b200_eltwise_copy_elem(uint32_t* d_out, const uint32_t* d_in, size_t count, void* stream) {
    NEED_GPU();
    if (count == 0) return nullptr;
    cudaError_t e = cudaMemcpyAsync(d_out, d_in, count * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) FAIL("b200_eltwise_copy_elem: %s", cudaGetErrorString(e));
    return nullptr;
}

const char* b200_eltwise_zeroize_elem(uint32_t* d_io, size_t count, void* stream) {
    NEED_GPU();
    if (count == 0) return nullptr;
    B200_LAUNCH(k_eltwise_zeroize_elem)<<<grid_for(count, 256, 8), 256, 0, (cudaStream_t)stream>>>(d_io, count);
    DONE("eltwise_zeroize_elem");
}

size_t b200_poly_divide_scratch_words(uint32_t size) { return poly_divide_scratch_words(size); }
const char* b200_poly_divide(uint32_t* d_poly, uint32_t size, uint32_t* d_remainder, const uint32_t* d_pow, uint32_t* d_scratch,
                             void* stream) {
    NEED_GPU();
    if (!aligned16(d_poly) || !aligned16(d_remainder) || !aligned16(d_pow) || !aligned16(d_scratch))
        FAIL("b200_poly_divide: ExtElem buffers must be 16-byte aligned");
    cudaError_t e = launch_poly_divide(d_poly, size, d_remainder, d_pow, d_scratch, (cudaStream_t)stream);
    if (e != cudaSuccess) FAIL("b200_poly_divide: %s", cudaGetErrorString(e));
    return nullptr;
}

size_t b200_prefix_products_scratch_words(uint32_t count) { return (size_t)8 * ((count + PP_CH - 1) / PP_CH) + 8; }
const char* b200_prefix_products(uint32_t* d_io, uint32_t count, uint32_t* d_scratch, void* stream) {
    NEED_GPU();
    if (count == 0) return nullptr;
    if (!aligned16(d_io) || !aligned16(d_scratch)) FAIL("b200_prefix_products: ExtElem buffers must be 16-byte aligned");
    const uint32_t nchunks = (count + PP_CH - 1) / PP_CH;
    uint32_t* prods = d_scratch;
    uint32_t* carry = d_scratch + (size_t)4 * nchunks;
    cudaStream_t s = (cudaStream_t)stream;
    B200_LAUNCH(k_pp_chunk_prod)<<<nchunks, PP_T, 0, s>>>(prods, d_io, count);
    B200_LAUNCH(k_pp_scan)<<<1, 256, 0, s>>>(carry, prods, nchunks);
    B200_LAUNCH(k_pp_apply)<<<nchunks, PP_T, 0, s>>>(d_io, carry, count);
    DONE("prefix_products");
}

const char* b200_gather_sample(uint32_t* d_dst, const uint32_t* d_src, size_t idx, uint32_t size, size_t stride, void* stream) {
    NEED_GPU();
    if (size == 0) return nullptr;
    B200_LAUNCH(k_gather_sample)<<<grid_for(size, 128, 8), 128, 0, (cudaStream_t)stream>>>(d_dst, d_src, idx, size, stride);
    DONE("gather_sample");
}

const char* b200_scatter(uint32_t* d_into, const uint32_t* d_index, uint32_t n_index, const uint32_t* d_offsets, const uint32_t* d_values,
                         uint32_t n_values, void* stream) {
    NEED_GPU();
    if (n_index < 2 || n_values == 0) return nullptr;
    B200_LAUNCH(k_scatter)<<<grid_for(n_values, 256, 8), 256, 0, (cudaStream_t)stream>>>(d_into, d_index, n_index, d_offsets, d_values);
    DONE("scatter");
}

size_t b200_merkle_open_words(uint32_t lg_rows, uint32_t cols, uint32_t top_size) {
    uint32_t lg_top = 0;
    while ((1u << lg_top) < top_size) lg_top++;
    return (size_t)cols + (lg_rows > lg_top ? (size_t)(lg_rows - lg_top) * 8 : 0);
}
const char* b200_merkle_open(uint32_t* d_out, const uint32_t* d_nodes, const uint32_t* d_matrix, uint32_t lg_rows, uint32_t cols,
                             uint32_t top_size, uint32_t idx, void* stream) {
    NEED_GPU();
    if (lg_rows > 26) FAIL("b200_merkle_open: lg_rows too large");
    const uint32_t rows = 1u << lg_rows;
    if (idx >= rows) FAIL("b200_merkle_open: index %u out of range (%u rows)", idx, rows);
    if (top_size == 0 || (top_size & (top_size - 1)) || top_size > rows) FAIL("b200_merkle_open: top_size must be a power of two <= rows");
    B200_LAUNCH(k_merkle_open)<<<1, 128, 0, (cudaStream_t)stream>>>(d_out, d_nodes, d_matrix, rows, cols, top_size, idx);
    DONE("merkle_open");
}

// PolyGroup::new: K1+K2 (fused), K3, K4+K5
const char* b200_commit_group(uint32_t* d_coeffs_io, uint32_t* d_evals, uint32_t* d_nodes, uint32_t lg_n, uint32_t count, void* stream) {
    const char* e = b200_batch_intt_zk_shift(d_coeffs_io, lg_n, count, stream);
    if (e) return e;
    e = b200_batch_expand_ntt(d_evals, d_coeffs_io, lg_n, INV_RATE_LG, count, stream);
    if (e) return e;
    return b200_merkle_tree(d_nodes, d_evals, lg_n + INV_RATE_LG, count, stream);
}

}  // extern "C"
