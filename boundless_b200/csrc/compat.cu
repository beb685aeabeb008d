// risc0-sys / sppark symbol-compatible exports (include/b200_risc0_sys_compat.h): the extern "C" names risc0-zkp 3.0.3's CUDA
// `Hal` binds (un-vendored crates.io dependencies of /root/reference: risc0-sys 1.5.0, sppark 0.1.14), implemented
// on the b200 kernels so the Rust side can link against libb200zkp.so unchanged.  Reached from
// /root/reference/prover/crates/workflow/src/tasks/prove.rs:44-52 (prove_segment) and :96-104 (lift) via ProverServer.
// Synchronous on the legacy stream like the originals; errors travel as sppark::Error {code, malloc()ed message} by value.
#include "../../include/b200_risc0_sys_compat.h"
#include "../../include/b200zkp.h"
#include "internal.h"
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace b200 {

// out[j << lg_b] = in[j], the other 2^lg_b - 1 slots of the group zero: zero padding of the high coefficients when both arrays
// are in bit-reversed order (bitrev_{n+b}(d) = bitrev_n(d) << b for d < 2^n).  One 16-byte store per input word for blow-up 4.
__global__ void __launch_bounds__(256) k_lde_spread4(uint4* __restrict__ out, const uint32_t* __restrict__ in, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        out[i] = make_uint4(__ldg(in + i), 0u, 0u, 0u);
}
__global__ void __launch_bounds__(256) k_lde_spread(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, size_t total_out,
                                                    uint32_t lg_b) {
    const size_t mask = ((size_t)1 << lg_b) - 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_out; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (i & mask) ? 0u : __ldg(in + (i >> lg_b));
}

}  // namespace b200

using namespace b200;

static sppark_error ok() { return sppark_error{0, nullptr}; }
static sppark_error fail(int32_t code, const char* msg) {
    const char* m = msg ? msg : "b200: unknown error";
    char* copy = (char*)malloc(strlen(m) + 1);
    if (copy) strcpy(copy, m);
    return sppark_error{code ? code : -1, copy};
}
// run a b200_* entry point on the legacy stream, then wait for it (the sppark originals return after the work is done)
static sppark_error finish(const char* err) {
    if (err) return fail(-1, err);
    cudaError_t e = cudaStreamSynchronize(nullptr);
    if (e != cudaSuccess) return fail((int32_t)e, cudaGetErrorString(e));
    return ok();
}
static bool no_device(sppark_error* out) {
    if (b200_device_count() > 0) return false;
    *out = fail(-1, "b200: no CUDA device available (this library has no CPU path)");
    return true;
}

// supra_poly_divide keeps one scratch arena per process (grown on demand), like the original's internal temporaries
static std::mutex g_div_mu;
static uint32_t* g_div_scratch = nullptr;
static size_t g_div_words = 0;

namespace b200 {
// b200_shutdown(): release the division arena
void compat_release() {
    std::lock_guard<std::mutex> lock(g_div_mu);
    if (g_div_scratch) cudaFree(g_div_scratch);
    g_div_scratch = nullptr; g_div_words = 0;
}
}  // namespace b200

extern "C" {

sppark_error sppark_init(void) {
    sppark_error r;
    if (no_device(&r)) return r;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail((int32_t)e, cudaGetErrorString(e));
    const char* err = b200_init(dev);
    return err ? fail(-1, err) : ok();
}

sppark_error sppark_batch_iNTT(uint32_t* d_inout, uint32_t lg_domain_size, uint32_t poly_count) {
    sppark_error r;
    if (no_device(&r)) return r;
    return finish(b200_batch_intt(d_inout, lg_domain_size, poly_count, nullptr));
}

sppark_error sppark_batch_NTT(uint32_t* d_inout, uint32_t lg_domain_size, uint32_t poly_count) {
    sppark_error r;
    if (no_device(&r)) return r;
    return finish(b200_batch_ntt(d_inout, lg_domain_size, poly_count, nullptr));
}

sppark_error sppark_batch_zk_shift(uint32_t* d_inout, uint32_t lg_domain_size, uint32_t poly_count) {
    sppark_error r;
    if (no_device(&r)) return r;
    return finish(b200_batch_zk_shift(d_inout, lg_domain_size, poly_count, nullptr));
}

sppark_error sppark_batch_expand(uint32_t* d_out, const uint32_t* d_in, uint32_t lg_domain_size, uint32_t lg_blowup,
                                 uint32_t poly_count) {
    sppark_error r;
    if (no_device(&r)) return r;
    if (lg_domain_size + lg_blowup > (uint32_t)MAX_LG) return fail(-1, "sppark_batch_expand: expanded domain larger than 2^26");
    if (poly_count == 0) return ok();
    const size_t total_in = (size_t)poly_count << lg_domain_size, total_out = total_in << lg_blowup;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t cap = (size_t)sms * 8;
    if (lg_blowup == 2 && ((uintptr_t)d_out & 15) == 0) {
        size_t g = (total_in + 255) / 256;
        B200_LAUNCH(k_lde_spread4)<<<(unsigned)(g < cap ? g : cap), 256>>>(reinterpret_cast<uint4*>(d_out), d_in, total_in);
    } else {
        size_t g = (total_out + 255) / 256;
        B200_LAUNCH(k_lde_spread)<<<(unsigned)(g < cap ? g : cap), 256>>>(d_out, d_in, total_out, lg_blowup);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int32_t)e, cudaGetErrorString(e));
    return finish(nullptr);
}

sppark_error sppark_poseidon2_rows(uint32_t* d_out, const uint32_t* d_in, uint32_t count, uint32_t col_size) {
    sppark_error r;
    if (no_device(&r)) return r;
    return finish(b200_poseidon2_rows(d_out, d_in, count, col_size, nullptr));
}

sppark_error sppark_poseidon2_fold(uint32_t* d_out, const uint32_t* d_in, size_t num_hashes) {
    sppark_error r;
    if (no_device(&r)) return r;
    if (num_hashes > 0xffffffffu) return fail(-1, "sppark_poseidon2_fold: num_hashes out of range");
    return finish(b200_poseidon2_fold(d_out, d_in, (uint32_t)num_hashes, nullptr));
}

sppark_error supra_poly_divide(uint32_t* d_polynomial, size_t poly_size, uint32_t* remainder, const uint32_t* pow) {
    sppark_error r;
    if (no_device(&r)) return r;
    if (!remainder || !pow) return fail(-1, "supra_poly_divide: null remainder / pow");
    if (poly_size > 0xffffffffu) return fail(-1, "supra_poly_divide: poly_size out of range");
    std::lock_guard<std::mutex> lock(g_div_mu);
    // arena: [0,4) pow, [4,8) remainder, [8, ..) the scan scratch of b200_poly_divide
    const size_t need = 8 + b200_poly_divide_scratch_words((uint32_t)poly_size);
    if (need > g_div_words) {
        if (g_div_scratch) cudaFree(g_div_scratch);
        g_div_scratch = nullptr; g_div_words = 0;
        cudaError_t e = cudaMalloc(&g_div_scratch, need * 4);
        if (e != cudaSuccess) return fail((int32_t)e, cudaGetErrorString(e));
        g_div_words = need;
    }
    cudaError_t e = cudaMemcpy(g_div_scratch, pow, 16, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail((int32_t)e, cudaGetErrorString(e));
    const char* err = b200_poly_divide(d_polynomial, (uint32_t)poly_size, g_div_scratch + 4, g_div_scratch, g_div_scratch + 8, nullptr);
    if (err) return fail(-1, err);
    e = cudaMemcpy(remainder, g_div_scratch + 4, 16, cudaMemcpyDeviceToHost);      // synchronises with the legacy stream
    if (e != cudaSuccess) return fail((int32_t)e, cudaGetErrorString(e));
    return ok();
}

}  // extern "C"
