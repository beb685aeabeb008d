// A/B harness: csrc/hash.cu alone as a tiny shared library, so that build variants of the Poseidon2 code (macros of csrc/poseidon2.cuh /
// csrc/field.cuh) can be timed in the REAL leaf-hash and fold kernels (tools/time_p2_variants.py), not only in the microbenchmark.
#include "../boundless_b200/csrc/hash.cu"
namespace b200 { std::atomic<uint64_t> g_kernel_launches{0}; }
extern "C" const char* p2_rows(uint32_t* out, const uint32_t* m, uint32_t rows, uint32_t cols, void* stream) {
    cudaError_t e = b200::launch_poseidon2_rows(out, m, rows, cols, rows, (cudaStream_t)stream);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
extern "C" const char* p2_fold(uint32_t* out, const uint32_t* in, uint32_t n_out, void* stream) {
    cudaError_t e = b200::launch_poseidon2_fold(out, in, n_out, (cudaStream_t)stream);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
