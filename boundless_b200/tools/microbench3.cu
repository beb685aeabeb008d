// Is the FP64 pipe of B200 a usable SECOND modular multiplier next to the integer pipe?
// Measures: raw DFMA rate; FP64 lazy mulmod (a*b mod p with error-free product, 7 FP64 ops) rate; and the combined rate when
// half / a third of the warps of every CTA run the FP64 formulation while the others run the integer Montgomery multiply.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench3 microbench3.cu
#include <cstdio>
#include <cstdlib>
#include "../csrc/field.cuh"

using namespace b200;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double dmulmod(double a, double b) {
    const double PD = 2013265921.0, PINV = 1.0 / 2013265921.0, MAGIC = 6755399441055744.0;   // 1.5 * 2^52
    double h = a * b;
    double l = fma(a, b, -h);
    double q = (h * PINV + MAGIC) - MAGIC;
    double r = fma(-q, PD, h);
    return r + l;
}

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters) {
    double a[8], b = 1.000000001, c = 0.5;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 0: all warps integer; 1: all warps fp64; 2: warps with (warp % 2 == 1) fp64; 3: (warp % 3 == 2) fp64; 4: (warp % 4 == 3) fp64
template <int MODE>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, unsigned long long* counts) {
    const int warp = threadIdx.x >> 5;
    const bool fp = MODE == 1 || (MODE == 2 && (warp & 1)) || (MODE == 3 && (warp % 3 == 2)) || (MODE == 4 && (warp & 3) == 3);
    double res = 0;
    if (fp) {
        double a[8], b[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { a[i] = (double)((threadIdx.x * 977u + i * 131u + 5u) % P); b[i] = (double)((threadIdx.x * 31u + i * 7u + 11u) % P); }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = dmulmod(a[i], b[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) res += a[i];
    } else {
        uint32_t a[8], b[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { a[i] = (threadIdx.x * 977u + i * 131u + 5u) % P; b[i] = (threadIdx.x * 31u + i * 7u + 11u) % P; }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = fp_mul(a[i], b[i]);
        }
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s ^= a[i];
        res = (double)s;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = res;
    (void)counts;
}

// correctness of dmulmod against the integer path
__global__ void k_check(int* bad) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u % P, y = (x * 40503u + 12345u) % P;
    double r = dmulmod((double)x, (double)y);
    long long ri = (long long)r; ri %= (long long)P; if (ri < 0) ri += P;
    unsigned long long ref = (unsigned long long)x * y % P;
    if ((unsigned long long)ri != ref) atomicAdd(bad, 1);
    // lazy inputs: values up to 2^40 in magnitude
    double xa = (double)x * 300.0 - 1e11, ya = (double)y * 250.0 + 3e10;
    double r2 = dmulmod(xa, ya);
    long long xi = (long long)xa % (long long)P, yi = (long long)ya % (long long)P;
    if (xi < 0) xi += P; if (yi < 0) yi += P;
    unsigned long long ref2 = (unsigned long long)xi * (unsigned long long)yi % P;
    long long r2i = (long long)r2 % (long long)P; if (r2i < 0) r2i += P;
    if ((unsigned long long)r2i != ref2 || fabs(r2) > 1.7 * 2013265921.0) atomicAdd(bad, 1);
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* d_out; CK(cudaMalloc(&d_out, 64u << 20));
    int* d_bad; CK(cudaMalloc(&d_bad, 4)); CK(cudaMemset(d_bad, 0, 4));
    k_check<<<4096, 256>>>(d_bad); CK(cudaDeviceSynchronize());
    int bad; CK(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
    printf("dmulmod check: %d mismatches of %d\n", bad, 2 * 4096 * 256);
    int blocks = sms * 8, iters = 1000;
    double per_thread = (double)iters * 32;
    {
        float ms = time_ms([&] { k_dfma<<<blocks, 256>>>(d_out, iters); });
        double ops = (double)blocks * 256 * per_thread;
        printf("dfma raw      %8.3f ms  %8.2f Gop/s  %6.2f lane-op/clk/SM\n", ms, ops / ms * 1e-6, ops / (ms * 1e-3) / sms / (prop.clockRate * 1e3));
    }
    const char* names[] = {"all int", "all fp64", "1/2 warps fp64", "1/3 warps fp64", "1/4 warps fp64"};
#define RUN(M) { float ms = time_ms([&] { k_mix<M><<<blocks, 256>>>(d_out, iters, nullptr); }); \
        double total = (double)blocks * 256 * per_thread; \
        printf("mix %-16s %8.3f ms  %8.2f Gmulmod/s total\n", names[M], ms, total / ms * 1e-6); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
    return 0;
}
