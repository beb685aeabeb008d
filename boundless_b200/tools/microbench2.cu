// Pipe-overlap microbenchmark: cycles per iteration of small instruction mixes (8 independent chains per thread,
// 16 warps per SMSP), to learn which B200 pipes the Montgomery-multiply instructions share.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench2 microbench2.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MIX>
__global__ void __launch_bounds__(256) k_mix(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], b[8], c = seed * 3u + 7u, d = seed | 1u;
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 8u + i + seed; b[i] = a[i] * 2654435761u + 1u; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint64_t w;
                if (MIX == 1) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(d)); a[i] = (uint32_t)w; b[i] ^= (uint32_t)(w >> 32); }
                if (MIX == 2) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(d)); a[i] = (uint32_t)w; b[i] = __viaddmin_u32(b[i], (uint32_t)(w >> 32), c); }
                if (MIX == 3) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(d)); a[i] = __viaddmin_u32((uint32_t)w, c, d); b[i] = __viaddmin_u32(b[i], (uint32_t)(w >> 32), c); }
                if (MIX == 4) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(d)); a[i] = (uint32_t)w; b[i] = __viaddmin_u32(b[i], (uint32_t)(w >> 32), c); b[i] = __viaddmin_u32(b[i], d, c); b[i] = __viaddmin_u32(b[i], c, d); }
                if (MIX == 5) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(d), "r"(c)); b[i] = __viaddmin_u32(b[i], a[i], c); }
                if (MIX == 6) { asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(d), "r"(b[i])); }
                if (MIX == 7) { asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(d)); }
                if (MIX == 8) { a[i] = a[i] + (b[i] << 4); b[i] = b[i] + (a[i] << 27); }                      // 2 LEA
                if (MIX == 9) { a[i] = __viaddmin_u32(a[i], b[i], c); }                                          // 1 VIADDMNMX
                if (MIX == 10) { a[i] = __viaddmin_u32(a[i], b[i], c); b[i] = __viaddmin_u32(b[i], d, a[i]); }    // 2 dependent-ish
                if (MIX == 11) {   // full Montgomery multiply, original formulation
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(b[i]));
                    uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32), m, t;
                    asm volatile("mul.lo.u32 %0, %1, 0x88000001;" : "=r"(m) : "r"(lo));
                    asm volatile("mul.hi.u32 %0, %1, 0x78000001;" : "=r"(t) : "r"(m));
                    uint32_t r = hi - t;
                    a[i] = __viaddmin_u32(r, 0x78000001u, r);
                }
                if (MIX == 12) {   // multiply + one modular add (IADD + VIADDMNMX)
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(b[i]));
                    uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32), m, t;
                    asm volatile("mul.lo.u32 %0, %1, 0x88000001;" : "=r"(m) : "r"(lo));
                    asm volatile("mul.hi.u32 %0, %1, 0x78000001;" : "=r"(t) : "r"(m));
                    uint32_t r = hi - t;
                    a[i] = __viaddmin_u32(r, 0x78000001u, r);
                    uint32_t s = a[i] + c;
                    b[i] = __viaddmin_u32(s, 0x87ffffffu, s);
                }
                if (MIX == 13) {   // multiply + two modular adds
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(b[i]));
                    uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32), m, t;
                    asm volatile("mul.lo.u32 %0, %1, 0x88000001;" : "=r"(m) : "r"(lo));
                    asm volatile("mul.hi.u32 %0, %1, 0x78000001;" : "=r"(t) : "r"(m));
                    uint32_t r = hi - t;
                    a[i] = __viaddmin_u32(r, 0x78000001u, r);
                    uint32_t s = a[i] + c;
                    s = __viaddmin_u32(s, 0x87ffffffu, s);
                    uint32_t s2 = s + d;
                    b[i] = __viaddmin_u32(s2, 0x87ffffffu, s2);
                }
                if (MIX == 14) { uint32_t s = a[i] + b[i]; a[i] = __viaddmin_u32(s, 0x87ffffffu, s); }     // modular add alone
                if (MIX == 15) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(d), "r"(c)); b[i] = b[i] + a[i]; }   // IMAD + IADD
                if (MIX == 16) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(d)); a[i] = (uint32_t)w; b[i] = b[i] + (uint32_t)(w >> 32); b[i] = b[i] ^ c; }  // WIDE + IADD + LOP
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
    uint32_t* d_out; CK(cudaMalloc(&d_out, 64u << 20));
    int blocks = sms * 8, iters = 1000;
    double groups = (double)blocks * 256 / 32 * iters * 32;   // warp-level mix executions
    const char* names[] = {"", "WIDE", "WIDE+1VIADDMNMX", "WIDE+2VIADDMNMX", "WIDE+3VIADDMNMX", "IMAD+1VIADDMNMX", "IMAD.HI acc", "IMAD.HI",
                           "2xLEA", "1xVIADDMNMX", "2xVIADDMNMX", "mulmod", "mulmod+1add", "mulmod+2add", "addmod", "IMAD+IADD", "WIDE+IADD+LOP"};
#define MIXRUN(M) { float ms = time_ms([&] { k_mix<M><<<blocks, 256>>>(d_out, iters, 12345u); }); \
        double cyc = ms * 1e-3 * prop.clockRate * 1e3 * sms * 4 / groups; \
        printf("mix %2d %-18s %8.3f ms  %6.2f cycles/warp-iter/SMSP (at %d MHz)\n", M, names[M], ms, cyc, prop.clockRate / 1000); }
    MIXRUN(1) MIXRUN(2) MIXRUN(3) MIXRUN(4) MIXRUN(5) MIXRUN(6) MIXRUN(7) MIXRUN(8) MIXRUN(9) MIXRUN(10) MIXRUN(11) MIXRUN(12) MIXRUN(13) MIXRUN(14) MIXRUN(15) MIXRUN(16)
    return 0;
}
