// INT32-pipe roofline microbenchmark for the Poseidon2 / NTT kernels (B200, sm_100a).
// Measures (a) raw issue rates of the instructions a Montgomery multiply is made of, (b) Montgomery
// mulmod/s with independent chains, (c) Poseidon2 permutations/s with register-resident state.
// The numbers it prints are the denominators DESIGN.md uses for the integer roofline of K4/K5.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench microbench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../csrc/poseidon2.cuh"

using namespace b200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int OP>
__global__ void __launch_bounds__(256) k_raw(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], b = seed | 1u, c = seed * 3u + 7u;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 8u + i + seed;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                else if (OP == 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                else if (OP == 2) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(b)); a[i] = (uint32_t)w ^ (uint32_t)(w >> 32); }
                else if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                else if (OP == 4) a[i] = __viaddmin_u32(a[i], b, c);
                else if (OP == 5) asm volatile("mad.lo.u32 %0, %0, 0x88000001, %1;" : "+r"(a[i]) : "r"(c));
                else if (OP == 6) asm volatile("mul.hi.u32 %0, %0, 0x78000001;" : "+r"(a[i]));
                else if (OP == 7) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
                else if (OP == 8) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_mulmod(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = (threadIdx.x * 977u + i * 131u + seed) % P; b[i] = (a[i] * 3u + 11u) % P; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fp_mul(a[i], b[i]);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_addmod(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = (threadIdx.x * 977u + i * 131u + seed) % P; b[i] = (a[i] * 3u + 11u) % P; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fp_add(a[i], b[i]);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_perm(uint32_t* out, int iters, uint32_t seed) {
    uint32_t st[24];
#pragma unroll
    for (int i = 0; i < 24; i++) st[i] = (uint32_t)(((uint64_t)(blockIdx.x * THREADS + threadIdx.x) * 24u + i + seed) % P);
    for (int it = 0; it < iters; it++) p2_permute(st);
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 24; i++) s ^= st[i];
    out[blockIdx.x * THREADS + threadIdx.x] = s;
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_perm2(uint32_t* out, int iters, uint32_t seed) {
    uint32_t sa[24], sb[24];
#pragma unroll
    for (int i = 0; i < 24; i++) { sa[i] = (uint32_t)(((uint64_t)(blockIdx.x * THREADS + threadIdx.x) * 48u + i + seed) % P); sb[i] = (sa[i] + 24u) % P; }
    for (int it = 0; it < iters; it++) p2_permute2(sa, sb);
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 24; i++) s ^= sa[i] ^ sb[i];
    out[blockIdx.x * THREADS + threadIdx.x] = s;
}
__global__ void k_kat2(uint32_t* out) {
    uint32_t a[24], b[24];
    for (int i = 0; i < 24; i++) { a[i] = fp_to_mont(i); b[i] = fp_to_mont(i); }
    p2_permute2(a, b);
    for (int i = 0; i < 24; i++) { out[i] = fp_from_mont(a[i]); out[24 + i] = fp_from_mont(b[i]); }
}

__global__ void k_kat(uint32_t* out) {
    uint32_t st[24];
    for (int i = 0; i < 24; i++) st[i] = fp_to_mont(i);
    p2_permute(st);
    for (int i = 0; i < 24; i++) out[i] = fp_from_mont(st[i]);
}

// the differential check of tests/host_emul/device_on_host.cpp ("perms 12345 2000") on the device: 3 x 2002 chained permutations of
// pseudo-random / all-zero / all-(p-1) states folded into one 64-bit word; every build variant must print the same word as the host
__global__ void k_perms_fold(unsigned long long* out, unsigned long long seed, int count) {
    unsigned long long fold = 0;
    uint32_t st[24];
    for (int k = 0; k < count + 2; k++) {
        for (int i = 0; i < 24; i++) {
            seed = seed * 6364136223846793005ull + 1442695040888963407ull;
            st[i] = k == 0 ? 0u : k == 1 ? P - 1 : (uint32_t)((seed >> 33) % P);
        }
        for (int rep = 0; rep < 3; rep++) {
            p2_permute(st);
            for (int i = 0; i < 24; i++) fold = (fold ^ st[i]) * 1099511628211ull + (unsigned long long)i;
        }
    }
    *out = fold;
}

static const uint32_t KAT[24] = {
    0x2ed3e23d, 0x12921fb0, 0x0e659e79, 0x61d81dc9, 0x32bae33b, 0x62486ae3, 0x1e681b60, 0x24b91325,
    0x2a2ef5b9, 0x50e8593e, 0x5bc818ec, 0x10691997, 0x35a14520, 0x2ba6a3c5, 0x279d47ec, 0x55014e81,
    0x5953a67f, 0x2f403111, 0x6b8828ff, 0x1801301f, 0x2749207a, 0x3dc9cf21, 0x3c985ba2, 0x57a99864};

template <typename F>
static float time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms %d clock_khz %d\n", prop.name, sms, prop.clockRate);
    uint32_t* d_out; CK(cudaMalloc(&d_out, 64u << 20));
    // KAT
    k_kat<<<1, 1>>>(d_out); CK(cudaDeviceSynchronize());
    uint32_t h[24]; CK(cudaMemcpy(h, d_out, 96, cudaMemcpyDeviceToHost));
    int ok = 1; for (int i = 0; i < 24; i++) ok &= (h[i] == KAT[i]);
    printf("poseidon2_kat %s\n", ok ? "PASS" : "FAIL");
    k_kat2<<<1, 1>>>(d_out); CK(cudaDeviceSynchronize());
    uint32_t h2[48]; CK(cudaMemcpy(h2, d_out, 192, cudaMemcpyDeviceToHost));
    int ok2 = 1; for (int i = 0; i < 24; i++) ok2 &= (h2[i] == KAT[i]) && (h2[24 + i] == KAT[i]);
    printf("poseidon2_kat2 (two interleaved states) %s\n", ok2 ? "PASS" : "FAIL");

    {
        k_perms_fold<<<1, 1>>>((unsigned long long*)d_out, 12345ull, 2000); CK(cudaDeviceSynchronize());
        unsigned long long f; CK(cudaMemcpy(&f, d_out, 8, cudaMemcpyDeviceToHost));
        const int okf = f == 0xa3c709b7e26b094dull;       // value printed by the host build of the same headers
        printf("poseidon2_perms_fold %016llx %s\n", f, okf ? "PASS" : "FAIL");
        ok &= okf;
    }
#ifdef B200_MB_QUICK      // variant sweeps: the KAT and three launch shapes of the permutation only
    {
        int pit = 64;
#define PERMQ(T, MB, BPS) { int nb = sms * BPS; float ms = time_ms([&] { k_perm<T, MB><<<nb, T>>>(d_out, pit, 7u); }); \
            double perms = (double)nb * T * pit; printf("perm threads=%d minb=%d blocks/sm=%d %.3f Gperm/s\n", T, MB, BPS, perms / ms * 1e-6); }
        PERMQ(256, 3, 6) PERMQ(256, 2, 6) PERMQ(512, 1, 2) PERMQ(1024, 1, 1)
    }
    CK(cudaFree(d_out));
    return ok ? 0 : 1;
#endif
    const char* names[] = {"imad_lo_rrr", "imad_hi_rr", "imad_wide", "iadd", "viaddmin_u32", "imad_lo_imm", "imad_hi_imm", "lop3", "shf"};
    int blocks = sms * 8, iters = 2000;
    double ops = (double)blocks * 256 * iters * 32;
#define RAW(OPN) { float ms = time_ms([&] { k_raw<OPN><<<blocks, 256>>>(d_out, iters, 12345u); }); \
                   printf("raw %-14s %8.3f ms  %8.2f Gop/s  %6.2f lane-op/clk/SM @%d MHz\n", names[OPN], ms, ops / ms * 1e-6, ops / (ms * 1e-3) / sms / (prop.clockRate * 1e3), prop.clockRate / 1000); }
    RAW(0) RAW(1) RAW(2) RAW(3) RAW(4) RAW(5) RAW(6) RAW(7) RAW(8)
    {
        float ms = time_ms([&] { k_mulmod<<<blocks, 256>>>(d_out, iters, 1u); });
        printf("mulmod          %8.3f ms  %8.2f Gmulmod/s\n", ms, ops / ms * 1e-6);
        ms = time_ms([&] { k_addmod<<<blocks, 256>>>(d_out, iters, 1u); });
        printf("addmod          %8.3f ms  %8.2f Gaddmod/s\n", ms, ops / ms * 1e-6);
    }
    {
        int pit = 64;
#define PERM(T, MB, BPS) { int nb = sms * BPS; float ms = time_ms([&] { k_perm<T, MB><<<nb, T>>>(d_out, pit, 7u); }); \
            double perms = (double)nb * T * pit; \
            printf("perm threads=%d minb=%d blocks/sm=%d  %8.3f ms  %8.3f Gperm/s  %8.2f Gmulmod/s-equiv\n", T, MB, BPS, ms, perms / ms * 1e-6, perms * 1356 / ms * 1e-6); }
        PERM(128, 1, 4) PERM(128, 1, 8) PERM(128, 1, 12) PERM(256, 1, 2) PERM(256, 1, 4) PERM(256, 2, 4) PERM(256, 3, 6) PERM(512, 1, 2) PERM(64, 1, 16) PERM(256, 4, 8) PERM(128, 8, 16) PERM(1024, 1, 1) PERM(384, 2, 4)
    }
    {
        int pit = 64;
#define PERM2(T, MB, BPS) { int nb = sms * BPS; float ms = time_ms([&] { k_perm2<T, MB><<<nb, T>>>(d_out, pit, 7u); }); \
            double perms = 2.0 * nb * T * pit; \
            printf("perm2 threads=%d minb=%d blocks/sm=%d  %8.3f ms  %8.3f Gperm/s\n", T, MB, BPS, ms, perms / ms * 1e-6); }
        PERM2(128, 1, 4) PERM2(128, 2, 8) PERM2(256, 1, 2) PERM2(256, 2, 4) PERM2(512, 1, 2) PERM2(64, 4, 16) PERM2(128, 4, 8)
    }
    CK(cudaFree(d_out));
    return 0;
}
