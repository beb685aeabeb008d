// Test harness only: compiles the PRODUCT's device arithmetic headers (csrc/field.cuh, csrc/poseidon2.cuh) for the host by defining the
// CUDA qualifiers away and emulating the three intrinsics they use, so that the CPU test-suite exercises the very source the GPU runs
// (field operations against big-integer arithmetic done by the caller, the Poseidon2 permutation against the known-answer vector).
// This is not a CPU path of the product: nothing under boundless_b200/ builds or loads it.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define B200_HOST_EMULATION 1       // field.cuh then skips <cuda_runtime.h>: only the qualifiers and a few vector types are needed
#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline
#define __restrict__
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint32_t __viaddmin_u32(uint32_t a, uint32_t b, uint32_t c) { uint32_t s = a + b; return s < c ? s : c; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __brev(uint32_t x) { uint32_t r = 0; for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i); return r; }

#include "../../boundless_b200/csrc/field.cuh"
#include "../../boundless_b200/csrc/poseidon2.cuh"

using namespace b200;

// usage: device_on_host kat            -> 24 canonical words of permute([0..24))
//        device_on_host mul a b        -> from_mont(mul(to_mont(a), to_mont(b)))   (canonical integers in, canonical out)
//        device_on_host ops a b        -> add sub neg dbl mul sqr pow(a, b) on Montgomery forms, printed canonical
//        device_on_host fp4 a0..a3 b0..b3 -> the 4 canonical words of the extension-field product
int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "kat")) {
        uint32_t st[24];
        for (int i = 0; i < 24; i++) st[i] = fp_to_mont((uint32_t)i);
        p2_permute(st);
        for (int i = 0; i < 24; i++) printf("%08x%c", fp_from_mont(st[i]), i == 23 ? '\n' : ' ');
        return 0;
    }
    if (argc >= 4 && !strcmp(argv[1], "ops")) {
        const uint32_t a = (uint32_t)strtoul(argv[2], 0, 10), b = (uint32_t)strtoul(argv[3], 0, 10);
        const uint32_t am = fp_to_mont(a), bm = fp_to_mont(b);
        printf("%u %u %u %u %u %u %u\n", fp_from_mont(fp_add(am, bm)), fp_from_mont(fp_sub(am, bm)), fp_from_mont(fp_neg(am)),
               fp_from_mont(fp_dbl(am)), fp_from_mont(fp_mul(am, bm)), fp_from_mont(fp_sqr(am)), fp_from_mont(fp_pow(am, b)));
        return 0;
    }
    if (argc >= 10 && !strcmp(argv[1], "fp4")) {
        Fp4 a, b;
        for (int i = 0; i < 4; i++) { a.c[i] = fp_to_mont((uint32_t)strtoul(argv[2 + i], 0, 10)); b.c[i] = fp_to_mont((uint32_t)strtoul(argv[6 + i], 0, 10)); }
        Fp4 r = fp4_mul(a, b);
        printf("%u %u %u %u\n", fp_from_mont(r.c[0]), fp_from_mont(r.c[1]), fp_from_mont(r.c[2]), fp_from_mont(r.c[3]));
        return 0;
    }
    if (argc >= 4 && !strcmp(argv[1], "perms")) {
        // differential form: `count` permutations of pseudo-random states (plus the all-0 and all-(p-1) states), chained, and one
        // 64-bit fold of every output word -- two builds of the header that agree here agree on ~24*count field elements
        uint64_t seed = strtoull(argv[2], 0, 10), fold = 0;
        const int count = atoi(argv[3]);
        uint32_t st[24];
        for (int k = 0; k < count + 2; k++) {
            for (int i = 0; i < 24; i++) {
                seed = seed * 6364136223846793005ull + 1442695040888963407ull;
                st[i] = k == 0 ? 0u : k == 1 ? P - 1 : (uint32_t)((seed >> 33) % P);      // canonical Montgomery words
            }
            for (int rep = 0; rep < 3; rep++) {
                p2_permute(st);
                for (int i = 0; i < 24; i++) {
                    if (st[i] >= P) { printf("non-canonical output\n"); return 1; }
                    fold = (fold ^ st[i]) * 1099511628211ull + (uint64_t)i;
                }
            }
        }
        printf("%016llx\n", (unsigned long long)fold);
        return 0;
    }
    fprintf(stderr, "usage: device_on_host kat | ops a b | fp4 a0 a1 a2 a3 b0 b1 b2 b3 | perms seed count\n");
    return 2;
}
