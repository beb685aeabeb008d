/*
 * b200_risc0_sys_compat.h -- the extern "C" symbols of risc0-sys 1.5.0 / sppark 0.1.14 that risc0-zkp 3.0.3's CUDA `Hal`
 * binds, exported by libb200zkp.so UNDER THEIR ORIGINAL NAMES so that the Rust side links against this library with no
 * source change (the b200_* entry points of b200zkp.h stay the native, stream-aware interface).
 *
 * The crates are crates.io dependencies of /root/reference that are not vendored (Cargo.lock: risc0-sys 1.5.0, sppark 0.1.14,
 * risc0-zkp 3.0.3; SURVEY.md 0 finding 1, 8b "Kernel-level API"), so these prototypes are restated from the published crate
 * sources [RECALL-hi in SURVEY.md 8b]; the call sites inside the reference are
 *   prover/crates/workflow/src/tasks/prove.rs:44-52 (prove_segment), :96-104 (lift), tasks/join.rs:52-56 (join),
 * which reach them through risc0_zkvm::ProverServer -> risc0-zkp hal::cuda::CudaHal:
 *   batch_interpolate_ntt            -> sppark_batch_iNTT
 *   zk_shift                         -> sppark_batch_zk_shift
 *   batch_expand_into_evaluate_ntt   -> sppark_batch_expand, then sppark_batch_NTT on the expanded buffer
 *   batch_evaluate_ntt               -> sppark_batch_NTT
 *   hash_rows / hash_fold (Poseidon2)-> sppark_poseidon2_rows / sppark_poseidon2_fold
 *   poly_divide                      -> supra_poly_divide
 *
 * Conventions (sppark's): every function returns `sppark::Error { code: i32, message: *mut c_char }` BY VALUE; code 0 = success
 * and message NULL, otherwise message is malloc()ed and the caller frees it (the Rust `Drop` of sppark::Error calls free()).
 * All calls are synchronous on the legacy default stream, as the originals are.  Buffers are raw device pointers owned by the
 * caller: u32 BabyBear Montgomery words, column-major; a digest is 8 words; an ExtElem is 4 words.
 * There is no CPU path: without a CUDA device every call returns a non-zero code and a message.
 */
#ifndef B200_RISC0_SYS_COMPAT_H
#define B200_RISC0_SYS_COMPAT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t code;     /* 0 = ok; otherwise a cudaError_t value, or -1 for argument errors */
    char* message;    /* NULL on success; malloc()ed otherwise (caller frees) */
} sppark_error;

/* builds the twiddle tables of the current device (idempotent) */
sppark_error sppark_init(void);

/* `poly_count` in-place iNTTs of size 2^lg_domain_size: natural-order evaluations -> bit-reversed coefficients, scaled by 1/n */
sppark_error sppark_batch_iNTT(uint32_t* d_inout, uint32_t lg_domain_size, uint32_t poly_count);
/* `poly_count` in-place forward NTTs: bit-reversed coefficients -> natural-order evaluations */
sppark_error sppark_batch_NTT(uint32_t* d_inout, uint32_t lg_domain_size, uint32_t poly_count);
/* coefficient of x^d *= 3^d, coefficients in bit-reversed order */
sppark_error sppark_batch_zk_shift(uint32_t* d_inout, uint32_t lg_domain_size, uint32_t poly_count);
/* low-degree-extension spread: polynomial c of d_in (2^lg_domain_size bit-reversed coefficients) becomes polynomial c of d_out
 * (2^(lg_domain_size+lg_blowup) bit-reversed coefficients, the new high-degree coefficients zero), so that a following
 * sppark_batch_NTT(d_out, lg_domain_size + lg_blowup, poly_count) yields the evaluations over the blown-up domain.
 * (b200_batch_expand_ntt of b200zkp.h does both in one fused pass and is what a patched Hal should call.) */
sppark_error sppark_batch_expand(uint32_t* d_out, const uint32_t* d_in, uint32_t lg_domain_size, uint32_t lg_blowup,
                                 uint32_t poly_count);

/* d_out[count][8]: leaf j = Poseidon2 sponge over d_in[c*count + j], c < col_size */
sppark_error sppark_poseidon2_rows(uint32_t* d_out, const uint32_t* d_in, uint32_t count, uint32_t col_size);
/* d_out[i] = hash_pair(d_in[2i], d_in[2i+1]), i < num_hashes */
sppark_error sppark_poseidon2_fold(uint32_t* d_out, const uint32_t* d_in, size_t num_hashes);

/* ExtElem polynomial on the device (natural coefficient order) /= (x - *pow) in place; *remainder = P(*pow).
 * `remainder` and `pow` are HOST pointers to one ExtElem (4 Montgomery words) each, as in the original. */
sppark_error supra_poly_divide(uint32_t* d_polynomial, size_t poly_size, uint32_t* remainder, const uint32_t* pow);

#ifdef __cplusplus
}
#endif
#endif
