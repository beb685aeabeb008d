// Internal (non-ABI) declarations shared by the translation units of libb200zkp.so.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <cuda_runtime.h>
#include <atomic>
#include <vector>

namespace b200 {

// every kernel launch of the library goes through B200_LAUNCH so that b200_kernel_launches() reports real launches
extern std::atomic<uint64_t> g_kernel_launches;
template <typename F>
inline F* count_launch(F* f) { g_kernel_launches.fetch_add(1, std::memory_order_relaxed); return f; }
#define B200_LAUNCH(...) ::b200::count_launch(__VA_ARGS__)

constexpr int MAX_LG = 26;          // largest transform size 2^26 (segment po2 24 x blow-up 4; upstream MAX_CYCLES_PO2 = 24)
constexpr int MAX_LG_2PASS = 24;    // up to here a transform is two HBM passes; 2^25 and 2^26 take a third (csrc/ntt.cu)
constexpr int BIG_N1 = 10;          // rows of the outer strided pass of the three-pass transforms
constexpr int QUERIES = 50;
constexpr int INV_RATE_LG = 2;      // blow-up 4
constexpr int FRI_FOLD = 16;
constexpr int FRI_MIN_DEGREE = 256;
constexpr int CHECK_COLS = 16;
constexpr int GLOBALS = 16;

// ---- per-device read-only tables (tables.cpp) --------------------------------------------------------
struct DeviceTables {
    int device;
    int sm_count;
    // stage twiddles, compact per-level layout: tw[2^(l-1) + i] = w_{2^l}^i, i < 2^(l-1), l <= 12
    // every twiddle table holds Shoup pairs {w, floor(w * 2^32 / p)} with w the PLAIN value (csrc/ntt.cu mul_tw)
    uint2* tw_fwd;         // 8192 pairs
    uint2* tw_inv;         // 8192 pairs
    // six-step inter-pass twiddle decomposition for transform size 2^m:  w_{2^m}^e = lo[e & (2^h-1)] * hi[e >> h], h = ceil(m/2)
    uint2* pow_fwd[MAX_LG + 1];      // lo (2^h) followed by hi (2^(m-h))
    uint2* pow_inv[MAX_LG + 1];      // inverse roots; hi table pre-scaled by 2^-m (iNTT normalisation); for m > MAX_LG_2PASS by
                                     // 2^-BIG_N1 only: the inner transforms of the three-pass route bring their own 2^-(m - BIG_N1)
    // zk_shift: 3^d = p3lo[d & 4095] * p3hi[d >> 12]   (p3hi has 2^(MAX_LG - 12) entries)
    uint2* p3lo;
    uint2* p3hi;
    uint32_t rou_fwd[28], rou_rev[28];  // host copies, Montgomery
    // [r2] per-element tables in the DATA layout of a two-pass transform of size 2^m split as 2^full_rows[..][m] rows (get_full_table):
    // one coalesced 8-byte load and ONE multiply per element instead of two table gathers, an exponent and two multiplies.
    uint2* full[3][MAX_LG + 1];
    uint32_t full_rows[3][MAX_LG + 1];
};
// kind of a full table: element (rho, pos) of a 2^lg_rows x 2^(lg_m - lg_rows) matrix, index rho * 2^(lg_m - lg_rows) + pos, holds
//   FULL_INV / FULL_FWD: the inter-pass twiddle pow_inv / pow_fwd [lg_m] ^ (pos * bitrev(rho))      (inverse: normalisation folded in)
//   FULL_ZK:             3^(bitrev(rho) + 2^lg_rows * bitrev(pos)), the zk_shift factor of that slot of the bit-reversed coefficients
enum { FULL_INV = 0, FULL_FWD = 1, FULL_ZK = 2 };
// Built on first use (host-computed, synchronous upload; thread-safe), 8 * 2^lg_m bytes each; nullptr when lg_m exceeds
// B200_NTT_FULL_MAX_LG (default 22), when B200_NTT_FULL=0, on allocation failure, or when the table of this size exists with another
// split -- callers then use the two-table decomposition above.
const uint2* get_full_table(const DeviceTables* T, int kind, uint32_t lg_m, uint32_t lg_rows);
// the host computation behind it (Montgomery-form values; rou_* as in DeviceTables); used by get_full_table and by tests/host_emul
void fill_full_table(int kind, uint32_t lg_m, uint32_t lg_rows, const uint32_t* rou_fwd, const uint32_t* rou_rev, std::vector<uint32_t>& v);
const DeviceTables* get_tables(int device);   // lazily built, thread-safe; nullptr + error string on failure
void free_tables();                           // b200_shutdown
void compat_release();                        // b200_shutdown: scratch arena of the risc0-sys compatible supra_poly_divide (compat.cu)
const char* last_error();
void set_error(const char* fmt, ...);

// host-side field helpers (tables.cpp)
uint32_t h_mul(uint32_t a, uint32_t b);
uint32_t h_add(uint32_t a, uint32_t b);
uint32_t h_sub(uint32_t a, uint32_t b);
uint32_t h_pow(uint32_t a, uint64_t e);
uint32_t h_inv(uint32_t a);
uint32_t h_to_mont(uint32_t x);
uint32_t h_from_mont(uint32_t a);

// ---- NTT (ntt.cu) ------------------------------------------------------------------------------------
// K1: `count` in-place iNTTs of size 2^lg_n, natural-order evaluations -> bit-reversed coefficients, scaled by 2^-lg_n.
cudaError_t launch_batch_intt(const DeviceTables* T, uint32_t* d_io, uint32_t lg_n, uint32_t count, cudaStream_t s);
// forward, bit-reversed coefficients -> natural-order evaluations, in place
cudaError_t launch_batch_ntt(const DeviceTables* T, uint32_t* d_io, uint32_t lg_n, uint32_t count, cudaStream_t s);
// K3: out (count x 2^(lg_n+lg_blowup)) = NTT of the zero-padded (== replicated, levels skipped) coefficients
cudaError_t launch_batch_expand_ntt(const DeviceTables* T, uint32_t* d_out, const uint32_t* d_in, uint32_t lg_n,
                                    uint32_t lg_blowup, uint32_t count, cudaStream_t s);
// K2: coefficient of x^d *= 3^d (slot j holds degree bitrev(j))
cudaError_t launch_zk_shift(const DeviceTables* T, uint32_t* d_io, uint32_t lg_n, uint32_t count, cudaStream_t s);
cudaError_t launch_bit_reverse(uint32_t* d_io, uint32_t lg_n, uint32_t count, cudaStream_t s);
// K1 + K2 fused (iNTT whose last pass applies the coset shift)
cudaError_t launch_batch_intt_shift(const DeviceTables* T, uint32_t* d_io, uint32_t lg_n, uint32_t count, cudaStream_t s);

// ---- Poseidon2 / Merkle / transcript (hash.cu) -------------------------------------------------------
// K4: leaf j = sponge(matrix[c*col_stride + j], c < cols); out = rows x 8 words
cudaError_t launch_poseidon2_rows(uint32_t* d_out, const uint32_t* d_matrix, uint32_t rows, uint32_t cols,
                                  size_t col_stride, cudaStream_t s);
// K5: nodes[2*rows*8]; leaves in nodes[rows..2rows) -> fills nodes[1..rows)
cudaError_t launch_poseidon2_fold_tree(uint32_t* d_nodes, uint32_t lg_rows, cudaStream_t s);
// single-layer fold: out[i] = hash_pair(in[2i], in[2i+1]), i < n_out
cudaError_t launch_poseidon2_fold(uint32_t* d_out, const uint32_t* d_in, uint32_t n_out, cudaStream_t s);

struct Transcript { uint32_t cells[24]; uint32_t pool_used; uint32_t pad[7]; };
cudaError_t launch_iop_init(Transcript* t, cudaStream_t s);
cudaError_t launch_iop_commit(Transcript* t, const uint32_t* d_digest8, cudaStream_t s);
// hash_elem_slice(d_elems[0..count)) then rng.mix(digest); optional copy of the digest to d_digest_out
cudaError_t launch_iop_commit_elems(Transcript* t, const uint32_t* d_elems, uint32_t count, uint32_t* d_digest_out, cudaStream_t s);
cudaError_t launch_iop_draw_ext(Transcript* t, uint32_t* d_out, uint32_t n_ext, cudaStream_t s);
cudaError_t launch_iop_draw_bits(Transcript* t, uint32_t* d_out, uint32_t n, uint32_t bits, cudaStream_t s);
cudaError_t launch_hash_elems(uint32_t* d_digest_out, const uint32_t* d_elems, uint32_t count, cudaStream_t s);
cudaError_t launch_hash_pair_one(uint32_t* d_out8, const uint32_t* d_a8, const uint32_t* d_b8, cudaStream_t s);

// ---- STARK plumbing (stark.cu) -----------------------------------------------------------------------
cudaError_t launch_gen_trace(uint32_t* d_out, uint64_t seed, const uint32_t* d_seed_words, uint64_t count, cudaStream_t s);
cudaError_t launch_set_globals(uint32_t* d_seal, uint32_t po2, uint32_t w_code, uint32_t w_data, uint32_t w_accum, uint32_t kind, uint64_t seed, int hash_seed, cudaStream_t s);
cudaError_t launch_accumulate(uint32_t* d_acc_io, uint32_t rows, uint32_t w_accum, const uint32_t* d_mix, cudaStream_t s);
cudaError_t launch_powers(uint32_t* d_out, const uint32_t* d_base, uint32_t count, cudaStream_t s);   // out[k] = base^k (Fp4)
cudaError_t launch_eval_check(uint32_t* d_planes, const uint32_t* d_evals, uint32_t lg_domain, uint32_t w_code,
                              uint32_t w_data, uint32_t w_accum, const uint32_t* d_pmix, cudaStream_t s);
// K6
cudaError_t launch_fri_fold(uint32_t* d_out, const uint32_t* d_in, uint32_t in_size, const uint32_t* d_mix, cudaStream_t s);
// K7: evaluate `count` bit-reversed coefficient columns (size 2^lg_n) at Fp4 point d_x; cols [b0,b1) also at d_xb (may be null).
cudaError_t launch_evaluate(uint32_t* d_out_a, uint32_t* d_out_b, const uint32_t* d_coeffs, uint32_t lg_n, uint32_t count,
                            const uint32_t* d_x, const uint32_t* d_xb, uint32_t b0, uint32_t b1, uint32_t* d_scratch,
                            cudaStream_t s);
size_t evaluate_scratch_words(uint32_t lg_n, uint32_t count);
// derive the DEEP points from z: pts = [z, z * w_N^-1, z^4] (3 Fp4)
cudaError_t launch_deep_points(uint32_t* d_pts, const uint32_t* d_z, uint32_t rou_rev_n, cudaStream_t s);
// K8: DEEP combination + division + sum -> F planes (4 x N, bit-reversed)
struct DeepArgs {
    const uint32_t* coeffs;        // W columns x N (code, data, accum)
    const uint32_t* check_coeffs;  // 16 columns x N
    const uint32_t* u;             // T Fp4 tap evaluations
    const uint32_t* mix_pows;      // T Fp4 powers of the DEEP mix
    const uint32_t* pts;           // 3 Fp4 points
    uint32_t* combos;              // scratch: 3 x N Fp4 (AoS, natural degree order)
    uint32_t* chunk_vals;          // scratch: 3 x nchunks Fp4
    uint32_t* chunk_carry;         // scratch: 3 x nchunks Fp4
    uint32_t* f_planes;            // out: 4 x N
    uint32_t lg_n, W, w_accum;
};
cudaError_t launch_deep(const DeepArgs& a, cudaStream_t s);
// supra_poly_divide: Fp4 polynomial (AoS, natural coefficient order) /= (x - *d_pow) in place; *d_remainder = P(*d_pow)
size_t poly_divide_scratch_words(uint32_t size);
cudaError_t launch_poly_divide(uint32_t* d_poly, uint32_t size, uint32_t* d_remainder, const uint32_t* d_pow, uint32_t* d_scratch,
                               cudaStream_t s);
// K9: gather one Merkle opening per (query, tree) into the seal
struct GatherTree {
    const uint32_t* matrix; const uint32_t* nodes; uint32_t rows, cols, top_size; uint32_t seal_off;  // offset within a query record
    uint32_t pos_shift_mod;   // rows for "pos % rows" chaining (FRI) or 0 for trace groups (use pos as is)
};
cudaError_t launch_gather_queries(uint32_t* d_seal, uint32_t query_base, uint32_t query_words, const uint32_t* d_pos,
                                  const GatherTree* d_trees, uint32_t n_trees, cudaStream_t s);

// ---- seal verification (verify.cu) ----------------------------------------------------------------------------
// device context words: verdicts, DEEP sums and a scratch root
constexpr uint32_t VCTX_WORDS = 128, VCTX_RC = 0, VCTX_RESULT = 1, VCTX_QRC = 2 /* .. 2+QUERIES */, VCTX_USUM = 64 /* 3 Fp4 */, VCTX_ROOT = 80;
struct VerifyShape {
    uint32_t po2, w_code, w_data, w_accum, W, T, rounds, final_size, final_lg;
    uint32_t off_top[4], off_u, off_fri_top[8], off_final, off_queries, query_words, q_off_group[4], q_off_fri[8];
    uint32_t fri_rows[8], fri_top[8];
    uint32_t rou_fwd[28];      // Montgomery
    uint32_t inv16;            // Montgomery 1/16
};
cudaError_t launch_verify_reset(uint32_t* ctx, cudaStream_t s);
cudaError_t launch_verify_canonical(uint32_t* ctx, const uint32_t* seal, uint32_t words, cudaStream_t s);
// seal[0..5) must be the circuit the caller expects, seal[5..8) zero: else verdict 103
cudaError_t launch_verify_header(uint32_t* ctx, const uint32_t* seal, uint32_t po2, uint32_t w_code, uint32_t w_data, uint32_t w_accum,
                                 uint32_t kind, cudaStream_t s);
cudaError_t launch_verify_fold_top(uint32_t* root_out, const uint32_t* top, uint32_t top_size, cudaStream_t s);
cudaError_t launch_verify_constraint(uint32_t* ctx, const uint32_t* u, const uint32_t* pm, const uint32_t* z, uint32_t w_code,
                                     uint32_t w_data, uint32_t w_accum, cudaStream_t s);
cudaError_t launch_verify_usum(uint32_t* ctx, const uint32_t* u, const uint32_t* mp, uint32_t W, uint32_t w_accum, uint32_t T, cudaStream_t s);
cudaError_t launch_verify_queries(uint32_t* ctx, const uint32_t* seal, const VerifyShape& sh, const uint32_t* mp, const uint32_t* pts,
                                  const uint32_t* fmix, const uint32_t* pos, cudaStream_t s);
cudaError_t launch_verify_finish(uint32_t* ctx, cudaStream_t s);

}  // namespace b200
