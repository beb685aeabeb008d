/*
 * b200zkp.h -- C ABI of libb200zkp.so: B200-native (sm_100a) kernels and prover pipeline for the
 * segment-proving hot path of boundless-xyz/boundless (bento GPU agent -> risc0_zkvm::ProverServer).
 *
 * Every entry point returns `const char*`: NULL on success, otherwise a thread-local message (the same
 * convention as risc0-sys' C wrappers; SURVEY.md 8b "Errors").  Device buffers are caller-owned raw device
 * pointers; elements are u32 BabyBear values in Montgomery form; matrices are column-major (column c occupies
 * [c*rows, (c+1)*rows)); a digest is 8 Montgomery words.  `stream` is a cudaStream_t passed as void*.
 * No entry point falls back to the CPU: without a usable CUDA device every call fails with a message.
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   - trait risc0_zkvm::ProverServer obtained at prover/crates/workflow/src/lib.rs:276-284 and called at
 *     prover/crates/workflow/src/tasks/prove.rs:44-52 (prove_segment), :96-104 (lift),
 *     tasks/join.rs:52-56 (join), tasks/resolve.rs:84-88 (resolve), tasks/union.rs:43-47 (union);
 *   - one layer down, the extern "C" surface of risc0-sys 1.5.0 / sppark 0.1.14 (Cargo.lock:9131,10315;
 *     un-vendored): sppark_batch_iNTT / sppark_batch_NTT / sppark_batch_expand / sppark_batch_zk_shift /
 *     sppark_poseidon2_rows / sppark_poseidon2_fold / fri_fold / batch_evaluate_any.
 */
#ifndef B200ZKP_H
#define B200ZKP_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_DIGEST_WORDS 8
#define B200_QUERIES 50
#define B200_CHECK_COLS 16

/* ---- library ------------------------------------------------------------------------------------------- */
/* replaces sppark_init(): selects the device and builds the twiddle tables (idempotent). */
const char* b200_init(int device);
const char* b200_last_error(void);
/* number of CUDA devices visible, or -1 (never touches the CPU path) */
int b200_device_count(void);
/* PCI bus id of `device` ("0000:1b:00.0") into out[len >= 16]: lets the host side look up the GPU's NUMA node for its pinned buffers */
const char* b200_device_pci_bus_id(int device, char* out, int len);
/* frees the per-device twiddle tables built by b200_init (destroy every b200_prover first) */
const char* b200_shutdown(void);

/* ---- kernel 1: NTT (replaces sppark_batch_iNTT / _NTT / _expand / _zk_shift) ------------------------------ */
/* Transform sizes up to 2^26 (a po2-24 segment at blow-up 4; upstream MAX_CYCLES_PO2 = 24): two HBM passes up to 2^24, three above.
 * K1: `count` in-place iNTTs of size 2^lg_n: natural-order evaluations -> bit-reversed coefficients, x 2^-lg_n */
const char* b200_batch_intt(uint32_t* d_io, uint32_t lg_n, uint32_t count, void* stream);
/* forward: bit-reversed coefficients -> natural-order evaluations, in place */
const char* b200_batch_ntt(uint32_t* d_io, uint32_t lg_n, uint32_t count, void* stream);
/* K3: d_out[count][2^(lg_n+lg_blowup)] = evaluations of the zero-padded polynomials (expand + NTT, levels skipped) */
const char* b200_batch_expand_ntt(uint32_t* d_out, const uint32_t* d_in, uint32_t lg_n, uint32_t lg_blowup,
                                  uint32_t count, void* stream);
/* K2: coefficient of x^d *= 3^d (slot j holds degree bitrev(j)) */
const char* b200_batch_zk_shift(uint32_t* d_io, uint32_t lg_n, uint32_t count, void* stream);
const char* b200_batch_bit_reverse(uint32_t* d_io, uint32_t lg_n, uint32_t count, void* stream);
/* K1 + K2 fused: b200_batch_intt followed by b200_batch_zk_shift in one pass over the data (what commit_group uses) */
const char* b200_batch_intt_zk_shift(uint32_t* d_io, uint32_t lg_n, uint32_t count, void* stream);

/* ---- kernel 2: Poseidon2 (replaces sppark_poseidon2_rows / sppark_poseidon2_fold) ---------------------------- */
/* K4: d_out[rows][8]; leaf j = sponge over d_matrix[c*rows + j], c < cols */
const char* b200_poseidon2_rows(uint32_t* d_out, const uint32_t* d_matrix, uint32_t rows, uint32_t cols, void* stream);
/* K5 (one layer): d_out[i] = hash_pair(d_in[2i], d_in[2i+1]), i < num_hashes */
const char* b200_poseidon2_fold(uint32_t* d_out, const uint32_t* d_in, uint32_t num_hashes, void* stream);
/* K4+K5: whole tree. d_nodes holds 2*rows digests: leaves at [rows,2rows), node i = hash_pair(2i, 2i+1), root = node 1 */
const char* b200_merkle_tree(uint32_t* d_nodes, const uint32_t* d_matrix, uint32_t lg_rows, uint32_t cols, void* stream);

/* ---- kernel 3: FRI fold + evaluation (replaces fri_fold, batch_evaluate_any) ------------------------------------ */
/* K6: d_in = 4 planes x in_size bit-reversed coefficients, d_out = 4 planes x in_size/16, d_mix = one Fp4 (device) */
const char* b200_fri_fold(uint32_t* d_out, const uint32_t* d_in, uint32_t in_size, const uint32_t* d_mix, void* stream);
/* K7: d_out[count][4] = value of each bit-reversed coefficient column at the Fp4 point d_x (device) */
size_t b200_evaluate_scratch_words(uint32_t lg_n, uint32_t count);
const char* b200_batch_evaluate_any(uint32_t* d_out, const uint32_t* d_coeffs, uint32_t lg_n, uint32_t count,
                                    const uint32_t* d_x, uint32_t* d_scratch, void* stream);

/* ---- the small HAL operations (K8, K9 and element-wise helpers; replace the risc0-sys C wrappers of SURVEY.md 8b) -------- */
/* ExtElem buffers are arrays of 4 Montgomery words (16-byte aligned); Elem buffers are plain word arrays.
 * K8a Hal::mix_poly_coeffs: d_out[combos[i]*count + idx] += mix_start * mix^i * d_in[i*count + idx]   (i < input_size, idx < count)
 *     d_out = n_combos*count ExtElems (accumulated into), d_mix_start / d_mix = one ExtElem each (device), d_combos = input_size
 *     words, each < n_combos (other entries are skipped).  input_size <= 10240. */
const char* b200_mix_poly_coeffs(uint32_t* d_out, const uint32_t* d_mix_start, const uint32_t* d_mix, const uint32_t* d_in,
                                 const uint32_t* d_combos, uint32_t input_size, uint32_t count, uint32_t n_combos, void* stream);
/* K8b Hal::eltwise_sum_extelem: d_out[j*count + idx] = (sum_{i<to_add} d_in[i*count + idx]).elems[j]   (ExtElems in, 4 planes out) */
const char* b200_eltwise_sum_extelem(uint32_t* d_out, const uint32_t* d_in, uint32_t count, uint32_t to_add, void* stream);
/* K8c supra_poly_divide: ExtElem polynomial (natural coefficient order) /= (x - *d_pow) in place, *d_remainder = P(*d_pow).
 *     d_scratch = b200_poly_divide_scratch_words(size) words, 16-byte aligned. */
size_t b200_poly_divide_scratch_words(uint32_t size);
const char* b200_poly_divide(uint32_t* d_poly, uint32_t size, uint32_t* d_remainder, const uint32_t* d_pow, uint32_t* d_scratch,
                             void* stream);
/* Hal::prefix_products: d_io[i] = d_io[0] * ... * d_io[i] over `count` ExtElems (the accum grand product, X2) */
size_t b200_prefix_products_scratch_words(uint32_t count);
const char* b200_prefix_products(uint32_t* d_io, uint32_t count, uint32_t* d_scratch, void* stream);
/* Hal::eltwise_add_elem / eltwise_copy_elem / eltwise_zeroize_elem (zeroize: the INVALID marker 0xFFFFFFFF becomes 0) */
const char* b200_eltwise_add_elem(uint32_t* d_out, const uint32_t* d_a, const uint32_t* d_b, size_t count, void* stream);
const char* b200_eltwise_copy_elem(uint32_t* d_out, const uint32_t* d_in, size_t count, void* stream);
const char* b200_eltwise_zeroize_elem(uint32_t* d_io, size_t count, void* stream);
/* K9 Hal::gather_sample: d_dst[g] = d_src[g*stride + idx], g < size */
const char* b200_gather_sample(uint32_t* d_dst, const uint32_t* d_src, size_t idx, uint32_t size, size_t stride, void* stream);
/* Hal::scatter: d_into[d_offsets[k]] = d_values[k] for k in [d_index[0], d_index[n_index-1]); n_values bounds the launch */
const char* b200_scatter(uint32_t* d_into, const uint32_t* d_index, uint32_t n_index, const uint32_t* d_offsets,
                         const uint32_t* d_values, uint32_t n_values, void* stream);
/* K9 MerkleTreeProver::prove(idx): d_out = the `cols` values of row idx, then the sibling digests from the leaf layer up to (not
 *     including) the layer of top_size nodes; b200_merkle_open_words gives the length.  d_nodes as built by b200_merkle_tree. */
size_t b200_merkle_open_words(uint32_t lg_rows, uint32_t cols, uint32_t top_size);
const char* b200_merkle_open(uint32_t* d_out, const uint32_t* d_nodes, const uint32_t* d_matrix, uint32_t lg_rows, uint32_t cols,
                             uint32_t top_size, uint32_t idx, void* stream);
/* composite PolyGroup::new: d_coeffs_io = count columns of 2^lg_n evaluations -> shifted bit-reversed coefficients (K1+K2);
 *     d_evals = count x 2^(lg_n+2) evaluations (K3); d_nodes = 2 * 2^(lg_n+2) digests, root = node 1 (K4+K5) */
const char* b200_commit_group(uint32_t* d_coeffs_io, uint32_t* d_evals, uint32_t* d_nodes, uint32_t lg_n, uint32_t count, void* stream);

/* ---- prover pipeline (the ProverServer operator boundary) ------------------------------------------------------- */
typedef struct {
    uint32_t po2;                       /* trace rows = 2^po2, 9..24; reference default 20 (workflow/src/lib.rs:83-84) */
    uint32_t w_code, w_data, w_accum;   /* synthetic column-group widths; segment default 16/208/32 */
    uint32_t kind;                      /* 0 segment, 1 lift, 2 join, 3 resolve, 4 union */
} b200_circuit;

typedef struct b200_prover b200_prover;

size_t b200_seal_words(const b200_circuit* c);
/* one prover per GPU process (the agent's `Rc<dyn ProverServer>`); `slots` proofs may be in flight at once */
const char* b200_prover_create(b200_prover** out, int device, const b200_circuit* max_circuit, uint32_t slots);
void b200_prover_destroy(b200_prover* p);
size_t b200_prover_device_bytes(const b200_prover* p);

/* prove_segment: enqueue the whole proof on the slot's stream and return immediately.
 * h_trace (optional, may be NULL) = (w_code + w_data) x 2^po2 Montgomery words in host memory (pinned for overlap):
 * the segment's witness; when NULL the witgen stand-in expands `seed` on the device.
 * h_seal receives b200_seal_words(c) words once b200_prover_wait(p, slot) returns. */
const char* b200_prove_segment_async(b200_prover* p, uint32_t slot, const b200_circuit* c, uint64_t seed,
                                     const uint32_t* h_trace, uint32_t* h_seal);
/* Overlap the host -> device copy of the NEXT segment's witness with the proof the slot is running: copies h_trace (pinned) into the
 * slot's second coefficient region on a separate copy stream; the next b200_prove_segment_async on this slot that passes the SAME
 * h_trace pointer uses it without copying on its own stream.  May be called while the slot is busy; the caller must have waited for
 * the slot's previous proof but one (true for any submit / wait loop).  The region (W x 2^po2 words) is allocated on first use. */
const char* b200_prefetch_trace_async(b200_prover* p, uint32_t slot, const b200_circuit* c, const uint32_t* h_trace);
/* lift / join / resolve / union: recursion-shaped proof (kind 1..4) over the digest(s) of the child seal(s) */
const char* b200_recursion_async(b200_prover* p, uint32_t slot, const b200_circuit* c, const uint32_t* h_seal_a,
                                 size_t words_a, const uint32_t* h_seal_b, size_t words_b, uint32_t* h_seal);
/* The same with the child seals in DEVICE memory (receipts kept on the GPU between prove, lift and join, or received from a peer
 * GPU over NVLink): no host staging.  Read on the slot's stream; the caller keeps the buffers valid until b200_prover_wait. */
const char* b200_recursion_dev_async(b200_prover* p, uint32_t slot, const b200_circuit* c, const uint32_t* d_seal_a,
                                     size_t words_a, const uint32_t* d_seal_b, size_t words_b, uint32_t* h_seal);
/* The agent's task bodies as SINGLE enqueues, no host round trip between their steps:
 * b200_prove_lift_async = tasks::prove::prover (prover/crates/workflow/src/tasks/prove.rs:44-108): prove_segment -> verify_integrity ->
 *   lift -> verify_integrity.  Optional outputs (NULL = skip): segment / lifted seal in pinned host memory, lifted seal in caller-owned
 *   device memory, verdicts[2] (0 = valid; NULL skips both verifications).
 * b200_recursion_verified_async = tasks::join::join (tasks/join.rs:41-79; union / resolve alike): verify_integrity of the left and right
 *   receipts (DEVICE memory) against the circuits the caller expects, the recursion proof, verify_integrity of the result; verdicts[3].
 * All outputs are valid after b200_prover_wait(p, slot). */
const char* b200_prove_lift_async(b200_prover* p, uint32_t slot, const b200_circuit* seg, uint64_t seed, const uint32_t* h_trace,
                                  const b200_circuit* lift, uint32_t* h_seg_seal, uint32_t* h_lift_seal, uint32_t* d_lift_seal,
                                  int* h_verdicts);
const char* b200_recursion_verified_async(b200_prover* p, uint32_t slot, const b200_circuit* c, const uint32_t* d_seal_a,
                                          const b200_circuit* circuit_a, const uint32_t* d_seal_b, const b200_circuit* circuit_b,
                                          uint32_t* h_seal, uint32_t* d_seal_out, int* h_verdicts);
/* D2D copy of the seal the slot produced last (ordered behind the proof on the slot's stream) into caller-owned device memory */
const char* b200_seal_to_device(b200_prover* p, uint32_t slot, uint32_t* d_dst, size_t words);
/* 1 = everything enqueued on the slot has completed, 0 = still running, -1 = error.  Never blocks. */
int b200_prover_query(b200_prover* p, uint32_t slot);
/* verify_integrity (the check the reference runs after every prove / lift / join: tasks/prove.rs:56-58, tasks/join.rs:77-79),
 * on the device: transcript replay + the 50 queries in parallel.  h_seal == NULL verifies the seal the slot produced last
 * (still resident; may be enqueued right behind b200_prove_segment_async on a busy slot), otherwise `words` words from host
 * memory (slot must be idle).  *h_result is valid after b200_prover_wait: 0 = valid, else the code of the first failed check
 * (100-102 malformed header/length, 106 non-canonical word, 110 constraint identity, 120+g Merkle path of group g,
 * 130+k / 140+k FRI round k path / value, 150 final polynomial). */
const char* b200_verify_async(b200_prover* p, uint32_t slot, const uint32_t* h_seal, size_t words, int* h_result);
/* verify_integrity against the circuit the CALLER expects: the seal's header must equal `expect` (po2, widths, kind), else code 103 --
 * b200_verify_async above trusts the header, so a small seal of another kind would pass as whatever it says it is.  `seal` NULL = the
 * slot's own last seal; else `words` words in host memory (seal_on_device == 0) or caller-owned device memory (!= 0). */
const char* b200_verify_circuit_async(b200_prover* p, uint32_t slot, const b200_circuit* expect, const uint32_t* seal, size_t words,
                                      int seal_on_device, int* h_result);
const char* b200_prover_wait(b200_prover* p, uint32_t slot);
/* device time (ms) between the first and last operation of the slot's last proof */
float b200_prover_last_ms(b200_prover* p, uint32_t slot);
/* timing marks: record event `which` (0/1) on the slot's stream; elapsed ms between two marks (-1 if unset) */
const char* b200_prover_mark(b200_prover* p, uint32_t slot, uint32_t which);
float b200_prover_marks_ms(b200_prover* p, uint32_t slot_a, uint32_t which_a, uint32_t slot_b, uint32_t which_b);
/* witgen stand-in: expand `seed` into the (w_code + w_data) x 2^po2 witness and copy it to host memory (synchronous) */
const char* b200_witgen_to_host(b200_prover* p, uint32_t slot, const b200_circuit* c, uint64_t seed, uint32_t* h_trace);
/* number of kernels this library launched since load (all threads) */
uint64_t b200_kernel_launches(void);
/* host-pinned allocation helpers for h_trace / h_seal */
const char* b200_host_alloc(void** out, size_t bytes);
void b200_host_free(void* p);

/* ---- join-tree planner (mirrors taskdb::planner::Planner, prover/crates/taskdb/src/planner/mod.rs:91-240) -------- */
typedef struct b200_planner b200_planner;
enum { B200_CMD_KECCAK = 0, B200_CMD_FINALIZE = 1, B200_CMD_JOIN = 2, B200_CMD_SEGMENT = 3, B200_CMD_UNION = 4 };
/* mirrors planner::task::Task (prover/crates/taskdb/src/planner/task.rs:17-24) */
typedef struct {
    uint32_t task_number, task_height, command;
    uint32_t n_depends_on, depends_on[2];
    uint32_t n_keccak_depends_on, keccak_depends_on[2];
} b200_task;
b200_planner* b200_planner_new(void);
void b200_planner_free(b200_planner* pl);
/* return the new task number, or -1 "PlanFinalized" */
int64_t b200_planner_enqueue_segment(b200_planner* pl);
int64_t b200_planner_enqueue_keccak(b200_planner* pl);
/* joins the remaining peaks right-to-left, appends Finalize; returns its task number, or -1 "PlanNotStarted" */
int64_t b200_planner_finish(b200_planner* pl);
size_t b200_planner_task_count(const b200_planner* pl);
/* 0 on success, -1 "Invalid task number" */
int b200_planner_get_task(const b200_planner* pl, size_t task_number, b200_task* out);
/* iterator over tasks in creation order (Planner::next_task); 0 = produced a task, 1 = none pending */
int b200_planner_next_task(b200_planner* pl, b200_task* out);

#ifdef __cplusplus
}
#endif
#endif
