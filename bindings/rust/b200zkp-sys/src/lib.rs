//! Raw bindings to `include/b200zkp.h` (1:1, same names).  Source-only in this repository: never compiled here.
//!
//! Every function returns `*const c_char`: NULL on success, otherwise a thread-local message -- the convention of
//! risc0-sys' C wrappers, so a `Hal` shim can keep `fn ffi_wrap(...) -> anyhow::Result<()>`.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200_circuit { pub po2: u32, pub w_code: u32, pub w_data: u32, pub w_accum: u32, pub kind: u32 }

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200_task {
    pub task_number: u32, pub task_height: u32, pub command: u32,
    pub n_depends_on: u32, pub depends_on: [u32; 2],
    pub n_keccak_depends_on: u32, pub keccak_depends_on: [u32; 2],
}

pub enum b200_prover {}
pub enum b200_planner {}

extern "C" {
    pub fn b200_init(device: c_int) -> *const c_char;
    pub fn b200_last_error() -> *const c_char;
    pub fn b200_device_count() -> c_int;
    pub fn b200_device_pci_bus_id(device: c_int, out: *mut c_char, len: c_int) -> *const c_char;

    // kernel 1: NTT (sppark_batch_iNTT / _NTT / _expand / _zk_shift)
    pub fn b200_batch_intt(d_io: *mut u32, lg_n: u32, count: u32, stream: *mut c_void) -> *const c_char;
    pub fn b200_batch_ntt(d_io: *mut u32, lg_n: u32, count: u32, stream: *mut c_void) -> *const c_char;
    pub fn b200_batch_expand_ntt(d_out: *mut u32, d_in: *const u32, lg_n: u32, lg_blowup: u32, count: u32, stream: *mut c_void) -> *const c_char;
    pub fn b200_batch_zk_shift(d_io: *mut u32, lg_n: u32, count: u32, stream: *mut c_void) -> *const c_char;
    pub fn b200_batch_bit_reverse(d_io: *mut u32, lg_n: u32, count: u32, stream: *mut c_void) -> *const c_char;
    pub fn b200_batch_intt_zk_shift(d_io: *mut u32, lg_n: u32, count: u32, stream: *mut c_void) -> *const c_char;

    // kernel 2: Poseidon2 (sppark_poseidon2_rows / _fold)
    pub fn b200_poseidon2_rows(d_out: *mut u32, d_matrix: *const u32, rows: u32, cols: u32, stream: *mut c_void) -> *const c_char;
    pub fn b200_poseidon2_fold(d_out: *mut u32, d_in: *const u32, num_hashes: u32, stream: *mut c_void) -> *const c_char;
    pub fn b200_merkle_tree(d_nodes: *mut u32, d_matrix: *const u32, lg_rows: u32, cols: u32, stream: *mut c_void) -> *const c_char;

    // kernel 3: FRI fold + evaluation (fri_fold, batch_evaluate_any)
    pub fn b200_fri_fold(d_out: *mut u32, d_in: *const u32, in_size: u32, d_mix: *const u32, stream: *mut c_void) -> *const c_char;
    pub fn b200_evaluate_scratch_words(lg_n: u32, count: u32) -> usize;
    pub fn b200_batch_evaluate_any(d_out: *mut u32, d_coeffs: *const u32, lg_n: u32, count: u32, d_x: *const u32,
                                   d_scratch: *mut u32, stream: *mut c_void) -> *const c_char;

    // small HAL operations (mix_poly_coeffs, eltwise_*, supra_poly_divide, prefix_products, gather_sample, scatter, Merkle opening)
    pub fn b200_shutdown() -> *const c_char;
    pub fn b200_mix_poly_coeffs(d_out: *mut u32, d_mix_start: *const u32, d_mix: *const u32, d_in: *const u32, d_combos: *const u32,
                                input_size: u32, count: u32, n_combos: u32, stream: *mut c_void) -> *const c_char;
    pub fn b200_eltwise_sum_extelem(d_out: *mut u32, d_in: *const u32, count: u32, to_add: u32, stream: *mut c_void) -> *const c_char;
    pub fn b200_poly_divide_scratch_words(size: u32) -> usize;
    pub fn b200_poly_divide(d_poly: *mut u32, size: u32, d_remainder: *mut u32, d_pow: *const u32, d_scratch: *mut u32,
                            stream: *mut c_void) -> *const c_char;
    pub fn b200_prefix_products_scratch_words(count: u32) -> usize;
    pub fn b200_prefix_products(d_io: *mut u32, count: u32, d_scratch: *mut u32, stream: *mut c_void) -> *const c_char;
    pub fn b200_eltwise_add_elem(d_out: *mut u32, d_a: *const u32, d_b: *const u32, count: usize, stream: *mut c_void) -> *const c_char;
    pub fn b200_eltwise_copy_elem(d_out: *mut u32, d_in: *const u32, count: usize, stream: *mut c_void) -> *const c_char;
    pub fn b200_eltwise_zeroize_elem(d_io: *mut u32, count: usize, stream: *mut c_void) -> *const c_char;
    pub fn b200_gather_sample(d_dst: *mut u32, d_src: *const u32, idx: usize, size: u32, stride: usize, stream: *mut c_void) -> *const c_char;
    pub fn b200_scatter(d_into: *mut u32, d_index: *const u32, n_index: u32, d_offsets: *const u32, d_values: *const u32,
                        n_values: u32, stream: *mut c_void) -> *const c_char;
    pub fn b200_merkle_open_words(lg_rows: u32, cols: u32, top_size: u32) -> usize;
    pub fn b200_merkle_open(d_out: *mut u32, d_nodes: *const u32, d_matrix: *const u32, lg_rows: u32, cols: u32, top_size: u32,
                            idx: u32, stream: *mut c_void) -> *const c_char;
    pub fn b200_commit_group(d_coeffs_io: *mut u32, d_evals: *mut u32, d_nodes: *mut u32, lg_n: u32, count: u32,
                             stream: *mut c_void) -> *const c_char;

    // operator level (ProverServer::prove_segment / lift / join / resolve / union)
    pub fn b200_seal_words(c: *const b200_circuit) -> usize;
    pub fn b200_prover_create(out: *mut *mut b200_prover, device: c_int, max_circuit: *const b200_circuit, slots: u32) -> *const c_char;
    pub fn b200_prover_destroy(p: *mut b200_prover);
    pub fn b200_prover_device_bytes(p: *const b200_prover) -> usize;
    pub fn b200_prove_segment_async(p: *mut b200_prover, slot: u32, c: *const b200_circuit, seed: u64, h_trace: *const u32,
                                    h_seal: *mut u32) -> *const c_char;
    pub fn b200_prefetch_trace_async(p: *mut b200_prover, slot: u32, c: *const b200_circuit, h_trace: *const u32) -> *const c_char;
    pub fn b200_recursion_async(p: *mut b200_prover, slot: u32, c: *const b200_circuit, h_seal_a: *const u32, words_a: usize,
                                h_seal_b: *const u32, words_b: usize, h_seal: *mut u32) -> *const c_char;
    pub fn b200_recursion_dev_async(p: *mut b200_prover, slot: u32, c: *const b200_circuit, d_seal_a: *const u32, words_a: usize,
                                    d_seal_b: *const u32, words_b: usize, h_seal: *mut u32) -> *const c_char;
    // the agent's task bodies as single enqueues: tasks::prove::prover (prove.rs:44-108) and tasks::join::join (join.rs:41-79)
    pub fn b200_prove_lift_async(p: *mut b200_prover, slot: u32, seg: *const b200_circuit, seed: u64, h_trace: *const u32,
                                 lift: *const b200_circuit, h_seg_seal: *mut u32, h_lift_seal: *mut u32, d_lift_seal: *mut u32,
                                 h_verdicts: *mut c_int) -> *const c_char;
    pub fn b200_recursion_verified_async(p: *mut b200_prover, slot: u32, c: *const b200_circuit, d_seal_a: *const u32,
                                         circuit_a: *const b200_circuit, d_seal_b: *const u32, circuit_b: *const b200_circuit,
                                         h_seal: *mut u32, d_seal_out: *mut u32, h_verdicts: *mut c_int) -> *const c_char;
    pub fn b200_seal_to_device(p: *mut b200_prover, slot: u32, d_dst: *mut u32, words: usize) -> *const c_char;
    pub fn b200_prover_query(p: *mut b200_prover, slot: u32) -> c_int;
    pub fn b200_verify_async(p: *mut b200_prover, slot: u32, h_seal: *const u32, words: usize, h_result: *mut c_int) -> *const c_char;
    pub fn b200_verify_circuit_async(p: *mut b200_prover, slot: u32, expect: *const b200_circuit, seal: *const u32, words: usize,
                                     seal_on_device: c_int, h_result: *mut c_int) -> *const c_char;
    pub fn b200_prover_wait(p: *mut b200_prover, slot: u32) -> *const c_char;
    pub fn b200_prover_last_ms(p: *mut b200_prover, slot: u32) -> f32;
    pub fn b200_prover_mark(p: *mut b200_prover, slot: u32, which: u32) -> *const c_char;
    pub fn b200_prover_marks_ms(p: *mut b200_prover, slot_a: u32, which_a: u32, slot_b: u32, which_b: u32) -> f32;
    pub fn b200_witgen_to_host(p: *mut b200_prover, slot: u32, c: *const b200_circuit, seed: u64, h_trace: *mut u32) -> *const c_char;
    pub fn b200_kernel_launches() -> u64;
    pub fn b200_host_alloc(out: *mut *mut c_void, bytes: usize) -> *const c_char;
    pub fn b200_host_free(p: *mut c_void);

    // join-tree planner (taskdb::planner::Planner)
    pub fn b200_planner_new() -> *mut b200_planner;
    pub fn b200_planner_free(pl: *mut b200_planner);
    pub fn b200_planner_enqueue_segment(pl: *mut b200_planner) -> i64;
    pub fn b200_planner_enqueue_keccak(pl: *mut b200_planner) -> i64;
    pub fn b200_planner_finish(pl: *mut b200_planner) -> i64;
    pub fn b200_planner_task_count(pl: *const b200_planner) -> usize;
    pub fn b200_planner_get_task(pl: *const b200_planner, task_number: usize, out: *mut b200_task) -> c_int;
    pub fn b200_planner_next_task(pl: *mut b200_planner, out: *mut b200_task) -> c_int;
}

/// NULL -> Ok(()), message -> Err (what `ffi_wrap` does in risc0-sys).
pub fn check(err: *const c_char) -> Result<(), String> {
    if err.is_null() { Ok(()) } else { Err(unsafe { std::ffi::CStr::from_ptr(err) }.to_string_lossy().into_owned()) }
}
