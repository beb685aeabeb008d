// Host pipeline behind the ProverServer operator boundary: prove_segment / lift / join as ONE stream of kernel
// launches with no host round trip (transcript, challenges and query positions live in device memory).
//
// Mirrors the call sequence of risc0-zkp 3.0.3 prove/prover.rs + prove/fri.rs + prove/merkle.rs (un-vendored;
// SURVEY.md Appendix A "Prover loop") as invoked through ProverServer::prove_segment / lift / join at
// /root/reference/prover/crates/workflow/src/tasks/prove.rs:44-52, :96-104 and tasks/join.rs:52-56.
// The circuit is synthetic (DESIGN.md "Protocol"): witgen and eval_check are stand-ins, everything between them
// (K1-K9) is the real work at the real shapes.
#include "../../include/b200zkp.h"
#include "internal.h"
#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges cost nothing unless a tool (nsys, ncu --nvtx) is attached
#include <atomic>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

namespace b200 {

static unsigned ilog2(size_t x) { unsigned r = 0; while (((size_t)1 << r) < x) r++; return r; }

// NVTX range per proof phase (upstream risc0-core links nvtx for the same purpose: reference Cargo.lock:9012).  Host-side ranges around
// the ENQUEUE of each phase; profilers attribute the kernels launched inside to the range (ncu --nvtx --nvtx-include "b200/<phase>/").
struct Phase {
    explicit Phase(const char* name) { nvtxRangePushA(name); }
    ~Phase() { nvtxRangePop(); }
};

struct MerkleShape {
    uint32_t rows, cols, layers, top_layer, top_size;
    MerkleShape(uint32_t r, uint32_t c) : rows(r), cols(c) {
        layers = ilog2(r); top_layer = layers < 5 ? layers : 5; top_size = 1u << top_layer;
    }
    uint32_t proof_words() const { return cols + (layers - top_layer) * 8; }
};

static unsigned fri_rounds(unsigned po2, size_t* final_size) {
    size_t size = (size_t)1 << po2; unsigned r = 0;
    while (size > (size_t)FRI_MIN_DEGREE) { size /= FRI_FOLD; r++; }
    if (final_size) *final_size = size;
    return r;
}

static const char* check_circuit(const b200_circuit* c) {
    if (!c) return "b200: null circuit";
    if (c->po2 < 9 || c->po2 > 24) return "b200: po2 out of range [9,24]";
    if (c->w_code == 0 || c->w_code % 4 || c->w_data % 4 || c->w_accum % 4 || c->w_accum == 0)
        return "b200: column widths must be positive multiples of 4";
    if (c->w_accum > c->w_data) return "b200: w_accum must not exceed w_data";
    if (c->w_code + c->w_data + c->w_accum > 512) return "b200: too many columns (max 512)";
    return nullptr;
}

struct SealLayout {
    uint32_t off_top[4], off_u, off_fri_top[8], off_final, off_queries, query_words, total;
    uint32_t q_off_group[4], q_off_fri[8];
    unsigned rounds; size_t final_size;
    explicit SealLayout(const b200_circuit& c) {
        const uint32_t N = 1u << c.po2, D = 4 * N, W = c.w_code + c.w_data + c.w_accum;
        const uint32_t widths[4] = {c.w_code, c.w_data, c.w_accum, (uint32_t)CHECK_COLS};
        uint32_t pos = GLOBALS, q = 0;
        for (int g = 0; g < 4; g++) {
            MerkleShape m(D, widths[g]);
            off_top[g] = pos; pos += m.top_size * 8;
            q_off_group[g] = q; q += m.proof_words();
        }
        off_u = pos; pos += (W + c.w_accum + CHECK_COLS) * 4;
        rounds = fri_rounds(c.po2, &final_size);
        uint32_t size = N;
        for (unsigned r = 0; r < rounds; r++) {
            MerkleShape m(4 * size / FRI_FOLD, 4 * FRI_FOLD);
            off_fri_top[r] = pos; pos += m.top_size * 8;
            q_off_fri[r] = q; q += m.proof_words();
            size /= FRI_FOLD;
        }
        off_final = pos; pos += 4 * (uint32_t)final_size;
        off_queries = pos; query_words = q;
        total = pos + QUERIES * q;
    }
};

std::atomic<uint64_t> g_kernel_launches{0};

// bump allocator over one cudaMalloc'd arena
struct Arena {
    uint32_t* base = nullptr; size_t words = 0, used = 0;
    uint32_t* take(size_t n) { n = (n + 63) & ~(size_t)63; uint32_t* p = base + used; used += n; return p; }
};

// Everything one verify_integrity run touches.  A slot verifies its own seal in place with `vmain` (the slot's stream and buffers); the
// two auxiliary contexts have their own stream and buffers so that the verification of a task's INPUT receipts (tasks/join.rs:41-46)
// and of a just-proved segment (tasks/prove.rs:56-58) runs beside the proof that follows instead of in front of it.
struct VerifyCtx {
    cudaStream_t stream = nullptr; cudaEvent_t ev_done = nullptr;
    uint32_t *seal = nullptr, *chal = nullptr, *pts = nullptr, *pos = nullptr, *mp = nullptr, *pmix = nullptr, *vctx = nullptr;
    Transcript* tr = nullptr;
    uint32_t* alloc = nullptr;           // aux contexts: the one cudaMalloc behind the pointers above
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_mark[2] = {nullptr, nullptr};
    Arena arena;
    // fixed regions (sized for the max circuit)
    uint32_t *coeffs, *evals, *check_coeffs, *check_evals, *nodes[4];
    uint32_t *fri_evals, *fri_nodes, *fri_coeffs;
    uint32_t *combos, *f_planes, *chunk_vals, *chunk_carry, *ev_scratch;
    uint32_t *pmix, *mp, *chal, *pts, *pos, *seal, *digests, *vctx;
    Transcript* tr;
    GatherTree* d_trees;
    std::vector<GatherTree> h_trees;
    uint32_t* h_seal_out = nullptr; size_t seal_words = 0;
    uint32_t* h_stage = nullptr; size_t h_stage_words = 0;   // pinned staging for child seals
    bool busy = false;
    // witness prefetch (b200_prefetch_trace_async): a second coefficient region filled by a copy stream while the slot proves
    uint32_t* coeffs_alt = nullptr; size_t coeffs_words = 0;
    uint32_t* alt_alloc = nullptr;            // the cudaMalloc'ed region (coeffs / coeffs_alt swap roles, this is what gets freed)
    cudaStream_t copy_stream = nullptr; cudaEvent_t ev_staged = nullptr;
    const uint32_t* staged_src = nullptr; size_t staged_words = 0;
    b200_circuit last_circuit{}; bool has_seal = false;     // what s.seal holds (for verify of the slot's own proof)
    // verdicts travel device -> PINNED host words (a D2H copy into pageable memory would block the enqueueing thread until the stream
    // drains); b200_prover_wait hands them to the caller's ints
    int* h_vpin = nullptr; int* user_verdict[8] = {}; int n_verdicts = 0;
    VerifyCtx vmain, vaux[2];
    cudaEvent_t ev_fork = nullptr;
};

}  // namespace b200

using namespace b200;

struct b200_prover {
    int device = 0;
    b200_circuit maxc{};
    const DeviceTables* T = nullptr;
    std::vector<Slot> slots;
    size_t device_bytes = 0;
};

namespace b200 {

static size_t arena_words(const b200_circuit& c) {
    const size_t N = (size_t)1 << c.po2, D = 4 * N, W = c.w_code + c.w_data + c.w_accum;
    size_t w = 0;
    auto add = [&](size_t n) { w += (n + 63) & ~(size_t)63; };
    add(W * N); add(W * D); add(CHECK_COLS * N); add(CHECK_COLS * D);
    for (int g = 0; g < 4; g++) add(2 * D * 8);
    add(16 * N + 16 * N / 8); add(8 * N + 4096); add(4 * N);            // fri evals / nodes / coeffs (geometric sums, padded)
    add(12 * N); add(4 * N); add(3 * 4 * (N / 2048 + 1)); add(3 * 4 * (N / 2048 + 1));
    add(evaluate_scratch_words(c.po2, (uint32_t)W));
    add(4 * (W / 4 + c.w_accum)); add(4 * (W + c.w_accum + CHECK_COLS)); add(64); add(16); add(64);
    add(SealLayout(c).total); add(64);
    add(sizeof(Transcript) / 4); add(16 * sizeof(GatherTree) / 4); add(VCTX_WORDS);
    return w + 1024;
}

#define CU(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { set_error("b200: %s failed: %s", #x, cudaGetErrorString(e__)); return last_error(); } } while (0)

static const char* slot_init(b200_prover* p, Slot& s) {
    CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&s.ev_begin)); CU(cudaEventCreate(&s.ev_end));
    const b200_circuit& c = p->maxc;
    const size_t N = (size_t)1 << c.po2, D = 4 * N, W = c.w_code + c.w_data + c.w_accum;
    s.arena.words = arena_words(c);
    CU(cudaMalloc(&s.arena.base, s.arena.words * 4));
    p->device_bytes += s.arena.words * 4;
    Arena& a = s.arena;
    s.coeffs = a.take(W * N); s.coeffs_words = W * N; s.evals = a.take(W * D); s.check_coeffs = a.take(CHECK_COLS * N); s.check_evals = a.take(CHECK_COLS * D);
    for (int g = 0; g < 4; g++) s.nodes[g] = a.take(2 * D * 8);
    s.fri_evals = a.take(16 * N + 16 * N / 8); s.fri_nodes = a.take(8 * N + 4096); s.fri_coeffs = a.take(4 * N);
    s.combos = a.take(12 * N); s.f_planes = a.take(4 * N);
    s.chunk_vals = a.take(3 * 4 * (N / 2048 + 1)); s.chunk_carry = a.take(3 * 4 * (N / 2048 + 1));
    s.ev_scratch = a.take(evaluate_scratch_words(c.po2, (uint32_t)W));
    s.pmix = a.take(4 * (W / 4 + c.w_accum)); s.mp = a.take(4 * (W + c.w_accum + CHECK_COLS));
    s.chal = a.take(64); s.pts = a.take(16); s.pos = a.take(64);
    s.seal = a.take(SealLayout(c).total); s.digests = a.take(64);
    s.tr = reinterpret_cast<Transcript*>(a.take(sizeof(Transcript) / 4));
    s.d_trees = reinterpret_cast<GatherTree*>(a.take(16 * sizeof(GatherTree) / 4));
    s.vctx = a.take(VCTX_WORDS);
    if (a.used > a.words) { set_error("b200: arena overflow (%zu > %zu words)", a.used, a.words); return last_error(); }
    s.vmain.stream = s.stream; s.vmain.seal = s.seal; s.vmain.chal = s.chal; s.vmain.pts = s.pts; s.vmain.pos = s.pos; s.vmain.mp = s.mp;
    s.vmain.pmix = s.pmix; s.vmain.vctx = s.vctx; s.vmain.tr = s.tr;
    CU(cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
    const size_t w_seal = (SealLayout(c).total + 63) & ~(size_t)63, w_mp = (4 * (W + c.w_accum + CHECK_COLS) + 63) & ~(size_t)63,
                 w_pmix = (4 * (W / 4 + c.w_accum) + 63) & ~(size_t)63;
    for (auto& v : s.vaux) {
        CU(cudaStreamCreateWithFlags(&v.stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&v.ev_done, cudaEventDisableTiming));
        const size_t words = w_seal + w_mp + w_pmix + 64 + 64 + 64 + VCTX_WORDS + sizeof(Transcript) / 4 + 64;
        CU(cudaMalloc(&v.alloc, words * 4));
        p->device_bytes += words * 4;
        uint32_t* q = v.alloc;
        v.seal = q; q += w_seal; v.mp = q; q += w_mp; v.pmix = q; q += w_pmix; v.chal = q; q += 64; v.pts = q; q += 64; v.pos = q; q += 64;
        v.vctx = q; q += VCTX_WORDS; v.tr = reinterpret_cast<Transcript*>(q);
    }
    return nullptr;
}

#define KL(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { set_error("b200: %s: %s", #x, cudaGetErrorString(e__)); return last_error(); } } while (0)

// iNTT + zk_shift (optional) + expand/NTT + leaf hash + tree + top layer to the seal + rng.mix(root)
static const char* commit_group(b200_prover* p, Slot& s, uint32_t* coeffs, uint32_t* evals, uint32_t* nodes, uint32_t po2,
                                uint32_t cols, bool interp_shift, uint32_t seal_off) {
    cudaStream_t st = s.stream;
    const uint32_t D = 4u << po2;
    if (interp_shift) {
        KL(launch_batch_intt_shift(p->T, coeffs, po2, cols, st));      // K1 + K2 (coset shift fused into the last iNTT pass)
    }
    KL(launch_batch_expand_ntt(p->T, evals, coeffs, po2, INV_RATE_LG, cols, st));   // K3
    KL(launch_poseidon2_rows(nodes + (size_t)D * 8, evals, D, cols, D, st));         // K4
    KL(launch_poseidon2_fold_tree(nodes, po2 + 2, st));                               // K5
    MerkleShape m(D, cols);
    CU(cudaMemcpyAsync(s.seal + seal_off, nodes + (size_t)m.top_size * 8, (size_t)m.top_size * 32, cudaMemcpyDeviceToDevice, st));
    KL(launch_iop_commit(s.tr, nodes + 8, st));
    return nullptr;
}

// The proof proper.  Precondition: s.seal[8..16) (input digest) is already being produced on s.stream, and for
// recursion kinds s.digests[0..2) holds the trace seed words.
static const char* prove_on_slot(b200_prover* p, Slot& s, const b200_circuit& c, uint64_t seed, bool seed_on_device,
                                 const uint32_t* h_trace, uint32_t* h_seal, bool witness_staged = false) {
    cudaStream_t st = s.stream;
    const uint32_t po2 = c.po2, N = 1u << po2, D = 4 * N;
    const uint32_t W = c.w_code + c.w_data + c.w_accum, T = W + c.w_accum + CHECK_COLS;
    const SealLayout L(c);
    uint32_t* code = s.coeffs; uint32_t* data = code + (size_t)c.w_code * N; uint32_t* accum = data + (size_t)c.w_data * N;
    uint32_t* ecode = s.evals; uint32_t* edata = ecode + (size_t)c.w_code * D; uint32_t* eaccum = edata + (size_t)c.w_data * D;

    Phase proof_range(c.kind == 0 ? "b200/prove_segment" : "b200/recursion");
    KL(launch_iop_init(s.tr, st));
    KL(launch_iop_commit_elems(s.tr, s.seal, GLOBALS, nullptr, st));

    // witness: host trace (H2D) or the witgen stand-in
    const size_t tw = (size_t)(c.w_code + c.w_data) * N;
    if (witness_staged) { /* already in `code` (prefetched; the stream waits on ev_staged) */ }
    else if (h_trace) CU(cudaMemcpyAsync(code, h_trace, tw * 4, cudaMemcpyHostToDevice, st));
    else KL(launch_gen_trace(code, seed, seed_on_device ? s.digests : nullptr, tw, st));
    CU(cudaMemcpyAsync(accum, data, (size_t)c.w_accum * N * 4, cudaMemcpyDeviceToDevice, st));   // raw data columns for accumulate

    const char* e;
    { Phase ph("commit_group(code)"); if ((e = commit_group(p, s, code, ecode, s.nodes[0], po2, c.w_code, true, L.off_top[0]))) return e; }
    { Phase ph("commit_group(data)"); if ((e = commit_group(p, s, data, edata, s.nodes[1], po2, c.w_data, true, L.off_top[1]))) return e; }

    uint32_t* accum_mix = s.chal; uint32_t* poly_mix = s.chal + 4; uint32_t* z = s.chal + 8; uint32_t* mix = s.chal + 12;
    uint32_t* fri_mix = s.chal + 16;
    {
        Phase ph("accumulate + commit_group(accum)");
        KL(launch_iop_draw_ext(s.tr, accum_mix, 1, st));
        KL(launch_accumulate(accum, N, c.w_accum, accum_mix, st));
        if ((e = commit_group(p, s, accum, eaccum, s.nodes[2], po2, c.w_accum, true, L.off_top[2]))) return e;
    }
    {   // constraint stand-in over the 4N domain -> 4 planes x 4N -> iNTT -> 16 columns x N
        Phase ph("eval_check + commit_group(check)");
        KL(launch_iop_draw_ext(s.tr, poly_mix, 1, st));
        KL(launch_powers(s.pmix, poly_mix, W / 4 + c.w_accum, st));
        KL(launch_eval_check(s.check_coeffs, s.evals, po2 + 2, c.w_code, c.w_data, c.w_accum, s.pmix, st));
        KL(launch_batch_intt(p->T, s.check_coeffs, po2 + 2, 4, st));
        if ((e = commit_group(p, s, s.check_coeffs, s.check_evals, s.nodes[3], po2, CHECK_COLS, false, L.off_top[3]))) return e;
    }

    // DEEP point and tap evaluations (K7), written straight into the seal
    nvtxRangePushA("taps (K7) + DEEP (K8)");
    KL(launch_iop_draw_ext(s.tr, z, 1, st));
    KL(launch_deep_points(s.pts, z, p->T->rou_rev[po2], st));
    uint32_t* u = s.seal + L.off_u;
    KL(launch_evaluate(u, u + 4 * (size_t)W, s.coeffs, po2, W, s.pts, s.pts + 4, W - c.w_accum, W, s.ev_scratch, st));
    KL(launch_evaluate(u + 4 * (size_t)(W + c.w_accum), nullptr, s.check_coeffs, po2, CHECK_COLS, s.pts + 8, nullptr, 0, 0, s.ev_scratch, st));
    KL(launch_iop_commit_elems(s.tr, u, T * 4, nullptr, st));

    // DEEP combination (K8)
    KL(launch_iop_draw_ext(s.tr, mix, 1, st));
    KL(launch_powers(s.mp, mix, T, st));
    DeepArgs da{s.coeffs, s.check_coeffs, u, s.mp, s.pts, s.combos, s.chunk_vals, s.chunk_carry, s.f_planes, po2, W, c.w_accum};
    KL(launch_deep(da, st));
    nvtxRangePop();

    // FRI commit phase
    nvtxRangePushA("FRI commit (K3, K4, K5, K6)");
    s.h_trees.clear();
    const uint32_t widths[4] = {c.w_code, c.w_data, c.w_accum, (uint32_t)CHECK_COLS};
    const uint32_t* gev[4] = {ecode, edata, eaccum, s.check_evals};
    for (int g = 0; g < 4; g++) {
        MerkleShape m(D, widths[g]);
        s.h_trees.push_back(GatherTree{gev[g], s.nodes[g], D, widths[g], m.top_size, L.q_off_group[g], 0});
    }
    uint32_t* cur = s.f_planes; uint32_t size = N, lg = po2;
    uint32_t* fev = s.fri_evals; uint32_t* fnodes = s.fri_nodes; uint32_t* fco = s.fri_coeffs;
    for (unsigned r = 0; r < L.rounds; r++) {
        const uint32_t dom = 4 * size, rows = dom / FRI_FOLD;
        KL(launch_batch_expand_ntt(p->T, fev, cur, lg, INV_RATE_LG, 4, st));
        KL(launch_poseidon2_rows(fnodes + (size_t)rows * 8, fev, rows, 4 * FRI_FOLD, rows, st));
        KL(launch_poseidon2_fold_tree(fnodes, ilog2(rows), st));
        MerkleShape m(rows, 4 * FRI_FOLD);
        CU(cudaMemcpyAsync(s.seal + L.off_fri_top[r], fnodes + (size_t)m.top_size * 8, (size_t)m.top_size * 32, cudaMemcpyDeviceToDevice, st));
        KL(launch_iop_commit(s.tr, fnodes + 8, st));
        KL(launch_iop_draw_ext(s.tr, fri_mix, 1, st));
        KL(launch_fri_fold(fco, cur, size, fri_mix, st));
        s.h_trees.push_back(GatherTree{fev, fnodes, rows, 4 * FRI_FOLD, m.top_size, L.q_off_fri[r], rows});
        cur = fco; fco += 4 * (size_t)(size / FRI_FOLD);
        fev += 4 * (size_t)dom; fnodes += 2 * (size_t)rows * 8;
        size /= FRI_FOLD; lg -= 4;
    }
    CU(cudaMemcpyAsync(s.seal + L.off_final, cur, (size_t)4 * size * 4, cudaMemcpyDeviceToDevice, st));
    KL(launch_iop_commit_elems(s.tr, s.seal + L.off_final, 4 * size, nullptr, st));
    nvtxRangePop();

    // query phase (K9)
    Phase qph("queries (K9)");
    KL(launch_iop_draw_bits(s.tr, s.pos, QUERIES, po2 + 2, st));
    CU(cudaMemcpyAsync(s.d_trees, s.h_trees.data(), s.h_trees.size() * sizeof(GatherTree), cudaMemcpyHostToDevice, st));
    KL(launch_gather_queries(s.seal, L.off_queries, L.query_words, s.pos, s.d_trees, (uint32_t)s.h_trees.size(), st));

    if (h_seal) CU(cudaMemcpyAsync(h_seal, s.seal, (size_t)L.total * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(s.ev_end, st));
    s.busy = true; s.h_seal_out = h_seal; s.seal_words = L.total;
    s.last_circuit = c; s.has_seal = true;
    return nullptr;
}

// verify_integrity of the seal held in s.seal (circuit c): replay the transcript, then check the 50 queries in parallel.
// Uses the slot's challenge / transcript scratch, so it is ordered after any proof on the same slot by the stream.
static const char* verify_run(b200_prover* p, Slot& s, VerifyCtx& v, const b200_circuit& c, int* h_result, bool check_header) {
    cudaStream_t st = v.stream;
    const uint32_t po2 = c.po2, N = 1u << po2, D = 4 * N;
    const uint32_t W = c.w_code + c.w_data + c.w_accum, T = W + c.w_accum + CHECK_COLS;
    const SealLayout L(c);
    Phase vrange("b200/verify_integrity");
    VerifyShape sh{};
    sh.po2 = po2; sh.w_code = c.w_code; sh.w_data = c.w_data; sh.w_accum = c.w_accum; sh.W = W; sh.T = T;
    sh.rounds = L.rounds; sh.final_size = (uint32_t)L.final_size; sh.final_lg = ilog2(L.final_size);
    for (int g = 0; g < 4; g++) { sh.off_top[g] = L.off_top[g]; sh.q_off_group[g] = L.q_off_group[g]; }
    sh.off_u = L.off_u; sh.off_final = L.off_final; sh.off_queries = L.off_queries; sh.query_words = L.query_words;
    memcpy(sh.rou_fwd, p->T->rou_fwd, sizeof sh.rou_fwd);
    sh.inv16 = h_inv(h_to_mont(16));
    uint32_t* seal = v.seal;
    uint32_t* accum_mix = v.chal; uint32_t* poly_mix = v.chal + 4; uint32_t* z = v.chal + 8; uint32_t* mix = v.chal + 12;
    uint32_t* fmix = v.chal + 16;
    uint32_t* root = v.vctx + VCTX_ROOT;
    const uint32_t widths[4] = {c.w_code, c.w_data, c.w_accum, (uint32_t)CHECK_COLS};

    KL(launch_verify_reset(v.vctx, st));
    if (check_header) KL(launch_verify_header(v.vctx, seal, c.po2, c.w_code, c.w_data, c.w_accum, c.kind, st));
    KL(launch_verify_canonical(v.vctx, seal, L.total, st));
    KL(launch_iop_init(v.tr, st));
    KL(launch_iop_commit_elems(v.tr, seal, GLOBALS, nullptr, st));
    auto commit_top = [&](uint32_t off, uint32_t rows, uint32_t cols) -> const char* {
        MerkleShape m(rows, cols);
        KL(launch_verify_fold_top(root, seal + off, m.top_size, st));
        KL(launch_iop_commit(v.tr, root, st));
        return nullptr;
    };
    const char* e;
    if ((e = commit_top(L.off_top[0], D, widths[0]))) return e;
    if ((e = commit_top(L.off_top[1], D, widths[1]))) return e;
    KL(launch_iop_draw_ext(v.tr, accum_mix, 1, st));          // binds the transcript; accum itself is witness
    if ((e = commit_top(L.off_top[2], D, widths[2]))) return e;
    KL(launch_iop_draw_ext(v.tr, poly_mix, 1, st));
    if ((e = commit_top(L.off_top[3], D, widths[3]))) return e;
    KL(launch_iop_draw_ext(v.tr, z, 1, st));
    KL(launch_deep_points(v.pts, z, p->T->rou_rev[po2], st));
    const uint32_t* u = seal + L.off_u;
    KL(launch_iop_commit_elems(v.tr, u, T * 4, nullptr, st));
    KL(launch_powers(v.pmix, poly_mix, W / 4 + c.w_accum, st));
    KL(launch_verify_constraint(v.vctx, u, v.pmix, z, c.w_code, c.w_data, c.w_accum, st));
    KL(launch_iop_draw_ext(v.tr, mix, 1, st));
    KL(launch_powers(v.mp, mix, T, st));
    KL(launch_verify_usum(v.vctx, u, v.mp, W, c.w_accum, T, st));
    uint32_t size = N;
    for (unsigned r = 0; r < L.rounds; r++) {
        const uint32_t rows = 4 * size / FRI_FOLD;
        MerkleShape m(rows, 4 * FRI_FOLD);
        sh.off_fri_top[r] = L.off_fri_top[r]; sh.q_off_fri[r] = L.q_off_fri[r]; sh.fri_rows[r] = rows; sh.fri_top[r] = m.top_size;
        if ((e = commit_top(L.off_fri_top[r], rows, 4 * FRI_FOLD))) return e;
        KL(launch_iop_draw_ext(v.tr, fmix + 4 * r, 1, st));
        size /= FRI_FOLD;
    }
    KL(launch_iop_commit_elems(v.tr, seal + L.off_final, 4 * size, nullptr, st));
    KL(launch_iop_draw_bits(v.tr, v.pos, QUERIES, po2 + 2, st));
    KL(launch_verify_queries(v.vctx, seal, sh, v.mp, v.pts, fmix, v.pos, st));
    KL(launch_verify_finish(v.vctx, st));
    if (s.n_verdicts >= 8) { set_error("b200: too many verifications pending on one slot (call b200_prover_wait)"); return last_error(); }
    CU(cudaMemcpyAsync(&s.h_vpin[s.n_verdicts], v.vctx + VCTX_RESULT, 4, cudaMemcpyDeviceToHost, st));
    s.user_verdict[s.n_verdicts++] = h_result;
    s.busy = true;
    return nullptr;
}
static const char* verify_on_slot(b200_prover* p, Slot& s, const b200_circuit& c, int* h_result, bool check_header = false) {
    return verify_run(p, s, s.vmain, c, h_result, check_header);
}
// verify `words` words of a seal in device memory on auxiliary context k, concurrently with whatever follows on the slot's stream;
// `after_slot`: the seal is produced on the slot's stream (the copy is ordered behind it).  join_aux() makes the slot's stream wait.
static const char* verify_aux(b200_prover* p, Slot& s, int k, const b200_circuit& c, const uint32_t* d_seal, size_t words, bool after_slot,
                              int* h_result) {
    VerifyCtx& v = s.vaux[k];
    if (after_slot) {      // the copy rides on the slot's stream: behind the proof that writes the seal, ahead of whatever overwrites it next
        CU(cudaMemcpyAsync(v.seal, d_seal, words * 4, cudaMemcpyDeviceToDevice, s.stream));
        CU(cudaEventRecord(s.ev_fork, s.stream));
        CU(cudaStreamWaitEvent(v.stream, s.ev_fork, 0));
    } else {
        CU(cudaMemcpyAsync(v.seal, d_seal, words * 4, cudaMemcpyDeviceToDevice, v.stream));
    }
    const char* e = verify_run(p, s, v, c, h_result, true);
    if (e) return e;
    CU(cudaEventRecord(v.ev_done, v.stream));
    return nullptr;
}
static const char* join_aux(Slot& s, int k) {
    CU(cudaStreamWaitEvent(s.stream, s.vaux[k].ev_done, 0));
    return nullptr;
}


// digest of a seal -> d_out8 (device): 1024-row column-major view, hash_rows, fold (K4/K5).  The seal is read from host memory
// (staged through the slot's pinned buffer) or, with on_device, straight from caller-owned device memory (no host bounce).
static const char* seal_digest_async(b200_prover* p, Slot& s, const uint32_t* seal, size_t words, bool on_device, uint32_t* h_stage,
                                     uint32_t* d_scratch_matrix, uint32_t* d_scratch_nodes, uint32_t* d_out8) {
    (void)p;
    const uint32_t rows = 1024; const size_t cols = (words + rows - 1) / rows;
    CU(cudaMemsetAsync(d_scratch_matrix, 0, cols * rows * 4, s.stream));
    if (on_device) {
        CU(cudaMemcpyAsync(d_scratch_matrix, seal, words * 4, cudaMemcpyDeviceToDevice, s.stream));
    } else {
        memcpy(h_stage, seal, words * 4);
        CU(cudaMemcpyAsync(d_scratch_matrix, h_stage, words * 4, cudaMemcpyHostToDevice, s.stream));
    }
    KL(launch_poseidon2_rows(d_scratch_nodes + (size_t)rows * 8, d_scratch_matrix, rows, (uint32_t)cols, rows, s.stream));
    KL(launch_poseidon2_fold_tree(d_scratch_nodes, 10, s.stream));
    CU(cudaMemcpyAsync(d_out8, d_scratch_nodes + 8, 32, cudaMemcpyDeviceToDevice, s.stream));
    return nullptr;
}

}  // namespace b200

extern "C" {

uint64_t b200_kernel_launches(void) { return g_kernel_launches.load(); }

size_t b200_seal_words(const b200_circuit* c) {
    if (check_circuit(c)) return 0;
    return SealLayout(*c).total;
}

const char* b200_prover_create(b200_prover** out, int device, const b200_circuit* maxc, uint32_t slots) {
    if (!out) return "b200: null out";
    *out = nullptr;
    const char* ce = check_circuit(maxc);
    if (ce) { set_error("%s", ce); return last_error(); }
    if (slots == 0 || slots > 16) { set_error("b200: slots must be in [1,16]"); return last_error(); }
    const DeviceTables* T = get_tables(device);
    if (!T) return last_error();
    CU(cudaSetDevice(device));
    b200_prover* p = new (std::nothrow) b200_prover();
    if (!p) { set_error("b200: out of host memory"); return last_error(); }
    p->device = device; p->maxc = *maxc; p->T = T;
    p->slots.resize(slots);
    for (auto& s : p->slots) {
        const char* e = slot_init(p, s);
        if (e) { b200_prover_destroy(p); return e; }
        const size_t stage_words = 2 * (SealLayout(*maxc).total + 4096);
        if (cudaHostAlloc((void**)&s.h_stage, stage_words * 4, cudaHostAllocDefault) != cudaSuccess) {
            set_error("b200: cudaHostAlloc failed"); b200_prover_destroy(p); return last_error();
        }
        s.h_stage_words = stage_words;
        if (cudaHostAlloc((void**)&s.h_vpin, 8 * sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
            set_error("b200: cudaHostAlloc failed"); b200_prover_destroy(p); return last_error();
        }
    }
    *out = p;
    return nullptr;
}

void b200_prover_destroy(b200_prover* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (auto& s : p->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.arena.base) cudaFree(s.arena.base);
        if (s.h_stage) cudaFreeHost(s.h_stage);
        if (s.h_vpin) cudaFreeHost(s.h_vpin);
        for (auto& v : s.vaux) {
            if (v.stream) { cudaStreamSynchronize(v.stream); cudaStreamDestroy(v.stream); }
            if (v.ev_done) cudaEventDestroy(v.ev_done);
            if (v.alloc) cudaFree(v.alloc);
        }
        if (s.ev_fork) cudaEventDestroy(s.ev_fork);
        if (s.alt_alloc) cudaFree(s.alt_alloc);
        if (s.ev_staged) cudaEventDestroy(s.ev_staged);
        if (s.copy_stream) cudaStreamDestroy(s.copy_stream);
        if (s.ev_begin) cudaEventDestroy(s.ev_begin);
        if (s.ev_end) cudaEventDestroy(s.ev_end);
        for (int i = 0; i < 2; i++) if (s.ev_mark[i]) cudaEventDestroy(s.ev_mark[i]);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    delete p;
}

size_t b200_prover_device_bytes(const b200_prover* p) { return p ? p->device_bytes : 0; }

static const char* slot_check(b200_prover* p, uint32_t slot, const b200_circuit* c) {
    if (!p) { set_error("b200: null prover"); return last_error(); }
    if (slot >= p->slots.size()) { set_error("b200: slot %u out of range", slot); return last_error(); }
    const char* ce = check_circuit(c);
    if (ce) { set_error("%s", ce); return last_error(); }
    if (c->po2 > p->maxc.po2 || c->w_code + c->w_data + c->w_accum > p->maxc.w_code + p->maxc.w_data + p->maxc.w_accum ||
        c->w_accum > p->maxc.w_accum) { set_error("b200: circuit exceeds the prover's max_circuit"); return last_error(); }
    if (p->slots[slot].busy) { set_error("b200: slot %u busy (call b200_prover_wait first)", slot); return last_error(); }
    CU(cudaSetDevice(p->device));
    return nullptr;
}

const char* b200_prove_segment_async(b200_prover* p, uint32_t slot, const b200_circuit* c, uint64_t seed, const uint32_t* h_trace,
                                     uint32_t* h_seal) {
    const char* e = slot_check(p, slot, c);
    if (e) return e;
    if (!h_seal) { set_error("b200: null h_seal"); return last_error(); }
    Slot& s = p->slots[slot];
    CU(cudaEventRecord(s.ev_begin, s.stream));
    KL(launch_set_globals(s.seal, c->po2, c->w_code, c->w_data, c->w_accum, c->kind, seed, 1, s.stream));
    bool staged = false;
    if (h_trace && s.staged_src == h_trace && s.staged_words == ((size_t)(c->w_code + c->w_data) << c->po2)) {
        // the witness was prefetched into the other coefficient region: make that region the current one
        CU(cudaStreamWaitEvent(s.stream, s.ev_staged, 0));
        uint32_t* t = s.coeffs; s.coeffs = s.coeffs_alt; s.coeffs_alt = t;
        staged = true;
    }
    s.staged_src = nullptr; s.staged_words = 0;
    return prove_on_slot(p, s, *c, seed, false, h_trace, h_seal, staged);
}

// Copy the NEXT segment's witness host -> device on the slot's copy stream while the slot is still proving the current one; the
// following b200_prove_segment_async with the same h_trace pointer picks it up without a copy on its own stream.
const char* b200_prefetch_trace_async(b200_prover* p, uint32_t slot, const b200_circuit* c, const uint32_t* h_trace) {
    if (!p) { set_error("b200: null prover"); return last_error(); }
    if (slot >= p->slots.size()) { set_error("b200: slot %u out of range", slot); return last_error(); }
    const char* ce = check_circuit(c);
    if (ce) { set_error("%s", ce); return last_error(); }
    if (!h_trace) { set_error("b200: null h_trace"); return last_error(); }
    Slot& s = p->slots[slot];
    const size_t tw = (size_t)(c->w_code + c->w_data) << c->po2;
    if (c->po2 > p->maxc.po2 || ((size_t)(c->w_code + c->w_data + c->w_accum) << c->po2) > s.coeffs_words) {
        set_error("b200: circuit exceeds the prover's max_circuit"); return last_error();
    }
    CU(cudaSetDevice(p->device));
    if (!s.coeffs_alt) {      // first use on this slot: a synchronising allocation, keep it out of timed regions (warm up with one prefetch)
        CU(cudaMalloc(&s.alt_alloc, s.coeffs_words * 4));
        s.coeffs_alt = s.alt_alloc;
        p->device_bytes += s.coeffs_words * 4;
        CU(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s.ev_staged, cudaEventDisableTiming));
    }
    // coeffs_alt was the region of this slot's previous proof, which the caller has already waited for
    CU(cudaMemcpyAsync(s.coeffs_alt, h_trace, tw * 4, cudaMemcpyHostToDevice, s.copy_stream));
    CU(cudaEventRecord(s.ev_staged, s.copy_stream));
    s.staged_src = h_trace; s.staged_words = tw;
    return nullptr;
}

// recursion proof on a slot that has been checked already.  The children are digested BEFORE the new globals are written, so a child
// may be the slot's own previous seal (lift right behind prove_segment, no copy).
static const char* recursion_on_slot(b200_prover* p, Slot& s, const b200_circuit* c, const uint32_t* h_a, size_t wa,
                                     const uint32_t* h_b, size_t wb, bool on_device, uint32_t* h_seal) {
    const char* e;
    if (wa + wb > s.h_stage_words) { set_error("b200: child seal too large"); return last_error(); }
    // scratch: the evaluation / node regions are free until the proof starts
    if ((e = seal_digest_async(p, s, h_a, wa, on_device, s.h_stage, s.evals, s.nodes[0], s.digests + 8))) return e;
    if (h_b && wb && (e = seal_digest_async(p, s, h_b, wb, on_device, s.h_stage + wa, s.evals, s.nodes[0], s.digests + 16))) return e;
    KL(launch_set_globals(s.seal, c->po2, c->w_code, c->w_data, c->w_accum, c->kind, 0, 0, s.stream));
    if (h_b && wb) {
        KL(launch_hash_pair_one(s.seal + 8, s.digests + 8, s.digests + 16, s.stream));
    } else {
        CU(cudaMemcpyAsync(s.seal + 8, s.digests + 8, 32, cudaMemcpyDeviceToDevice, s.stream));
    }
    // trace seed = first two words of the input digest
    CU(cudaMemcpyAsync(s.digests, s.seal + 8, 8, cudaMemcpyDeviceToDevice, s.stream));
    return prove_on_slot(p, s, *c, 0, true, nullptr, h_seal);
}
static const char* recursion_common(b200_prover* p, uint32_t slot, const b200_circuit* c, const uint32_t* h_a, size_t wa,
                                    const uint32_t* h_b, size_t wb, bool on_device, uint32_t* h_seal) {
    const char* e = slot_check(p, slot, c);
    if (e) return e;
    if (!h_seal || !h_a || wa == 0) { set_error("b200: null argument"); return last_error(); }
    Slot& s = p->slots[slot];
    CU(cudaEventRecord(s.ev_begin, s.stream));
    return recursion_on_slot(p, s, c, h_a, wa, h_b, wb, on_device, h_seal);
}

// ---- the agent's task bodies as single enqueues (no host round trip between their steps) ----------------------------------------
// tasks::prove::prover (prover/crates/workflow/src/tasks/prove.rs:44-108): prove_segment -> verify_integrity -> lift -> verify_integrity.
// The lift reads the segment seal where the proof left it (the slot's seal buffer).  Outputs, all optional: the segment seal and the
// lifted seal in pinned host memory, the lifted seal in caller-owned device memory (for the join that follows, local or on a peer GPU),
// and the two verdicts (0 = valid; NULL skips both verifications).  Everything is valid after b200_prover_wait.
const char* b200_prove_lift_async(b200_prover* p, uint32_t slot, const b200_circuit* seg, uint64_t seed, const uint32_t* h_trace,
                                  const b200_circuit* lift, uint32_t* h_seg_seal, uint32_t* h_lift_seal, uint32_t* d_lift_seal,
                                  int* h_verdicts) {
    const char* e = slot_check(p, slot, seg);
    if (e) return e;
    const char* ce = check_circuit(lift);
    if (ce) { set_error("%s", ce); return last_error(); }
    if (lift->po2 > p->maxc.po2 || lift->w_code + lift->w_data + lift->w_accum > p->maxc.w_code + p->maxc.w_data + p->maxc.w_accum ||
        lift->w_accum > p->maxc.w_accum) { set_error("b200: lift circuit exceeds the prover's max_circuit"); return last_error(); }
    Slot& s = p->slots[slot];
    CU(cudaEventRecord(s.ev_begin, s.stream));
    KL(launch_set_globals(s.seal, seg->po2, seg->w_code, seg->w_data, seg->w_accum, seg->kind, seed, 1, s.stream));
    bool staged = false;
    if (h_trace && s.staged_src == h_trace && s.staged_words == ((size_t)(seg->w_code + seg->w_data) << seg->po2)) {
        CU(cudaStreamWaitEvent(s.stream, s.ev_staged, 0));
        uint32_t* t = s.coeffs; s.coeffs = s.coeffs_alt; s.coeffs_alt = t;
        staged = true;
    }
    s.staged_src = nullptr; s.staged_words = 0;
    if ((e = prove_on_slot(p, s, *seg, seed, false, h_trace, h_seg_seal, staged))) return e;
    const size_t seg_words = SealLayout(*seg).total;
    // the segment receipt is verified on an auxiliary stream (from a copy of the seal) WHILE the lift runs; both verdicts gate the task
    if (h_verdicts) { h_verdicts[0] = h_verdicts[1] = -1; if ((e = verify_aux(p, s, 0, *seg, s.seal, seg_words, true, &h_verdicts[0]))) return e; }
    if ((e = recursion_on_slot(p, s, lift, s.seal, seg_words, nullptr, 0, true, h_lift_seal))) return e;
    if (h_verdicts && (e = verify_on_slot(p, s, *lift, &h_verdicts[1], true))) return e;
    if (h_verdicts && (e = join_aux(s, 0))) return e;
    if (d_lift_seal) CU(cudaMemcpyAsync(d_lift_seal, s.seal, (size_t)SealLayout(*lift).total * 4, cudaMemcpyDeviceToDevice, s.stream));
    CU(cudaEventRecord(s.ev_end, s.stream));
    return nullptr;
}

// tasks::join::join (tasks/join.rs:41-79; union and resolve have the same shape): verify_integrity of the left and right receipt
// against the circuits the caller expects them to have, the recursion proof over both, verify_integrity of the result.  The children
// live in caller-owned DEVICE memory.  h_verdicts[3] = left, right, result (NULL skips the three verifications).
const char* b200_recursion_verified_async(b200_prover* p, uint32_t slot, const b200_circuit* c, const uint32_t* d_a,
                                          const b200_circuit* ca, const uint32_t* d_b, const b200_circuit* cb, uint32_t* h_seal,
                                          uint32_t* d_seal_out, int* h_verdicts) {
    const char* e = slot_check(p, slot, c);
    if (e) return e;
    if (!d_a || !ca || (d_b && !cb)) { set_error("b200: null argument"); return last_error(); }
    Slot& s = p->slots[slot];
    const b200_circuit* kids[2] = {ca, d_b ? cb : nullptr};
    const uint32_t* seals[2] = {d_a, d_b};
    size_t words[2] = {0, 0};
    for (int k = 0; k < 2; k++) {
        if (!kids[k]) continue;
        const char* ce = check_circuit(kids[k]);
        if (ce) { set_error("%s", ce); return last_error(); }
        words[k] = SealLayout(*kids[k]).total;
        if (words[k] > (size_t)SealLayout(p->maxc).total || kids[k]->po2 > p->maxc.po2) { set_error("b200: child circuit exceeds the prover's max_circuit"); return last_error(); }
    }
    CU(cudaEventRecord(s.ev_begin, s.stream));
    if (h_verdicts) {          // left and right are verified on the two auxiliary streams, beside the join proof
        h_verdicts[0] = h_verdicts[2] = -1; h_verdicts[1] = d_b ? -1 : 0;
        for (int k = 0; k < 2; k++)
            if (kids[k] && (e = verify_aux(p, s, k, *kids[k], seals[k], words[k], false, &h_verdicts[k]))) return e;
    }
    if ((e = recursion_on_slot(p, s, c, d_a, words[0], d_b, words[1], true, h_seal))) return e;
    if (h_verdicts && (e = verify_on_slot(p, s, *c, &h_verdicts[2], true))) return e;
    if (h_verdicts) for (int k = 0; k < 2; k++) if (kids[k] && (e = join_aux(s, k))) return e;
    if (d_seal_out) CU(cudaMemcpyAsync(d_seal_out, s.seal, (size_t)SealLayout(*c).total * 4, cudaMemcpyDeviceToDevice, s.stream));
    CU(cudaEventRecord(s.ev_end, s.stream));
    return nullptr;
}

const char* b200_recursion_async(b200_prover* p, uint32_t slot, const b200_circuit* c, const uint32_t* h_a, size_t wa,
                                 const uint32_t* h_b, size_t wb, uint32_t* h_seal) {
    return recursion_common(p, slot, c, h_a, wa, h_b, wb, false, h_seal);
}
// The same with the child seals in DEVICE memory (receipts that stay on the GPU between prove, lift and join, or arrive from a peer
// GPU over NVLink): no host staging.  The buffers are read on the slot's stream; the caller keeps them valid until b200_prover_wait.
const char* b200_recursion_dev_async(b200_prover* p, uint32_t slot, const b200_circuit* c, const uint32_t* d_a, size_t wa,
                                     const uint32_t* d_b, size_t wb, uint32_t* h_seal) {
    return recursion_common(p, slot, c, d_a, wa, d_b, wb, true, h_seal);
}

// Copy the seal the slot produced last (or is producing: the copy is ordered behind it on the slot's stream) into caller-owned device
// memory, so that a receipt can be handed to the next recursion step or to NCCL without leaving the GPU.
const char* b200_seal_to_device(b200_prover* p, uint32_t slot, uint32_t* d_dst, size_t words) {
    if (!p || slot >= p->slots.size()) { set_error("b200: bad prover/slot"); return last_error(); }
    Slot& s = p->slots[slot];
    if (!d_dst) { set_error("b200: null d_dst"); return last_error(); }
    if (!s.has_seal || words != (size_t)SealLayout(s.last_circuit).total) { set_error("b200: slot %u holds no seal of %zu words", slot, words); return last_error(); }
    CU(cudaSetDevice(p->device));
    CU(cudaMemcpyAsync(d_dst, s.seal, words * 4, cudaMemcpyDeviceToDevice, s.stream));
    return nullptr;
}

// 1 when everything enqueued on the slot has completed, 0 while it is still running, -1 on error (b200_last_error).  Never blocks:
// lets one host thread keep several slots and NCCL transfers in flight and react to whichever finishes first.
int b200_prover_query(b200_prover* p, uint32_t slot) {
    if (!p || slot >= p->slots.size()) { set_error("b200: bad prover/slot"); return -1; }
    if (cudaSetDevice(p->device) != cudaSuccess) { set_error("b200: cudaSetDevice failed"); return -1; }
    cudaError_t e = cudaStreamQuery(p->slots[slot].stream);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) return 0;
    set_error("b200: cudaStreamQuery: %s", cudaGetErrorString(e));
    return -1;
}

// verify_integrity against the circuit the CALLER expects (shape and kind): a seal whose header says anything else is rejected with
// code 103 instead of being verified as whatever it claims to be.  seal == NULL: the seal the slot produced last; otherwise `words`
// words in host memory (seal_on_device == 0) or caller-owned device memory (!= 0, read on the slot's stream).
const char* b200_verify_circuit_async(b200_prover* p, uint32_t slot, const b200_circuit* expect, const uint32_t* seal, size_t words,
                                      int seal_on_device, int* h_result) {
    if (!p || slot >= p->slots.size()) { set_error("b200: bad prover/slot"); return last_error(); }
    if (!h_result) { set_error("b200: null h_result"); return last_error(); }
    const char* ce = check_circuit(expect);
    if (ce) { set_error("%s", ce); return last_error(); }
    Slot& s = p->slots[slot];
    const b200_circuit c = *expect;
    if (c.po2 > p->maxc.po2 || c.w_code + c.w_data + c.w_accum > p->maxc.w_code + p->maxc.w_data + p->maxc.w_accum ||
        c.w_accum > p->maxc.w_accum || SealLayout(c).total > SealLayout(p->maxc).total) {
        set_error("b200: expected circuit exceeds the prover's max_circuit"); return last_error();
    }
    CU(cudaSetDevice(p->device));
    if (seal) {
        if (s.busy) { set_error("b200: slot %u busy (call b200_prover_wait first)", slot); return last_error(); }
        if ((size_t)SealLayout(c).total != words) { *h_result = 102; return nullptr; }
        if (seal_on_device) {
            CU(cudaMemcpyAsync(s.seal, seal, words * 4, cudaMemcpyDeviceToDevice, s.stream));
        } else {
            if (words > s.h_stage_words) { set_error("b200: seal too large for the staging buffer"); return last_error(); }
            memcpy(s.h_stage, seal, words * 4);
            CU(cudaMemcpyAsync(s.seal, s.h_stage, words * 4, cudaMemcpyHostToDevice, s.stream));
        }
        s.last_circuit = c; s.has_seal = true;
    } else if (!s.has_seal) {
        set_error("b200: slot %u holds no seal to verify", slot); return last_error();
    } else if ((size_t)SealLayout(c).total != (size_t)SealLayout(s.last_circuit).total) {
        *h_result = 102; return nullptr;
    }
    *h_result = -1;
    return verify_on_slot(p, s, c, h_result, true);
}

// verify_integrity: h_seal == NULL verifies the seal the slot produced last (still resident on the device)
const char* b200_verify_async(b200_prover* p, uint32_t slot, const uint32_t* h_seal, size_t words, int* h_result) {
    if (!p || slot >= p->slots.size()) { set_error("b200: bad prover/slot"); return last_error(); }
    if (!h_result) { set_error("b200: null h_result"); return last_error(); }
    Slot& s = p->slots[slot];
    b200_circuit c;
    if (h_seal) {
        if (s.busy) { set_error("b200: slot %u busy (call b200_prover_wait first)", slot); return last_error(); }
        if (words < (size_t)GLOBALS) { *h_result = 100; return nullptr; }
        c = b200_circuit{h_seal[0], h_seal[1], h_seal[2], h_seal[3], h_seal[4]};
        if (check_circuit(&c)) { *h_result = 101; return nullptr; }
        if ((size_t)SealLayout(c).total != words) { *h_result = 102; return nullptr; }
    } else {
        if (!s.has_seal) { set_error("b200: slot %u holds no seal to verify", slot); return last_error(); }
        c = s.last_circuit;
    }
    if (c.po2 > p->maxc.po2 || c.w_code + c.w_data + c.w_accum > p->maxc.w_code + p->maxc.w_data + p->maxc.w_accum ||
        c.w_accum > p->maxc.w_accum || SealLayout(c).total > SealLayout(p->maxc).total) {
        set_error("b200: seal's circuit exceeds the prover's max_circuit"); return last_error();
    }
    CU(cudaSetDevice(p->device));
    if (h_seal) {
        if (words > s.h_stage_words) { set_error("b200: seal too large for the staging buffer"); return last_error(); }
        memcpy(s.h_stage, h_seal, words * 4);
        CU(cudaMemcpyAsync(s.seal, s.h_stage, words * 4, cudaMemcpyHostToDevice, s.stream));
        s.last_circuit = c; s.has_seal = true;
    }
    *h_result = -1;
    return verify_on_slot(p, s, c, h_result);
}

const char* b200_prover_wait(b200_prover* p, uint32_t slot) {
    if (!p || slot >= p->slots.size()) { set_error("b200: bad prover/slot"); return last_error(); }
    Slot& s = p->slots[slot];
    s.busy = false;                       // whatever happens below, the slot is reusable afterwards
    const int nv = s.n_verdicts; s.n_verdicts = 0;
    CU(cudaSetDevice(p->device));
    CU(cudaStreamSynchronize(s.stream));
    for (int i = 0; i < nv; i++) if (s.user_verdict[i]) *s.user_verdict[i] = s.h_vpin[i];
    return nullptr;
}

float b200_prover_last_ms(b200_prover* p, uint32_t slot) {
    if (!p || slot >= p->slots.size()) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, p->slots[slot].ev_begin, p->slots[slot].ev_end) != cudaSuccess) return -1.f;
    return ms;
}

const char* b200_prover_mark(b200_prover* p, uint32_t slot, uint32_t which) {
    if (!p || slot >= p->slots.size() || which > 1) { set_error("b200: bad prover/slot/mark"); return last_error(); }
    CU(cudaSetDevice(p->device));
    Slot& s = p->slots[slot];
    if (!s.ev_mark[which]) CU(cudaEventCreate(&s.ev_mark[which]));
    CU(cudaEventRecord(s.ev_mark[which], s.stream));
    return nullptr;
}
float b200_prover_marks_ms(b200_prover* p, uint32_t slot_a, uint32_t which_a, uint32_t slot_b, uint32_t which_b) {
    if (!p || slot_a >= p->slots.size() || slot_b >= p->slots.size() || which_a > 1 || which_b > 1) return -1.f;
    cudaEvent_t a = p->slots[slot_a].ev_mark[which_a], b = p->slots[slot_b].ev_mark[which_b];
    float ms = -1.f;
    if (!a || !b || cudaEventSynchronize(b) != cudaSuccess || cudaEventElapsedTime(&ms, a, b) != cudaSuccess) return -1.f;
    return ms;
}

const char* b200_witgen_to_host(b200_prover* p, uint32_t slot, const b200_circuit* c, uint64_t seed, uint32_t* h_trace) {
    const char* e = slot_check(p, slot, c);
    if (e) return e;
    if (!h_trace) { set_error("b200: null h_trace"); return last_error(); }
    Slot& s = p->slots[slot];
    const size_t tw = (size_t)(c->w_code + c->w_data) << c->po2;
    KL(launch_gen_trace(s.coeffs, seed, nullptr, tw, s.stream));
    CU(cudaMemcpyAsync(h_trace, s.coeffs, tw * 4, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    return nullptr;
}

const char* b200_host_alloc(void** out, size_t bytes) {
    if (!out) return "b200: null out";
    CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return nullptr;
}
void b200_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
