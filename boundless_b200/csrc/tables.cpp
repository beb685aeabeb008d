// Per-device read-only tables (stage twiddles, six-step twiddle decomposition, zk_shift powers) and the
// host-side BabyBear helpers used to build them.  Product code: independent of oracle/.
#include "internal.h"
#include "constants.inc"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace b200 {

static constexpr uint32_t HP = 2013265921u, HPINV = 0x88000001u, HR2 = 1172168163u;

uint32_t h_mul(uint32_t a, uint32_t b) {
    uint64_t o = (uint64_t)a * b;
    uint32_t m = (uint32_t)o * HPINV;
    uint32_t t = (uint32_t)(((uint64_t)m * HP) >> 32);
    uint32_t hi = (uint32_t)(o >> 32);
    return hi >= t ? hi - t : hi - t + HP;
}
uint32_t h_add(uint32_t a, uint32_t b) { uint32_t s = a + b; return s >= HP ? s - HP : s; }
uint32_t h_sub(uint32_t a, uint32_t b) { return a >= b ? a - b : a - b + HP; }
uint32_t h_to_mont(uint32_t x) { return h_mul(x % HP, HR2); }
uint32_t h_from_mont(uint32_t a) { return h_mul(a, 1u); }
uint32_t h_pow(uint32_t a, uint64_t e) {
    uint32_t r = h_to_mont(1);
    while (e) { if (e & 1) r = h_mul(r, a); a = h_mul(a, a); e >>= 1; }
    return r;
}
uint32_t h_inv(uint32_t a) { return h_pow(a, HP - 2); }

static thread_local char g_err[1024];
const char* last_error() { return g_err; }
void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}

static std::mutex g_mu;
static DeviceTables* g_tables[64];

// tables are built in Montgomery form and uploaded as Shoup pairs {plain w, floor(w * 2^32 / p)}
static uint2* upload(const std::vector<uint32_t>& mont) {
    std::vector<uint2> v(mont.size());
    for (size_t i = 0; i < mont.size(); i++) {
        const uint32_t w = h_from_mont(mont[i]);
        v[i] = make_uint2(w, (uint32_t)(((uint64_t)w << 32) / HP));
    }
    uint2* d = nullptr;
    if (cudaMalloc(&d, v.size() * sizeof(uint2)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, v.data(), v.size() * sizeof(uint2), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
    return d;
}

const DeviceTables* get_tables(int device) {
    if (device < 0 || device >= 64) { set_error("b200: bad device %d", device); return nullptr; }
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_tables[device]) return g_tables[device];
    int prev = 0; cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) { set_error("b200: cudaSetDevice(%d) failed", device); return nullptr; }
    DeviceTables* T = new DeviceTables();
    memset(T, 0, sizeof *T);
    T->device = device;
    if (cudaDeviceGetAttribute(&T->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || T->sm_count <= 0) T->sm_count = 148;
    memcpy(T->rou_fwd, B200_ROU_FWD_MONT, sizeof T->rou_fwd);
    memcpy(T->rou_rev, B200_ROU_REV_MONT, sizeof T->rou_rev);
    bool ok = true;
    // stage twiddles: tw[2^(l-1) + i] = w_{2^l}^i for l <= 13
    const int TWL = 13;
    for (int inv = 0; inv < 2; inv++) {
        std::vector<uint32_t> tw((size_t)1 << TWL, h_to_mont(1));
        for (int l = 1; l <= TWL; l++) {
            uint32_t w = inv ? T->rou_rev[l] : T->rou_fwd[l], cur = h_to_mont(1);
            for (uint32_t i = 0; i < (1u << (l - 1)); i++) { tw[(1u << (l - 1)) + i] = cur; cur = h_mul(cur, w); }
        }
        uint2* d = upload(tw);
        ok &= d != nullptr;
        (inv ? T->tw_inv : T->tw_fwd) = d;
    }
    // six-step decomposition tables
    for (int m = 1; m <= MAX_LG && ok; m++) {
        const int h = (m + 1) / 2;
        for (int inv = 0; inv < 2; inv++) {
            const uint32_t w = inv ? T->rou_rev[m] : T->rou_fwd[m];
            std::vector<uint32_t> t(((size_t)1 << h) + ((size_t)1 << (m - h)));
            uint32_t cur = h_to_mont(1);
            for (uint32_t i = 0; i < (1u << h); i++) { t[i] = cur; cur = h_mul(cur, w); }
            const uint32_t wh = h_pow(w, (uint64_t)1 << h);
            // fold the normalisation into the inverse hi table: 1/2^m, or 1/2^BIG_N1 for the sizes that take the three-pass route
            cur = inv ? h_inv(h_to_mont(1u << (m > MAX_LG_2PASS ? BIG_N1 : m))) : h_to_mont(1);
            for (uint32_t i = 0; i < (1u << (m - h)); i++) { t[((size_t)1 << h) + i] = cur; cur = h_mul(cur, wh); }
            uint2* d = upload(t);
            ok &= d != nullptr;
            (inv ? T->pow_inv : T->pow_fwd)[m] = d;
        }
    }
    // zk_shift tables
    {
        std::vector<uint32_t> lo(4096), hi((size_t)1 << (MAX_LG - 12));
        const uint32_t three = h_to_mont(3), step = h_pow(three, 4096);
        uint32_t cur = h_to_mont(1);
        for (int i = 0; i < 4096; i++) { lo[i] = cur; cur = h_mul(cur, three); }
        cur = h_to_mont(1);
        for (size_t i = 0; i < hi.size(); i++) { hi[i] = cur; cur = h_mul(cur, step); }
        T->p3lo = upload(lo); T->p3hi = upload(hi);
        ok &= T->p3lo && T->p3hi;
    }
    cudaSetDevice(prev);
    if (!ok) { set_error("b200: table upload failed: %s", cudaGetErrorString(cudaGetLastError())); delete T; return nullptr; }
    g_tables[device] = T;
    return T;
}

static uint32_t h_bitrev(uint32_t x, uint32_t bits) {
    if (bits == 0) return 0;
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    x = (x >> 16) | (x << 16);
    return x >> (32 - bits);
}
static int env_int_t(const char* name, int dflt) { const char* v = getenv(name); return v && *v ? atoi(v) : dflt; }

// Host computation of a per-element table (Montgomery form, index = rho * 2^(lg_m - lg_rows) + pos); see internal.h for the three kinds.
// Separate from the upload so that tests/host_emul can check it against big-integer arithmetic without a device.
void fill_full_table(int kind, uint32_t lg_m, uint32_t lg_rows, const uint32_t* rou_fwd, const uint32_t* rou_rev, std::vector<uint32_t>& v) {
    const uint32_t lg_cols = lg_m - lg_rows;
    const size_t n = (size_t)1 << lg_m;
    v.resize(n);
    std::vector<uint32_t> lo, hi;
    if (kind == FULL_ZK) {
        lo.resize(4096); hi.resize(lg_m > 12 ? (size_t)1 << (lg_m - 12) : 1);
        const uint32_t three = h_to_mont(3), step = h_pow(three, 4096);
        uint32_t cur = h_to_mont(1);
        for (auto& x : lo) { x = cur; cur = h_mul(cur, three); }
        cur = h_to_mont(1);
        for (auto& x : hi) { x = cur; cur = h_mul(cur, step); }
        for (size_t i = 0; i < n; i++) {
            const uint32_t rho = (uint32_t)(i >> lg_cols), pos = (uint32_t)(i & (((size_t)1 << lg_cols) - 1));
            const uint32_t d = h_bitrev(rho, lg_rows) + (h_bitrev(pos, lg_cols) << lg_rows);
            v[i] = h_mul(lo[d & 4095], hi[d >> 12]);
        }
        return;
    }
    // the same two factors get_tables() uploads for this size (normalisation folded into the inverse hi factor)
    const bool inv = kind == FULL_INV;
    const uint32_t h = (lg_m + 1) / 2;
    const uint32_t w = inv ? rou_rev[lg_m] : rou_fwd[lg_m];
    lo.resize((size_t)1 << h); hi.resize((size_t)1 << (lg_m - h));
    uint32_t cur = h_to_mont(1);
    for (auto& x : lo) { x = cur; cur = h_mul(cur, w); }
    const uint32_t wh = h_pow(w, (uint64_t)1 << h);
    cur = inv ? h_inv(h_to_mont(1u << (lg_m > (uint32_t)MAX_LG_2PASS ? (uint32_t)BIG_N1 : lg_m))) : h_to_mont(1);
    for (auto& x : hi) { x = cur; cur = h_mul(cur, wh); }
    const uint32_t mmask = (uint32_t)(n - 1), lmask = (1u << h) - 1;
    for (size_t i = 0; i < n; i++) {
        const uint32_t rho = (uint32_t)(i >> lg_cols), pos = (uint32_t)(i & (((size_t)1 << lg_cols) - 1));
        const uint32_t e = (uint32_t)(((uint64_t)pos * h_bitrev(rho, lg_rows)) & mmask);
        v[i] = h_mul(lo[e & lmask], hi[e >> h]);
    }
}

const uint2* get_full_table(const DeviceTables* Tc, int kind, uint32_t lg_m, uint32_t lg_rows) {
    if (kind < 0 || kind > 2 || lg_m > (uint32_t)MAX_LG || lg_rows > lg_m) return nullptr;
    if (!env_int_t("B200_NTT_FULL", 1) || lg_m > (uint32_t)env_int_t("B200_NTT_FULL_MAX_LG", 22)) return nullptr;
    DeviceTables* T = const_cast<DeviceTables*>(Tc);
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (T->full[kind][lg_m]) return T->full_rows[kind][lg_m] == lg_rows ? T->full[kind][lg_m] : nullptr;
    std::vector<uint32_t> v;
    fill_full_table(kind, lg_m, lg_rows, T->rou_fwd, T->rou_rev, v);
    int prev = 0; cudaGetDevice(&prev);
    if (cudaSetDevice(T->device) != cudaSuccess) return nullptr;
    uint2* d = upload(v);
    cudaSetDevice(prev);
    if (!d) { cudaGetLastError(); return nullptr; }
    T->full[kind][lg_m] = d; T->full_rows[kind][lg_m] = lg_rows;
    return d;
}

// b200_shutdown(): release every device table (callers must have destroyed their provers first)
void free_tables() {
    std::lock_guard<std::mutex> lk(g_mu);
    int prev = 0; cudaGetDevice(&prev);
    for (int dev = 0; dev < 64; dev++) {
        DeviceTables* T = g_tables[dev];
        if (!T) continue;
        if (cudaSetDevice(dev) == cudaSuccess) {
            cudaFree(T->tw_fwd); cudaFree(T->tw_inv); cudaFree(T->p3lo); cudaFree(T->p3hi);
            for (int m = 0; m <= MAX_LG; m++) { cudaFree(T->pow_fwd[m]); cudaFree(T->pow_inv[m]); for (int k = 0; k < 3; k++) cudaFree(T->full[k][m]); }
        }
        delete T;
        g_tables[dev] = nullptr;
    }
    cudaSetDevice(prev);
}

}  // namespace b200
