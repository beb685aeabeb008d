// extern "C" kernel-level entry points (the risc0-sys / sppark FFI shape; include/b200zkp.h).
#include "../../include/b200zkp.h"
#include "internal.h"

using namespace b200;

static const DeviceTables* cur_tables() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("b200: no CUDA device available (this library has no CPU path)"); return nullptr; }
    return get_tables(dev);
}
#define RET(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error("b200: %s: %s", #call, cudaGetErrorString(e__)); return last_error(); } return nullptr; } while (0)
#define TABLES() const DeviceTables* T = cur_tables(); if (!T) return last_error()

extern "C" {

const char* b200_last_error(void) { return last_error(); }

int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return -1;
    return n;
}

// PCI bus id of a device ("0000:1b:00.0"), so that the host side can find the GPU's NUMA node in sysfs and keep its pinned witness
// buffers on that node (boundless_b200/feed.py bind_to_gpu_numa_node)
const char* b200_device_pci_bus_id(int device, char* out, int len) {
    if (!out || len < 16) { set_error("b200: pci bus id buffer too small"); return last_error(); }
    cudaError_t e = cudaDeviceGetPCIBusId(out, len, device);
    if (e != cudaSuccess) { set_error("b200: cudaDeviceGetPCIBusId(%d): %s", device, cudaGetErrorString(e)); return last_error(); }
    return nullptr;
}

const char* b200_init(int device) {
    int n = b200_device_count();
    if (n <= 0) { set_error("b200: no CUDA device available (this library has no CPU path)"); return last_error(); }
    if (device < 0 || device >= n) { set_error("b200: device %d out of range (%d visible)", device, n); return last_error(); }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("b200: cudaSetDevice(%d) failed", device); return last_error(); }
    return get_tables(device) ? nullptr : last_error();
}

const char* b200_batch_intt(uint32_t* d_io, uint32_t lg_n, uint32_t count, void* stream) {
    TABLES(); RET(launch_batch_intt(T, d_io, lg_n, count, (cudaStream_t)stream));
}
const char* b200_batch_ntt(uint32_t* d_io, uint32_t lg_n, uint32_t count, void* stream) {
    TABLES(); RET(launch_batch_ntt(T, d_io, lg_n, count, (cudaStream_t)stream));
}
const char* b200_batch_expand_ntt(uint32_t* d_out, const uint32_t* d_in, uint32_t lg_n, uint32_t lg_blowup, uint32_t count, void* stream) {
    TABLES(); RET(launch_batch_expand_ntt(T, d_out, d_in, lg_n, lg_blowup, count, (cudaStream_t)stream));
}
const char* b200_batch_zk_shift(uint32_t* d_io, uint32_t lg_n, uint32_t count, void* stream) {
    TABLES(); RET(launch_zk_shift(T, d_io, lg_n, count, (cudaStream_t)stream));
}
const char* b200_batch_intt_zk_shift(uint32_t* d_io, uint32_t lg_n, uint32_t count, void* stream) {
    TABLES(); RET(launch_batch_intt_shift(T, d_io, lg_n, count, (cudaStream_t)stream));
}
const char* b200_batch_bit_reverse(uint32_t* d_io, uint32_t lg_n, uint32_t count, void* stream) {
    RET(launch_bit_reverse(d_io, lg_n, count, (cudaStream_t)stream));
}
const char* b200_poseidon2_rows(uint32_t* d_out, const uint32_t* d_matrix, uint32_t rows, uint32_t cols, void* stream) {
    RET(launch_poseidon2_rows(d_out, d_matrix, rows, cols, rows, (cudaStream_t)stream));
}
const char* b200_poseidon2_fold(uint32_t* d_out, const uint32_t* d_in, uint32_t num_hashes, void* stream) {
    RET(launch_poseidon2_fold(d_out, d_in, num_hashes, (cudaStream_t)stream));
}
const char* b200_merkle_tree(uint32_t* d_nodes, const uint32_t* d_matrix, uint32_t lg_rows, uint32_t cols, void* stream) {
    if (lg_rows > 26) { set_error("b200: lg_rows too large"); return last_error(); }
    const uint32_t rows = 1u << lg_rows;
    cudaError_t e = launch_poseidon2_rows(d_nodes + (size_t)rows * 8, d_matrix, rows, cols, rows, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("b200: poseidon2_rows: %s", cudaGetErrorString(e)); return last_error(); }
    RET(launch_poseidon2_fold_tree(d_nodes, lg_rows, (cudaStream_t)stream));
}
const char* b200_fri_fold(uint32_t* d_out, const uint32_t* d_in, uint32_t in_size, const uint32_t* d_mix, void* stream) {
    RET(launch_fri_fold(d_out, d_in, in_size, d_mix, (cudaStream_t)stream));
}
size_t b200_evaluate_scratch_words(uint32_t lg_n, uint32_t count) { return evaluate_scratch_words(lg_n, count); }
const char* b200_batch_evaluate_any(uint32_t* d_out, const uint32_t* d_coeffs, uint32_t lg_n, uint32_t count, const uint32_t* d_x,
                                    uint32_t* d_scratch, void* stream) {
    RET(launch_evaluate(d_out, nullptr, d_coeffs, lg_n, count, d_x, nullptr, 0, 0, d_scratch, (cudaStream_t)stream));
}

}  // extern "C"
