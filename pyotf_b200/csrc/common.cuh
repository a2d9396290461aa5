// Shared host-side helpers for the C ABI: error reporting and handle definitions.
#pragma once
#include <vector>
#include <mutex>
#include <cuda_runtime.h>
#include <cstdio>
#include <string>

#include "../../include/pyotf_b200.h"

namespace pyotf {

void set_error(const std::string& msg);

#define PYOTF_CUDA_OK(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            ::pyotf::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
            return 2;                                                                            \
        }                                                                                        \
    } while (0)

#define PYOTF_REQUIRE(cond, msg)                                                                 \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            ::pyotf::set_error(std::string(msg));                                                \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

// Stream-ordered scratch (cudaMallocAsync) comes from the device's default pool; keep freed blocks cached
// instead of returning them to the driver at every synchronisation (the default release threshold is 0).
inline void keep_scratch_pool() {
    static thread_local int done_for = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done_for = dev;
}

constexpr int PYOTF_MAX_DEVICES = 64;
constexpr int PYOTF_MAX_PEERS = 16;
inline int current_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < PYOTF_MAX_DEVICES ? dev : 0;
}

// Pinned, device-mapped progress words of the phase-retrieval loops (two ints per slot: stopped, iterations finished).
// cudaHostAlloc / cudaFreeHost cost a fraction of a millisecond each and synchronise the device, so the slots come out
// of one page that lives as long as the process; a state borrows a slot for its lifetime.
struct ProgressSlots {
    static constexpr int NSLOT = 512;
    std::mutex mu;
    int* host = nullptr;
    void* dev = nullptr;
    std::vector<int> free_list;
    static ProgressSlots& get() { static ProgressSlots* p = new ProgressSlots; return *p; }   // never destroyed (no CUDA calls at exit)
    // returns a slot index, or -1 (out of slots / allocation failed: the caller allocates its own words)
    int acquire(int** h, void** d) {
        std::lock_guard<std::mutex> lk(mu);
        if (!host) {
            if (cudaHostAlloc(&host, sizeof(int) * 2 * NSLOT, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
                cudaHostGetDevicePointer(&dev, host, 0) != cudaSuccess) { cudaGetLastError(); host = nullptr; return -1; }
            for (int i = NSLOT - 1; i >= 0; --i) free_list.push_back(i);
        }
        if (free_list.empty()) return -1;
        const int i = free_list.back(); free_list.pop_back();
        *h = host + 2 * i; *d = (char*)dev + sizeof(int) * 2 * i;
        return i;
    }
    void release(int i) { std::lock_guard<std::mutex> lk(mu); free_list.push_back(i); }
};

constexpr int PYOTF_PAD_ROWS = 64;   // >= the largest rows-per-CTA M of K1
inline bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

}  // namespace pyotf

struct pyotf_hanser {
    pyotf_hanser_params p;
    int N, K, KR, KP, f0, nvec, device;   // KR = K + PYOTF_PAD_ROWS rows allocated (zero padded)
    double* kzc = nullptr;          // [K*K]
    void* ampc = nullptr;           // [nvec*K*K] dtype
    unsigned char* maskc = nullptr; // [K*K]
    void* tw = nullptr;             // [N] complex dtype
    double* octkz = nullptr;        // octant lists (symmetric box only): kz, amp of component 0, packed (i, j)
    void* octamp = nullptr;
    int* octij = nullptr;
    void* dtab = nullptr;           // [planes][dtab_size] per-plane defocus tables (scratch, grown on demand)
    size_t dtab_cap = 0;            // capacity of dtab in bytes
    int overlap = 0;                // streaming mode (pyotf_b200_hanser_set_overlap)
    void* pscratch = nullptr;       // pupil * amp of the last launch (scalar model), grown on demand
    size_t pscratch_cap = 0;
    double* zbuf = nullptr;         // plane positions uploaded by pyotf_b200_hanser_psf_zhost (stacks too long for the kernel arguments)
    size_t zbuf_cap = 0;
};

struct pyotf_comm;

// phase-retrieval state (K2); buffers are in the handle's dtype unless noted
struct pyotf_pr {
    pyotf_hanser* h = nullptr;      // owned geometry (support = -2, vec_corr / condition = none)
    int nz = 0, M = 0, nth = 256, nparts = 0, nmse = 0, nblk = 0, cap_iters = 0;
    int generic = 0;                // 1: four-pass path (sizes outside the fused kernel's set)
    void* W = nullptr;              // generic: [nz][K][N] complex
    void* P = nullptr;              // generic: [nz][N][N] complex
    long long nz_total = 0;         // planes over all ranks of a z-sharded retrieval
    int cur = 0, applied = 0;       // pupil buffer holding the final / the last applied pupil
    int launched = 0;               // iterations enqueued by the last run
    int kernel_launches = 0;        // fused loop: plane-kernel launches of the last run (each runs up to 8 iterations)
    double* z = nullptr;            // [nz]
    void* data = nullptr;           // [nz][N][N]
    void* pupil[2] = {nullptr, nullptr};   // [KR][K] complex
    void* partial = nullptr;        // [nparts][K][K] complex
    double* mse_part = nullptr;     // [nparts]
    void* sums = nullptr;           // [K*K + 1] double2 (all-reduce buffer)
    double* diffpart = nullptr;     // [2][nblk][2]
    double *mse = nullptr, *mse_diff = nullptr, *pupil_diff = nullptr;   // [cap_iters]
    void* dtab = nullptr;           // [nz][dtab_size] defocus tables of this stack (built once)
    void* state = nullptr;          // PrState
    unsigned* arrive = nullptr;     // grid-arrival counter of the fused (one kernel per iteration) loop
    int fuse = 0;                   // the plane kernel finalizes the iteration itself (single GPU, <= 1 CTA per SM)
    cudaStream_t stream = nullptr;  // creation stream: the state's buffers are stream-ordered allocations on it
    // z-sharded runs: peer-mapped window (CUDA IPC; layout in pr.cuh) into which every rank's finalize kernel pushes
    // its partial sums over NVLink
    void* win = nullptr;            // this rank's window (cudaMalloc, exported)
    size_t win_bytes = 0;
    void* peer[pyotf::PYOTF_MAX_PEERS] = {};   // every rank's window as mapped here (peer[win_rank] == win)
    int win_rank = 0, win_world = 0, win_alloc_world = 0;
    long long epoch_base = 0;       // epochs are monotonic over the life of the window
    int* h_prog = nullptr;          // pinned, mapped: [0] stopped, [1] iterations finished (written by the kernels)
    void* d_prog = nullptr;         // its device address
    int prog_slot = -1;             // >= 0: borrowed from pyotf::ProgressSlots; -1: own allocation
    int in_flight = 0;              // a run was enqueued and has not been seen to finish (it failed half-way)
};

namespace pyotf {
// K2 (pr_f32.cu / pr_f64.cu)
template <typename T> int pr_setup(pyotf_pr* s, const void* data, int data_dtype, int want_rows, cudaStream_t st);
template <typename T> int pr_set_pupil(pyotf_pr* s, const void* pupil_c128, cudaStream_t st);
template <typename T> int pr_get_pupil(const pyotf_pr* s, int which, void* out_c128, cudaStream_t st);
template <typename T>
int pr_run(pyotf_pr* s, int max_iters, double pupil_tol, double mse_tol, int phase_only, pyotf_comm* comm,
           double* mse_host, double* mse_diff_host, double* pupil_diff_host, int* iters_run_host, cudaStream_t st);
// per-dtype launchers (hanser_f32.cu / hanser_f64.cu)
template <typename T> int hanser_build_tables(pyotf_hanser* h, cudaStream_t s);
template <typename T>
int hanser_launch(const pyotf_hanser* h, const double* z, int nz, const void* pupilc, int batch,
                  long long pupil_stride, void* psfa, void* psfi, cudaStream_t s, const double* z_host = nullptr);
template <typename T>
int hanser_pack(const pyotf_hanser* h, const void* src, int batch, void* dst, cudaStream_t s);
template <typename T>
int fftn_run(const void* in, void* out, int ndim, const long long* shape, const int* axes, int naxes, int inverse,
             int real_in, int shift_in, int shift_out, cudaStream_t s, const unsigned char* const* pass_flags = nullptr);
template <typename T> int abs2_sum0_run(const void* a, int nvec, long long n, void* out, cudaStream_t s);
}  // namespace pyotf
