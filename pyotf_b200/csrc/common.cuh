// Shared host-side helpers for the C ABI: error reporting and handle definitions.
#pragma once
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
};

namespace pyotf {
// per-dtype launchers (hanser_f32.cu / hanser_f64.cu)
template <typename T> int hanser_build_tables(pyotf_hanser* h, cudaStream_t s);
template <typename T>
int hanser_launch(const pyotf_hanser* h, const double* z, int nz, const void* pupilc, int batch,
                  long long pupil_stride, void* psfa, void* psfi, cudaStream_t s);
template <typename T>
int hanser_pack(const pyotf_hanser* h, const void* src, int batch, void* dst, cudaStream_t s);
template <typename T>
int fftn_run(const void* in, void* out, int ndim, const long long* shape, const int* axes, int naxes, int inverse,
             int real_in, int shift_in, int shift_out, cudaStream_t s);
template <typename T> int abs2_sum0_run(const void* a, int nvec, long long n, void* out, cudaStream_t s);
}  // namespace pyotf
