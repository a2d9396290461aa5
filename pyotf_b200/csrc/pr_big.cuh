// K2b: one phase-retrieval iteration (pyotf/phaseretrieval.py:95-130) for sizes the fused kernel K2 does
// not cover (N > 256 or not a power of two), as four line-FFT passes over L2-resident intermediates:
//   A  rows   : pupil * defocus, inverse FFT along kx                 -> W[plane][K][N]        (BigRow, hanser_big.cuh)
//   B  columns: inverse FFT along ky, /N^2, PSFi = |a|^2, (data - PSFi)^2 accumulated,
//               a <- sqrt(max(data,0)) a/|a|                          -> P[plane][N][N]        (unshifted)
//   C  columns: forward FFT along y, only the K box rows kept         -> W[plane][K][N]
//   D  rows   : forward FFT along x, only the K box columns kept, * conj(defocus)
//                                                                     -> partial[plane][K][K]
// followed by the same pr_finalize_kernel as the fused path (one partial pupil per plane).
#pragma once
#include "hanser_big.cuh"
#include "pr.cuh"

namespace pyotf {

template <typename T> struct PrBigArgs {
    const double* kzc;        // [KR][K]
    const double* z;          // [nz]
    const T* data;            // [nz][N][N] fftshift-ed
    cplx<T>* W;               // [nz][K][N]
    cplx<T>* P;               // [nz][N][N] unshifted new_psf
    cplx<T>* partial;         // [nz][K][K]
    double* mse_part;         // one per CTA of pass B
    const cplx<T>* tw;        // [N]
    const PrState* state;
    int N, K, f0;
    double scale;             // 1 / N^2
};

struct PrSkip {
    const PrState* st;
    __device__ __forceinline__ bool skip() const { return st->stopped != 0; }
};

// ---- pass B ---------------------------------------------------------------------------------------------
template <typename T> struct PrColInv {
    struct Line { const cplx<T>* w; const T* drow; cplx<T>* p; };
    using Acc = double;
    static constexpr int LOAD_BATCH = 8;
    static constexpr bool HAS_EMPTY = false;
    PrBigArgs<T> a;
    __device__ __forceinline__ static Acc acc_init() { return 0.0; }
    __device__ __forceinline__ bool skip() const { return a.state->stopped != 0; }
    __device__ __forceinline__ cplx<T> twiddle(int k) const { return a.tw[k]; }
    __device__ __forceinline__ Line info(long long line) const {
        const int x = (int)(line % a.N), plane = (int)(line / a.N);
        const int xs = (x + a.N / 2) % a.N;
        Line d;
        d.w = a.W + (size_t)plane * a.K * a.N + x;
        d.drow = a.data + (size_t)plane * a.N * a.N + xs;      // + ys * N per element
        d.p = a.P + (size_t)plane * a.N * a.N + x;             // + y * N per element
        return d;
    }
    __device__ __forceinline__ cplx<T> load(const Line& d, int l) const {
        const int f = l < (a.N + 1) / 2 ? l : l - a.N;
        const int iy = f - a.f0;
        if (iy < 0 || iy >= a.K) return mk<T>(0, 0);
        return d.w[(size_t)iy * a.N];
    }
    __device__ __forceinline__ void store(const Line& d, int o, cplx<T> v, Acc& acc) const {
        int ys = o + a.N / 2; if (ys >= a.N) ys -= a.N;          // data is stored fftshift-ed
        v = cscale(v, (T)a.scale);
        const T dv = d.drow[(size_t)ys * a.N];
        const T I = v.x * v.x + v.y * v.y;                        // model PSFi
        const T df = dv - I;
        acc += (double)df * (double)df;                           // _calc_mse  phaseretrieval.py:401-403
        d.p[(size_t)o * a.N] = replace_magnitude(v, I, dv);       // :79, :120-122
    }
    __device__ __forceinline__ void finish(Acc acc) const {
        __shared__ double red[8];
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
            a.mse_part[blockIdx.x] = s;
        }
    }
};

// ---- pass C ---------------------------------------------------------------------------------------------
template <typename T> struct PrColFwd : NoAcc {
    struct Line { const cplx<T>* p; cplx<T>* w; };
    PrBigArgs<T> a;
    __device__ __forceinline__ bool skip() const { return a.state->stopped != 0; }
    __device__ __forceinline__ cplx<T> twiddle(int k) const { return a.tw[k]; }
    __device__ __forceinline__ Line info(long long line) const {
        const int x = (int)(line % a.N), plane = (int)(line / a.N);
        Line d;
        d.p = a.P + (size_t)plane * a.N * a.N + x;
        d.w = a.W + (size_t)plane * a.K * a.N + x;
        return d;
    }
    __device__ __forceinline__ cplx<T> load(const Line& d, int l) const { return d.p[(size_t)l * a.N]; }
    __device__ __forceinline__ void store(const Line& d, int o, cplx<T> v, int&) const {
        const int f = o < (a.N + 1) / 2 ? o : o - a.N;
        const int iy = f - a.f0;
        if (iy >= 0 && iy < a.K) d.w[(size_t)iy * a.N] = v;
    }
};

// ---- pass D ---------------------------------------------------------------------------------------------
template <typename T> struct PrRowFwd : NoAcc {
    struct Line { const cplx<T>* w; cplx<T>* out; const double* kz; double z; };
    PrBigArgs<T> a;
    __device__ __forceinline__ bool skip() const { return a.state->stopped != 0; }
    __device__ __forceinline__ cplx<T> twiddle(int k) const { return a.tw[k]; }
    __device__ __forceinline__ Line info(long long line) const {
        const int iy = (int)(line % a.K), plane = (int)(line / a.K);
        Line d;
        d.w = a.W + (size_t)line * a.N;
        d.out = a.partial + ((size_t)plane * a.K + iy) * a.K;
        d.kz = a.kzc + (size_t)iy * a.K;
        d.z = a.z[plane];
        return d;
    }
    __device__ __forceinline__ cplx<T> load(const Line& d, int l) const { return d.w[l]; }
    __device__ __forceinline__ void store(const Line& d, int o, cplx<T> v, int&) const {
        const int f = o < (a.N + 1) / 2 ? o : o - a.N;
        const int ix = f - a.f0;
        if (ix < 0 || ix >= a.K) return;
        cplx<T> dph;
        cis_cycles(d.kz[ix] * d.z, dph);                          // new_pupils /= defocus   :126
        d.out[ix] = cmulc(v, dph);
    }
};

}  // namespace pyotf
