// Zernike polynomials on the device (pyotf/zernike.py:251-365) and the aberrated-pupil
// builder of apply_aberration (pyotf/otf.py:626-642), batched over coefficient vectors for
// the PSF-library mode.  All arithmetic is fp64; only the final pupil is cast to T.
#pragma once
#include "fft_regs.cuh"

namespace pyotf {

// P_k^{(a,0)}(x): three-term recurrence (A&S 22.7.1, beta = 0) -- restates
// scipy.special.eval_jacobi(k, a, 0, x) used at zernike.py:332.
__device__ __forceinline__ double jacobi_a0(int k, int a, double x) {
    if (k == 0) return 1.0;
    double p0 = 1.0, p1 = 0.5 * ((double)a + (double)(a + 2) * x);
    for (int i = 2; i <= k; ++i) {
        double c = (double)(2 * i + a);
        double a1 = 2.0 * i * (double)(i + a) * (c - 2.0);
        double a2 = (c - 1.0) * (c * (c - 2.0) * x + (double)(a * a));
        double a3 = 2.0 * (double)(i + a - 1) * (double)(i - 1) * c;
        double p2 = (a2 * p1 - a3 * p0) / a1;
        p0 = p1; p1 = p2;
    }
    return p1;
}

__device__ __forceinline__ double ipow(double r, int m) {
    double out = 1.0, b = r;
    while (m) { if (m & 1) out *= b; b *= b; m >>= 1; }
    return out;
}

// One Zernike value Z_n^m(r, theta) (zernike.py:315-365); zero for r > 1 and for odd n-m.
__device__ __forceinline__ double zernike_value(double r, double theta, int n, int m, int norm) {
    int ma = m < 0 ? -m : m;
    if ((n - ma) & 1) return 0.0;
    if (!(r <= 1.0)) return 0.0;
    double rad;
    if (n == 0 && ma == 0) rad = 1.0;
    else {
        int k = (n - ma) / 2;
        rad = ipow(r, ma) * jacobi_a0(k, ma, 1.0 - 2.0 * r * r);
        if (k & 1) rad = -rad;
    }
    double ang = m < 0 ? sin((double)ma * theta) : cos((double)ma * theta);
    double z = rad * ang;
    if (norm) z *= (ma == 0) ? sqrt((double)(n + 1)) : sqrt(2.0 * (double)(n + 1));
    return z;
}

// out[mode][pt] for arbitrary (r, theta) point lists
static __global__ void zernike_eval_kernel(const double* __restrict__ r, const double* __restrict__ theta, long long npts,
                                           const int* __restrict__ n, const int* __restrict__ m, int nmodes, int norm,
                                           double* __restrict__ out) {
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    double rr = r[p], th = theta[p];
    for (int j = 0; j < nmodes; ++j) out[(size_t)j * npts + p] = zernike_value(rr, th, n[j], m[j], norm);
}

// Normal equations of the Zernike least-squares fit (phaseretrieval._fit_to_zerns, pyotf/phaseretrieval.py:406-432):
// G[i][j] = sum_p Z_i(p) Z_j(p), rhs[k][i] = sum_p Z_i(p) b_k(p) over the npts in-pupil pixels.  Z [nmodes][npts] is
// the output of zernike_eval_kernel, b [nrhs][npts] the data vectors (magnitude, phase).  One CTA per (i, j) with
// j <= nmodes + nrhs - 1: columns j >= nmodes are the right-hand sides.  fp64, fixed-order reduction.
static __global__ void __launch_bounds__(128) zernike_gram_kernel(const double* __restrict__ Z, const double* __restrict__ b,
                                                                  long long npts, int nmodes, int nrhs,
                                                                  double* __restrict__ G /* [nmodes][nmodes + nrhs] */) {
    __shared__ double red[4];
    const int i = blockIdx.y, j = blockIdx.x;
    const double* __restrict__ u = Z + (size_t)i * npts;
    const double* __restrict__ v = j < nmodes ? Z + (size_t)j * npts : b + (size_t)(j - nmodes) * npts;
    double acc = 0.0;
    for (long long p = threadIdx.x; p < npts; p += blockDim.x) acc += u[p] * v[p];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) G[(size_t)i * (nmodes + nrhs) + j] = (red[0] + red[1]) + (red[2] + red[3]);
}

struct AberrationParams {
    int N, K, KR, f0;
    int nmodes, batch;
    int has_mag, has_phase;
    double kstep, wl, na, diff_limit;
};

// pupilc[b][pix] = (sum_j mc[b][j] Z_j + [kr <= na/wl]) * exp(i sum_j pc[b][j] Z_j)
// One CTA owns PIX pixels of the compact K x K box: the Z_j values are computed once into
// shared memory and reused for every coefficient vector of the batch.
template <typename T, int PIX>
__global__ void __launch_bounds__(256) aberration_pupil_kernel(AberrationParams p, const int* __restrict__ zn,
                                                               const int* __restrict__ zm,
                                                               const double* __restrict__ mcoefs,
                                                               const double* __restrict__ pcoefs,
                                                               cplx<T>* __restrict__ pupilc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Z = reinterpret_cast<double*>(smem_raw);            // [nmodes][PIX]
    double* ideal = Z + (size_t)p.nmodes * PIX;                 // [PIX]
    const int KK = p.K * p.K;                 // real pixels; the padded layout has KR * K
    const int KKP = p.KR * p.K;
    const int pix0 = blockIdx.x * PIX;
    {
        // a thread evaluates its modes j = t / PIX, t / PIX + blockDim / PIX, ... for ONE pixel (PIX divides blockDim): the
        // polar coordinates are computed once per thread, not once per (pixel, mode)
        const int q = threadIdx.x % PIX, idx = pix0 + q;
        double r = 0.0, phi = 0.0;
        bool inpix = idx < KK;
        if (inpix) {
            const int iy = idx / p.K, ix = idx - iy * p.K;
            const double ky = __dmul_rn((double)(p.f0 + iy), p.kstep);
            const double kx = __dmul_rn((double)(p.f0 + ix), p.kstep);
            const double kr = hypot_ref(ky, kx);
            phi = atan2(ky, kx);
            r = __dmul_rn(kr, p.wl) / p.na;                      // otf.py:632
            if (threadIdx.x < PIX) ideal[q] = kr <= p.diff_limit ? 1.0 : 0.0;
        }
        for (int j = threadIdx.x / PIX; j < p.nmodes; j += blockDim.x / PIX)
            Z[j * PIX + q] = inpix ? zernike_value(r, phi, zn[j], zm[j], 1) : 0.0;
    }
    __syncthreads();
    // thread <-> (pixel q, batch lane); lanes of a warp cover consecutive pixels
    const int q = threadIdx.x % PIX;
    const int bl = threadIdx.x / PIX;
    const int nbl = blockDim.x / PIX;
    const int idx = pix0 + q;
    if (idx >= KKP) return;
    if (idx >= KK) {                                             // zero padding rows
        for (int b = blockIdx.y * nbl + bl; b < p.batch; b += nbl * gridDim.y)
            pupilc[(size_t)b * KKP + idx] = mk<T>(0, 0);
        return;
    }
    // EB entries at a time: EB independent accumulation chains per thread (one chain of nmodes dependent DFMAs per entry
    // left the kernel latency-bound)
    constexpr int EB = 4;
    const int bstep = nbl * gridDim.y;
    for (int b0 = blockIdx.y * nbl + bl; b0 < p.batch; b0 += EB * bstep) {
        double mag[EB], ph[EB];
#pragma unroll
        for (int e = 0; e < EB; ++e) { mag[e] = 0.0; ph[e] = 0.0; }
        if (p.has_mag) {
            for (int j = 0; j < p.nmodes; ++j) {
                const double zj = Z[j * PIX + q];
#pragma unroll
                for (int e = 0; e < EB; ++e) { const int b = b0 + e * bstep; if (b < p.batch) mag[e] += zj * mcoefs[(size_t)b * p.nmodes + j]; }
            }
        }
        if (p.has_phase) {
            for (int j = 0; j < p.nmodes; ++j) {
                const double zj = Z[j * PIX + q];
#pragma unroll
                for (int e = 0; e < EB; ++e) { const int b = b0 + e * bstep; if (b < p.batch) ph[e] += zj * pcoefs[(size_t)b * p.nmodes + j]; }
            }
        }
#pragma unroll
        for (int e = 0; e < EB; ++e) {
            const int b = b0 + e * bstep;
            if (b >= p.batch) continue;
            const double m = mag[e] + ideal[q];                      // otf.py:639
            double s, c;
            sincos(ph[e], &s, &c);
            pupilc[(size_t)b * KKP + idx] = mk<T>((T)(m * c), (T)(m * s));
        }
    }
}

// full unshifted grid: kr, phi, kz (HanserPSF._gen_kr, otf.py:335-344) for the API hooks
static __global__ void hanser_kspace_kernel(int N, double kstep, double kmag, double* __restrict__ kr_out,
                                            double* __restrict__ phi_out, double* __restrict__ kz_out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * N) return;
    int y = idx / N, x = idx - y * N;
    int fy = y < (N + 1) / 2 ? y : y - N, fx = x < (N + 1) / 2 ? x : x - N;
    double ky = __dmul_rn((double)fy, kstep), kx = __dmul_rn((double)fx, kstep);
    double kr = hypot_ref(ky, kx);
    if (kr_out) kr_out[idx] = kr;
    if (phi_out) phi_out[idx] = atan2(ky, kx);
    if (kz_out) kz_out[idx] = sqrt(fmax(0.0, __dsub_rn(__dmul_rn(kmag, kmag), __dmul_rn(kr, kr))));
}

// exp(2 pi i kz z) on the full grid (HanserPSF._calc_defocus, otf.py:357-360), complex128
static __global__ void hanser_defocus_kernel(int NN, const double* __restrict__ kz, const double* __restrict__ z, int nz,
                                             double2* __restrict__ out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int plane = blockIdx.y;
    if (idx >= NN || plane >= nz) return;
    cplx<double> d;
    cis_cycles(kz[idx] * z[plane], d);
    out[(size_t)plane * NN + idx] = make_double2(d.x, d.y);
}

}  // namespace pyotf
