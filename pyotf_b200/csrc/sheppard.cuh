// K3 (generator): the amplitude OTF of SheppardPSF -- a spherical cap of radius ni/wl drawn on
// the 3D k-space grid (pyotf/otf.py:497-582).  One thread per voxel of the UNSHIFTED grid
// evaluates the reference's float64 predicates in the reference's operation order
//   kr   = norm((kz, ky, kx))                     otf.py:503
//   dkr  = norm((kz dkz, ky dk, kx dk)) / kr      otf.py:517-522   (dk when dk == dkz)
//   valid = |kr - kmag| < dkr  and  kzz > kz_min + dkr             otf.py:532
// and writes the value (1 for the scalar model, P * a for the vectorial components,
// otf.py:542-575) at the fftshift-ed position (otf.py:582).
#pragma once
#include "fft_regs.cuh"

namespace pyotf {

struct SheppardParams {
    int N, NZ, nvec, vec_mode, condition, dual, same_dk;
    double kstep, kzstep;     // fftfreq steps 1/(N res), 1/(NZ zres)
    double dk, dkz;           // k[1] - k[0], kz[1] - kz[0]
    double kmag, kz_min;
    double r2_lo, r2_hi;      // conservative bounds on kr^2 for a voxel to be on the shell (prefilter)
};

// fused PSFa / PSFi path (sheppard_impl.cuh, instantiated in sheppard_f32.cu / sheppard_f64.cu)
template <typename T> int sheppard_fused(const SheppardParams& p, void* psfa_out, void* psfi_out, cudaStream_t s);

__device__ __forceinline__ double fftfreq_at(int i, int n, double step) {
    // numpy.fft.fftfreq: integers [0, ..., (n-1)//2, -(n//2), ..., -1] * (1 / (n d))
    int f = i < (n + 1) / 2 ? i : i - n;
    return __dmul_rn((double)f, step);
}

// rowflag [NZ][N] / planeflag [NZ] (optional, zero-initialised, indexed in the SHIFTED layout of the output):
// set to 1 when the row / plane holds a shell voxel; the 3D inverse FFT skips lines that are entirely zero.
template <typename T>
__global__ void __launch_bounds__(256) sheppard_otf_kernel(SheppardParams p, cplx<T>* __restrict__ otf,
                                                           unsigned char* __restrict__ valid_out,
                                                           unsigned char* __restrict__ rowflag,
                                                           unsigned char* __restrict__ planeflag) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, zi = blockIdx.z;
    if (x >= p.N) return;
    const double kz = fftfreq_at(zi, p.NZ, p.kzstep);
    const double ky = fftfreq_at(y, p.N, p.kstep);
    const double kx = fftfreq_at(x, p.N, p.kstep);
    const size_t plane = (size_t)p.N * p.N, vol = plane * p.NZ;
    // fftshift: index i moves to (i + n//2) mod n
    int xs = x + p.N / 2, ys = y + p.N / 2, zs = zi + p.NZ / 2;
    if (xs >= p.N) xs -= p.N;
    if (ys >= p.N) ys -= p.N;
    if (zs >= p.NZ) zs -= p.NZ;
    const size_t o = (size_t)zs * plane + (size_t)ys * p.N + xs;
    // numpy.linalg.norm(axis=0): sqrt(add.reduce(x * x)) in array order (kz, ky, kx)
    const double kr2 = __dadd_rn(__dadd_rn(__dmul_rn(kz, kz), __dmul_rn(ky, ky)), __dmul_rn(kx, kx));
    // prefilter: |kr - kmag| < dkr <= max(dk, dkz) bounds kr^2; 99.7 % of the voxels stop here
    if (kr2 < p.r2_lo || kr2 > p.r2_hi) {
        if (valid_out) valid_out[(size_t)zi * plane + (size_t)y * p.N + x] = 0;
        if (otf) for (int v = 0; v < p.nvec; ++v) otf[(size_t)v * vol + o] = mk<T>(0, 0);
        return;
    }
    const double kr = sqrt(kr2);
    double dkr;
    if (p.same_dk) dkr = p.dk;
    else if (zi == 0 && y == 0 && x == 0) dkr = 0.0;
    else {
        const double a = __dmul_rn(kz, p.dkz), b = __dmul_rn(ky, p.dk), c = __dmul_rn(kx, p.dk);
        dkr = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c))) / kr;
    }
    const double kzz = p.dual ? fabs(kz) : kz;
    const bool valid = (fabs(__dsub_rn(kr, p.kmag)) < dkr) && (kzz > __dadd_rn(p.kz_min, dkr));
    if (valid_out) valid_out[(size_t)zi * plane + (size_t)y * p.N + x] = valid ? 1 : 0;
    if (valid && rowflag) { rowflag[(size_t)zs * p.N + ys] = 1; planeflag[zs] = 1; }
    if (!otf) return;
    if (!valid) {
        for (int v = 0; v < p.nvec; ++v) otf[(size_t)v * vol + o] = mk<T>(0, 0);
        return;
    }
    if (p.vec_mode == 0) { otf[o] = mk<T>((T)1, 0); return; }        // scalar model ignores `condition`  otf.py:577-579
    // direction cosines: (kx, ky, kz) / norm((kx, ky, kz))             otf.py:543
    const double kn = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(kx, kx), __dmul_rn(ky, ky)), __dmul_rn(kz, kz)));
    const double m = kx / kn, n = ky / kn, s = kz / kn;
    double a = 1.0;
    if (p.condition == 1) a = 1.0 / sqrt(s);
    else if (p.condition == 2) a = 1.0 / s;
    int v = 0;
    if (p.vec_mode == 3 || p.vec_mode == 4) {                         // otf.py:555-557
        otf[(size_t)(v++) * vol + o] = mk<T>((T)(-m * a), 0);
        otf[(size_t)(v++) * vol + o] = mk<T>((T)(-n * a), 0);
    }
    if (p.vec_mode == 2 || p.vec_mode == 4) {                         // otf.py:558-560
        otf[(size_t)(v++) * vol + o] = mk<T>((T)((-n * m / (1 + s)) * a), 0);
        otf[(size_t)(v++) * vol + o] = mk<T>((T)((1 - n * n / (1 + s)) * a), 0);
    }
    if (p.vec_mode == 1 || p.vec_mode == 4) {                         // otf.py:561-563
        otf[(size_t)(v++) * vol + o] = mk<T>((T)((1 - m * m / (1 + s)) * a), 0);
        otf[(size_t)(v++) * vol + o] = mk<T>((T)((-m * n / (1 + s)) * a), 0);
    }
}

// per-line flags of the first two passes of the 3D inverse FFT of OTFa [nvec][NZ][N][N]:
// pass over x: line (v, z, y) <- rowflag[z][y];  pass over y: line (v, z, x) <- planeflag[z]
static __global__ void sheppard_flags_kernel(int nvec, int NZ, int N, const unsigned char* __restrict__ rowflag,
                                             const unsigned char* __restrict__ planeflag,
                                             unsigned char* __restrict__ f_x, unsigned char* __restrict__ f_y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = nvec * NZ * N;
    if (i >= n) return;
    const int zy = i % (NZ * N);
    f_x[i] = rowflag[zy];
    f_y[i] = planeflag[zy / N];
}

}  // namespace pyotf
