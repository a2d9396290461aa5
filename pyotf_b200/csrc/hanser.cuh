// K1: fused pupil -> defocus -> 2D inverse FFT -> |.|^2 per z-plane (HanserPSF._gen_psf,
// pyotf/otf.py:362-428, and BasePSF.PSFi, pyotf/otf.py:204-207), plus the fp64 pupil-plane
// geometry it consumes (HanserPSF._gen_kr/_gen_pupil, pyotf/otf.py:335-355 and the
// apodisation / vector factors of pyotf/otf.py:392-417).
//
// Decomposition: the pupil is non-zero only inside a K x K box of
// frequencies f in [f0, f0+K) around DC (K = 109 for the 256^2 configs).  One CTA owns the
// output rows y = R*y' + r (y' < M = N/R) of one z-plane:
//   fold   : F[b][kx] = sum_{ky == b mod M} pupil[ky][kx] * D_z[ky][kx] * w_N^{ky r}
//            (the first radix-R decimation-in-frequency stage, kept only for residue r)
//   cols   : U[y'][kx] = FFT_M over b                (shared memory, in place)
//   rows   : out[y][x] = FFT_N over kx of U[y'][.]   (registers + warp-private exchange)
//   epilog : |.|^2 / N^4, sum over polarisation components, store (fftshift folded into
//            the store index).
// No intermediate ever leaves the SM; the only HBM traffic is the PSFi/PSFa store.
#pragma once
#include "fft_regs.cuh"

namespace pyotf {

// ---------------------------------------------------------------------------------------
// plans: length L = R1 * R2, G = R2 threads cooperate on one transform
// ---------------------------------------------------------------------------------------
template <int L> struct FftPlan;
template <> struct FftPlan<16>  { static constexpr int R1 = 16, R2 = 1,  G = 1;  };
template <> struct FftPlan<32>  { static constexpr int R1 = 8,  R2 = 4,  G = 4;  };
template <> struct FftPlan<64>  { static constexpr int R1 = 8,  R2 = 8,  G = 8;  };
template <> struct FftPlan<128> { static constexpr int R1 = 16, R2 = 8,  G = 8;  };
template <> struct FftPlan<256> { static constexpr int R1 = 16, R2 = 16, G = 16; };

struct HanserGeomParams {
    int N;            // grid size
    int K;            // support box edge
    int f0;           // first frequency index of the box (negative)
    int nvec;         // 1, 2 or 6
    int vec_mode;     // 0 none, 1 x, 2 y, 3 z, 4 total
    int condition;    // 0 none, 1 sine, 2 herschel
    int apply_mask;   // 1: fold the ideal-pupil mask into amp (no user pupil)
    double kstep;     // 1.0 / (N * res)            (numpy.fft.fftfreq's `val`)
    double kmag;      // ni / wl
    double diff_limit;// na / wl
};

// fp64 geometry in the reference's operation order (no FMA contraction on the predicates).
template <typename T>
__global__ void hanser_geometry_kernel(HanserGeomParams p, double* __restrict__ kzc, T* __restrict__ ampc,
                                       unsigned char* __restrict__ maskc) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int KK = p.K * p.K;
    if (idx >= KK) return;
    int iy = idx / p.K, ix = idx - iy * p.K;
    double ky = __dmul_rn((double)(p.f0 + iy), p.kstep);
    double kx = __dmul_rn((double)(p.f0 + ix), p.kstep);
    double kr = hypot(ky, kx);                                  // cart2pol(kyy, kxx)  utils.py:25-29
    double phi = atan2(ky, kx);
    double kz2 = __dsub_rn(__dmul_rn(p.kmag, p.kmag), __dmul_rn(kr, kr));
    double kz = sqrt(fmax(0.0, kz2));                           // psqrt              otf.py:344
    bool inside = kr <= p.diff_limit;                           // _gen_pupil         otf.py:355
    double st = (kr < p.kmag) ? kr / p.kmag : 0.0;              // sin(theta)         otf.py:392
    double theta = asin(st);
    double ct = cos(theta);
    double a = 1.0;
    if (p.condition == 1) a = 1.0 / sqrt(ct);                   // otf.py:398
    else if (p.condition == 2) a = 1.0 / ct;                    // otf.py:400
    if (p.apply_mask && !inside) a = 0.0;
    kzc[idx] = kz;
    if (maskc) maskc[idx] = inside ? 1 : 0;
    if (p.vec_mode == 0) { ampc[idx] = (T)a; return; }
    double sp = sin(phi), cp = cos(phi), sth = sin(theta);
    int v = 0;
    if (p.vec_mode == 3 || p.vec_mode == 4) {                   // otf.py:407-409
        ampc[(size_t)(v++) * KK + idx] = (T)(a * (sth * cp));
        ampc[(size_t)(v++) * KK + idx] = (T)(a * (sth * sp));
    }
    if (p.vec_mode == 2 || p.vec_mode == 4) {                   // otf.py:410-412
        ampc[(size_t)(v++) * KK + idx] = (T)(a * ((ct - 1) * sp * cp));
        ampc[(size_t)(v++) * KK + idx] = (T)(a * (ct * sp * sp + cp * cp));
    }
    if (p.vec_mode == 1 || p.vec_mode == 4) {                   // otf.py:413-415
        ampc[(size_t)(v++) * KK + idx] = (T)(a * (ct * cp * cp + sp * sp));
        ampc[(size_t)(v++) * KK + idx] = (T)(a * ((ct - 1) * sp * cp));
    }
}

// twiddle table tw[k] = e^{+2 pi i k / N}, computed in fp64
template <typename T>
__global__ void twiddle_kernel(int N, cplx<T>* __restrict__ tw) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    double s, c;
    sincospi(2.0 * (double)k / (double)N, &s, &c);
    tw[k] = mk<T>((T)c, (T)s);
}

// user pupil (N x N complex128, unshifted like fftfreq) -> compact centred K x K in T
template <typename T>
__global__ void pack_pupil_kernel(int N, int K, int f0, const double2* __restrict__ src, cplx<T>* __restrict__ dst,
                                  int batch) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int KK = K * K;
    if (idx >= KK * batch) return;
    int b = idx / KK, r = idx - b * KK;
    int iy = r / K, ix = r - iy * K;
    int y = (f0 + iy) & (N - 1), x = (f0 + ix) & (N - 1);
    double2 v = src[(size_t)b * N * N + (size_t)y * N + x];
    dst[idx] = mk<T>((T)v.x, (T)v.y);
}

// bounding half-width (in frequency index) of the non-zero support of an unshifted N x N pupil
static __global__ void pupil_support_kernel(int N, const double2* __restrict__ src, int* __restrict__ hmax) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int h = 0;
    if (idx < N * N) {
        double2 v = src[idx];
        if (v.x != 0.0 || v.y != 0.0) {
            int y = idx / N, x = idx - y * N;
            int fy = y < N / 2 ? y : N - y, fx = x < N / 2 ? x : N - x;   // |f|, with -N/2 -> N/2
            h = max(fy, fx);
        }
    }
    for (int o = 16; o; o >>= 1) h = max(h, __shfl_xor_sync(0xffffffffu, h, o));
    if ((threadIdx.x & 31) == 0 && h > 0) atomicMax(hmax, h);
}

template <typename T> struct HanserArgs {
    const double* kzc;        // [K*K]   kz in cycles per length unit
    const T* ampc;            // [nvec][K*K]
    const cplx<T>* pupilc;    // [B][K*K] or nullptr
    const cplx<T>* tw;        // [N] e^{+2 pi i k/N}
    const double* z;          // [nz]
    T* psfi;                  // [B][nz][N][N] or nullptr
    cplx<T>* psfa;            // [B][nvec][nz][N][N] or nullptr
    int K, KP, f0, nz, nvec;
    long long pupil_stride;   // elements between pupils in pupilc (0 = shared)
};

template <int N, int M, typename T> struct HanserSmem {
    using RP = FftPlan<N>;
    static constexpr int NT = 256;
    static constexpr int NG = NT / RP::G;                     // row groups per CTA
    static constexpr int PITCH = RP::R2 + 1;                  // exchange pitch (elements)
    static constexpr int EXSZ = ((RP::R1 * PITCH + 7) / 8) * 8 + (RP::G == 8 ? 8 : 0);
    __host__ __device__ static size_t bytes(int KP) {
        return sizeof(cplx<T>) * ((size_t)N + (size_t)M * KP + (size_t)NG * EXSZ);
    }
};

template <typename T, int N, int M>
__global__ void __launch_bounds__(256) hanser_psf_kernel(HanserArgs<T> a) {
    using RP = FftPlan<N>;
    using CP = FftPlan<M>;
    using SM = HanserSmem<N, M, T>;
    constexpr int NT = SM::NT;
    constexpr int R = N / M;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* tws = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* U = tws + N;
    cplx<T>* EX = U + (size_t)M * a.KP;

    const int tid = threadIdx.x;
    const int r = blockIdx.x % R;
    const int plane = blockIdx.x / R;
    const int bidx = blockIdx.y;
    const int K = a.K, KP = a.KP, f0 = a.f0, KK = K * K;
    const double z = a.z[plane];
    const cplx<T>* pupil = a.pupilc ? a.pupilc + (size_t)bidx * a.pupil_stride : nullptr;

    for (int i = tid; i < N; i += NT) tws[i] = a.tw[i];

    // per-thread row-pass twiddles w_N^{b' c}, b' = lane in group
    const int g = tid % RP::G;
    const int grp = tid / RP::G;
    cplx<T> twr[RP::R1];
#pragma unroll
    for (int c = 0; c < RP::R1; ++c) twr[c] = a.tw[(g * c) & (N - 1)];

    const T scale = (T)1.0 / ((T)N * (T)N);
    const size_t plane_elems = (size_t)N * N;

    for (int v = 0; v < a.nvec; ++v) {
        __syncthreads();
        // ---------------- fold ----------------
        const T* amp = a.ampc + (size_t)v * KK;
        for (int w = tid; w < M * K; w += NT) {
            int b = w / K, kx = w - b * K;
            int off = (b - f0) % M; if (off < 0) off += M;
            cplx<T> acc = mk<T>(0, 0);
            for (int iy = off; iy < K; iy += M) {
                int idx = iy * K + kx;
                T am = amp[idx];
                cplx<T> p = pupil ? cscale(pupil[idx], am) : mk<T>(am, 0);
                if (p.x != (T)0 || p.y != (T)0) {
                    cplx<T> d;
                    cis_cycles(a.kzc[idx] * z, d);             // _calc_defocus   otf.py:357-360
                    p = cmul(p, d);
                    if (R > 1) p = cmul(p, tws[((f0 + iy) * r) & (N - 1)]);
                    acc = acc + p;
                }
            }
            U[b * KP + kx] = acc;
        }
        __syncthreads();
        // ---------------- columns: FFT_M over b, in place, digit-reversed rows ----------------
        {
            constexpr int Q1 = CP::R1, L1 = CP::R2;
            for (int w = tid; w < L1 * K; w += NT) {
                int bp = w / K, kx = w - bp * K;
                cplx<T> x[Q1];
#pragma unroll
                for (int q = 0; q < Q1; ++q) x[q] = U[(L1 * q + bp) * KP + kx];
                fft_radix<1, Q1>(x);
#pragma unroll
                for (int c = 0; c < Q1; ++c) {
                    if (L1 > 1 && c > 0) x[c] = cmul(x[c], tws[(bp * c * R) & (N - 1)]);
                    U[(c * L1 + bp) * KP + kx] = x[c];
                }
            }
            if constexpr (L1 > 1) {
                __syncthreads();
                for (int w = tid; w < Q1 * K; w += NT) {
                    int c1 = w / K, kx = w - c1 * K;
                    cplx<T> x[L1];
#pragma unroll
                    for (int q = 0; q < L1; ++q) x[q] = U[(c1 * L1 + q) * KP + kx];
                    fft_radix<1, L1>(x);
#pragma unroll
                    for (int q = 0; q < L1; ++q) U[(c1 * L1 + q) * KP + kx] = x[q];
                }
            }
        }
        __syncthreads();
        // ---------------- rows: FFT_N over kx + epilogue ----------------
        {
            constexpr int R1 = RP::R1, L1 = RP::R2, G = RP::G, NG = SM::NG, PITCH = SM::PITCH;
            constexpr int CL1 = CP::R2, CQ1 = CP::R1;
            cplx<T>* ex = EX + (size_t)grp * SM::EXSZ;
            for (int l = grp; l < M; l += NG) {
                const int yp = (l / CL1) + CQ1 * (l % CL1);
                const int y = R * yp + r;
                const int ys = (y + N / 2) & (N - 1);
                const cplx<T>* Ul = U + (size_t)l * KP;
                cplx<T> x[R1];
#pragma unroll
                for (int q = 0; q < R1; ++q) {
                    int k = L1 * q + g;
                    int f = k < N / 2 ? k : k - N;
                    int i = f - f0;
                    x[q] = (i >= 0 && i < K) ? Ul[i] : mk<T>(0, 0);
                }
                fft_radix<1, R1>(x);
                __syncwarp();
#pragma unroll
                for (int c = 0; c < R1; ++c) {
                    if (c > 0) x[c] = cmul(x[c], twr[c]);
                    ex[c * PITCH + g] = x[c];
                }
                __syncwarp();
#pragma unroll
                for (int u = 0; u < R1 / G; ++u) {
                    const int c1 = g + G * u;
                    cplx<T> t[L1];
#pragma unroll
                    for (int q = 0; q < L1; ++q) t[q] = ex[c1 * PITCH + q];
                    fft_radix<1, L1>(t);
#pragma unroll
                    for (int c2 = 0; c2 < L1; ++c2) {
                        const int xo = c1 + R1 * c2;
                        const int xs = (xo + N / 2) & (N - 1);
                        cplx<T> o = cscale(t[c2], scale);
                        if (a.psfa) {
                            size_t pa = (((size_t)bidx * a.nvec + v) * a.nz + plane) * plane_elems + (size_t)ys * N + xs;
                            a.psfa[pa] = o;
                        }
                        if (a.psfi) {
                            size_t pi = ((size_t)bidx * a.nz + plane) * plane_elems + (size_t)ys * N + xs;
                            T in = o.x * o.x + o.y * o.y;
                            if (v > 0) in += a.psfi[pi];
                            a.psfi[pi] = in;
                        }
                    }
                }
            }
        }
    }
}

}  // namespace pyotf
