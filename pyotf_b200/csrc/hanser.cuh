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
#include <cooperative_groups.h>

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
    int KR;           // allocated rows of the compact tables (K + zero padding rows)
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
    int KK = p.KR * p.K;
    if (idx >= KK) return;
    int iy = idx / p.K, ix = idx - iy * p.K;
    if (iy >= p.K) {                                            // zero padding rows
        kzc[idx] = 0.0;
        if (maskc) maskc[idx] = 0;
        for (int v = 0; v < p.nvec; ++v) ampc[(size_t)v * KK + idx] = (T)0;
        return;
    }
    double ky = __dmul_rn((double)(p.f0 + iy), p.kstep);
    double kx = __dmul_rn((double)(p.f0 + ix), p.kstep);
    double kr = hypot_ref(ky, kx);                                  // cart2pol(kyy, kxx)  utils.py:25-29
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

// user pupil (N x N complex128, unshifted like fftfreq) -> compact centred [KR][K] in T
template <typename T>
__global__ void pack_pupil_kernel(int N, int K, int KR, int f0, const double2* __restrict__ src,
                                  cplx<T>* __restrict__ dst, int batch) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int KK = KR * K;
    if (idx >= KK * batch) return;
    int b = idx / KK, r = idx - b * KK;
    int iy = r / K, ix = r - iy * K;
    if (iy >= K) { dst[idx] = mk<T>(0, 0); return; }
    int y = ((f0 + iy) % N + N) % N, x = ((f0 + ix) % N + N) % N;       // fftfreq order, any N
    double2 v = src[(size_t)b * N * N + (size_t)y * N + x];
    dst[idx] = mk<T>((T)v.x, (T)v.y);
}

// pupil * amp for the scalar model, once per launch: every CTA of every plane then reads one table
// instead of two in its fold (a library launch has 256 CTAs per entry re-reading them from L2)
template <typename T>
__global__ void pupil_premul_kernel(const cplx<T>* __restrict__ pupil, const T* __restrict__ amp, cplx<T>* __restrict__ out,
                                    int KKP, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    out[i] = cscale(pupil[i], amp[(int)(i % KKP)]);
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

// plane positions that travel by value in the kernel arguments (pyotf_b200_hanser_psf_zhost)
#define PYOTF_Z_PARAM_MAX 256

template <typename T> struct HanserArgs {
    const double* kzc;        // [KR][K]   kz in cycles per length unit (rows >= K are zero)
    const T* ampc;            // [nvec][KR][K]
    const cplx<T>* pupilc;    // [B][KR][K] or nullptr
    const cplx<T>* tw;        // [N] e^{+2 pi i k/N}
    const cplx<T>* dtab;      // [nz][dtab_size] per-plane defocus tables (TAB variants) or nullptr
    int amp_in_tab;           // the tables already hold amp * D (ideal pupil, scalar model)
    int overlap;              // host-side: streaming mode, launch with programmatic stream serialization
    int pupil_has_amp;        // pupilc was pre-multiplied by the (single) apodisation table
    const double* octkz;      // octant lists for the in-kernel table builders (symmetric box) or nullptr
    const T* octamp;
    const int* octij;
    const double* z;          // [nz]
    T* psfi;                  // [B][nz][N][N] or nullptr
    cplx<T>* psfa;            // [B][nvec][nz][N][N] or nullptr
    int K, KR, KP, f0, nz, nvec;
    long long pupil_stride;   // elements between pupils in pupilc (0 = shared)
    int z_in_params;          // plane positions are in zv (host-side zrange, <= PYOTF_Z_PARAM_MAX planes), z unused
    double zv[PYOTF_Z_PARAM_MAX];
    // K1s only (default model): PSFi(-z) == PSFi(z), so planes at +-z are computed once and stored twice.  CTA c of the
    // npair launched per stack computes plane pa[c] and stores it to planes pa[c] and pb[c] (pb[c] == pa[c]: no partner).
    int npair;                // 0: no pairing, one CTA per plane
    unsigned char pa[PYOTF_Z_PARAM_MAX], pb[PYOTF_Z_PARAM_MAX];
};

template <int N, int M, typename T, int NTHREADS = 256> struct HanserSmem {
    using RP = FftPlan<N>;
    static constexpr int NT = NTHREADS;
    static constexpr int NG = NT / RP::G;                     // row groups per CTA
    static constexpr int PITCH = RP::R2 + 1;                  // exchange pitch (elements)
    static constexpr int EXSZ = ((RP::R1 * PITCH + 7) / 8) * 8 + (RP::G == 8 ? 8 : 0);
    static constexpr int EXTOT = NG * EXSZ;                   // elements of the exchange region (also stages the defocus table)
    // twiddles + column/row work area U + the row-exchange buffers
    __host__ __device__ static size_t bytes(int KP) {
        return sizeof(cplx<T>) * ((size_t)N + (size_t)M * KP + (size_t)NG * EXSZ);
    }
};

// Quadrant tables of the defocus factor, one per plane, in global memory (L2 / L1 resident):
//   dtab[plane][|fy|][|fx|] = exp(2 pi i kz(|fy|,|fx|) z[plane])   (* amp when `amp` is given: the scalar
//   model's apodisation is radially symmetric too), row pitch dtab_pitch(f0).
// kz depends on fy^2 + fx^2 only, so each octant entry (i >= j >= 0) is evaluated once and mirrored
// (HanserPSF._calc_defocus, pyotf/otf.py:357-360).  The phase kz*z (hundreds of cycles) is formed and
// reduced in fp64.  One tiny launch per stack serves every CTA of every batch entry; phase retrieval
// builds its tables once per measured stack.  Entries are dealt to threads by row pairs (rows p and
// H-p hold H+2 octant entries together), which keeps the index decode integer-only.
__host__ __device__ __forceinline__ int dtab_pitch(int f0) { return (1 - f0) | 1; }
__host__ __device__ __forceinline__ int dtab_size(int f0) { return (1 - f0) * dtab_pitch(f0); }

template <typename T>
__global__ void __launch_bounds__(256) hanser_dtab_kernel(const double* __restrict__ kzc, const T* __restrict__ amp,
                                                          const double* __restrict__ zr, int f0, int K,
                                                          cplx<T>* __restrict__ dtab) {
    const int H = -f0, HP = dtab_pitch(f0), W = H + 2, ntot = (H / 2 + 1) * W;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ntot) return;
    const int p = e / W, t = e - p * W;
    int i, j;
    if (t <= p) { i = p; j = t; } else { i = H - p; j = t - p - 1; }
    if (t > p && i == p) return;                       // the middle row of an even H is its own partner
    const unsigned o = (unsigned)((H + i) * K + (H + j));
    cplx<T> d;
    cis_cycles(kzc[o] * zr[blockIdx.y], d);
    if (amp) d = cscale(d, amp[o]);
    cplx<T>* __restrict__ tab = dtab + (size_t)blockIdx.y * dtab_size(f0);
    tab[i * HP + j] = d;
    tab[j * HP + i] = d;
}

// Fold fused with the first column-FFT stage, for one thread's column kx and sub-row bp.
// The pupil rows that reach this thread's Q1 bins (b = L1*q + bp) are exactly the rows with
// f0 + iy == bp (mod L1), an arithmetic progression iy = iy0 + L1*j.  Row j lands in bin
// q = (m0 + j) mod Q1, so it is accumulated into the compile-time slot xs[j mod Q1] and the
// cyclic offset m0 is repaid after the butterfly as the phase w_Q1^{m0 c} (shift theorem),
// merged into the stage twiddle.  Each group of Q1 rows is branch-free (clamped address,
// zero weight when out of range) so its global loads issue back to back.
// Octant lists for the in-kernel table builders: entry e (row-pair order, see hanser_dtab_kernel) ->
// kz, amp of component 0 and the packed (i << 16 | j) position.  Built once per handle so that the
// builders read three coalesced streams instead of decoding indices and gathering from the K x K tables.
template <typename T>
__global__ void hanser_oct_kernel(const double* __restrict__ kzc, const T* __restrict__ amp, int f0, int K,
                                  double* __restrict__ octkz, T* __restrict__ octamp, int* __restrict__ octij) {
    const int H = -f0, W = H + 2, ntot = (H / 2 + 1) * W;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ntot) return;
    const int p = e / W, t = e - p * W;
    int i, j;
    if (t <= p) { i = p; j = t; } else { i = H - p; j = t - p - 1; }
    if (t > p && i == p) { octij[e] = -1; octkz[e] = 0.0; octamp[e] = (T)0; return; }   // unused slot (even H)
    const unsigned o = (unsigned)((H + i) * K + (H + j));
    octkz[e] = kzc[o]; octamp[e] = amp[o]; octij[e] = (i << 16) | j;
}

// The defocus table of one plane built inside the PSF kernel.  R == 1: every CTA builds its own copy.
// R > 1: the R CTAs of the plane (a thread-block cluster) each evaluate 1/R of the entries and write
// them into the shared memory of every CTA of the cluster through distributed shared memory.
template <typename T, int NT, int R>
__device__ __forceinline__ void hanser_build_dtab_cluster(cplx<T>* Dq, const double* __restrict__ octkz,
                                                          const T* __restrict__ octamp, const int* __restrict__ octij,
                                                          bool with_amp, double z, int f0, int tid, int rank) {
    const int H = -f0, HP = dtab_pitch(f0), ntot = (H / 2 + 1) * (H + 2);
    cplx<T>* dst[R];
    if constexpr (R > 1) {
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
        for (int c = 0; c < R; ++c) dst[c] = cluster.map_shared_rank(Dq, c);
    } else {
        dst[0] = Dq;
    }
    constexpr int UNR = 4;
    for (int e0 = rank * NT + tid; e0 < ntot; e0 += UNR * R * NT) {
        double kz[UNR];
        T am[UNR];
        int ij[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {                 // all loads of the batch before the first sincos
            const int e = e0 + u * R * NT;
            const bool ok = e < ntot;
            ij[u] = ok ? octij[e] : -1;
            kz[u] = ok ? octkz[e] : 0.0;
            am[u] = (ok && with_amp) ? octamp[e] : (T)1;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (ij[u] >= 0) {
                const int i = ij[u] >> 16, j = ij[u] & 0xffff;
                cplx<T> d;
                cis_cycles(kz[u] * z, d);                // _calc_defocus   otf.py:357-360
                d = cscale(d, am[u]);
#pragma unroll
                for (int c = 0; c < R; ++c) { dst[c][i * HP + j] = d; dst[c][j * HP + i] = d; }
            }
        }
    }
}

// copy one plane's defocus table from global to shared memory, all of a thread's loads in flight at once
template <typename T, int NT>
__device__ __forceinline__ void stage_dtab(cplx<T>* __restrict__ dst, const cplx<T>* __restrict__ src, int n, int tid) {
    constexpr int UNR = 8;
    for (int i0 = tid; i0 < n; i0 += UNR * NT) {
        cplx<T> v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) { const int i = i0 + u * NT; v[u] = i < n ? src[i] : mk<T>(0, 0); }
#pragma unroll
        for (int u = 0; u < UNR; ++u) { const int i = i0 + u * NT; if (i < n) dst[i] = v[u]; }
    }
}

// what multiplies the defocus factor in the fold
enum FoldMode {
    FOLD_TAB_ONLY = 0,    // nothing: the table already holds amp * D      (ideal pupil, scalar model)
    FOLD_AMP = 1,         // amp                                           (ideal pupil, vectorial components)
    FOLD_PUPIL_AMP = 2,   // pupil * amp                                   (user / aberrated pupil)
    FOLD_PUPIL = 3        // pupil                                         (phase retrieval: unit apodisation)
};

template <bool TAB, int PM, typename T, int N, int M>
__device__ __forceinline__ void hanser_fold_cols(cplx<T>* __restrict__ U, const cplx<T>* __restrict__ tws,
                                                 const cplx<T>* __restrict__ Dq, const T* __restrict__ amp,
                                                 const cplx<T>* __restrict__ pupil, const double* __restrict__ kzc,
                                                 double z, int K, int KP, int f0, int r, int kx, int rl, int nrl) {
    using CP = FftPlan<M>;
    constexpr int R = N / M, Q1 = CP::R1, L1 = CP::R2, JMAX = N / L1;
    static_assert(TAB || PM != FOLD_TAB_ONLY, "FOLD_TAB_ONLY needs the table");
    const int fx = f0 + kx, ax = fx < 0 ? -fx : fx;
    const int H = -f0, HP = dtab_pitch(f0);
    const cplx<T>* __restrict__ dcol = Dq + ax;
    // rot[t] = w_R^{t r}: phase between consecutive groups of Q1 rows (they are M rows apart)
    cplx<T> rot[R];
#pragma unroll
    for (int t = 0; t < R; ++t) rot[t] = tws[(t * r * M) & (N - 1)];
    for (int bp = rl; bp < L1; bp += nrl) {
        const int iy0 = (bp - f0) & (L1 - 1);
        const int m0 = (f0 + iy0 - bp) / L1;                      // exact
        // one 32-bit element index serves every table (uniform base + index: one IMAD.WIDE per load)
        unsigned idx = (unsigned)(iy0 * K + kx);
        const unsigned step = (unsigned)(L1 * K);
        int fy = f0 + iy0;
        cplx<T> xs[Q1];
#pragma unroll
        for (int q = 0; q < Q1; ++q) xs[q] = mk<T>(0, 0);
#pragma unroll
        for (int jb = 0; jb < JMAX; jb += Q1) {
            if (iy0 + L1 * jb >= K) break;                        // warp-uniform
#pragma unroll
            for (int u = 0; u < Q1; ++u) {
                // rows >= K read the zero padding of the amp / pupil tables (KR = K + 64 rows)
                cplx<T> d;
                if constexpr (TAB) {
                    int ay = fy < 0 ? -fy : fy;
                    ay = ay < H ? ay : H;                         // padded rows: any in-table entry will do
                    d = dcol[ay * HP];
                } else {
                    cis_cycles(kzc[idx] * z, d);                   // _calc_defocus   otf.py:357-360
                }
                cplx<T> p;
                if constexpr (PM == FOLD_TAB_ONLY) {
                    p = d;
                    if (fy > H) p = mk<T>(0, 0);                  // warp-uniform: past the last pupil row
                } else if constexpr (PM == FOLD_AMP) {
                    p = cscale(d, amp[idx]);
                } else if constexpr (PM == FOLD_PUPIL_AMP) {
                    p = cmul(cscale(pupil[idx], amp[idx]), d);
                } else {
                    p = cmul(pupil[idx], d);
                }
                // w_N^{fy r} = w_N^{(bp + L1 (m0 + u)) r} * w_R^{(jb / Q1) r}: the second factor is
                // CTA-uniform per group, the first is applied once per slot after the loop
                if (jb == 0 || R == 1) xs[u] = xs[u] + p;
                else {
                    const cplx<T> w = rot[jb / Q1];
                    xs[u].x += p.x * w.x - p.y * w.y;
                    xs[u].y += p.x * w.y + p.y * w.x;
                }
                idx += step;
                fy += L1;
            }
        }
        if constexpr (R > 1) {
            if (r != 0) {                                         // CTA-uniform; w^0 = 1 for the residue-0 CTA
#pragma unroll
                for (int u = 0; u < Q1; ++u) xs[u] = cmul(xs[u], tws[((bp + L1 * (m0 + u)) * r) & (N - 1)]);
            }
        }
        fft_radix<1, Q1>(xs);
#pragma unroll
        for (int c = 0; c < Q1; ++c) {
            if (c > 0) xs[c] = cmul(xs[c], tws[(bp * c * R + m0 * c * (N / Q1)) & (N - 1)]);
            U[(c * L1 + bp) * KP + kx] = xs[c];
        }
    }
}

// Phase timestamps of K1 (debug builds only: -DPYOTF_K1_TIMING, tools/k1_phases.py); compiled out otherwise.
#ifdef PYOTF_K1_TIMING
static __device__ unsigned long long g_k1_marks[4096 * 8];
__device__ __forceinline__ unsigned long long k1_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define K1_MARK(i) do { if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 4096) g_k1_marks[blockIdx.x * 8 + (i)] = ((i) >= 6 ? k1_globaltimer() : (unsigned long long)clock64()); } while (0)
#else
#define K1_MARK(i) do { } while (0)
#endif

// TAB: the defocus factor exp(2 pi i kz z) depends on (|fy|, |fx|) only through fy^2 + fx^2, so
// when the support box is symmetric each CTA evaluates it once per octant entry (i >= j >= 0)
// into a shared-memory quadrant table instead of once per pupil pixel (6x fewer sincos for the
// 256^2 configs).
// TABM: 0 = defocus evaluated on the fly, 1 = per-plane tables precomputed in global memory and staged,
// 2 = table built by the plane's R-CTA cluster through distributed shared memory, 3 = every CTA builds
// its own copy (one stack per launch: cheaper than a separate table launch).
// SYM: ideal pupil of the scalar model.  pupil * amp * D then depends on (|ky|, |kx|) only, so the PSF is
// even in y and in x: the fold runs over the columns kx >= 0 only (the row pass reads column |kx|) and
// only the rows y <= N/2 are transformed, each stored to y and to N - y.  Halves both halves of the work
// for the reference's default model (HanserPSF with no aberration), bit-identical otherwise.
template <typename T, int N, int M, int TABM, bool SYM>
__global__ void __launch_bounds__(256, (M == 128 && sizeof(T) == 4 ? 2 : 0)) hanser_psf_kernel(HanserArgs<T> a) {
    constexpr bool TAB = TABM != 0;
    static_assert(!SYM || TAB, "the symmetric path needs the (symmetric-box) table");
    K1_MARK(6); K1_MARK(0);
    pdl_launch_dependents();          // a following K1 launch may be scheduled while this grid drains
    // default: wait for the preceding kernel before reading anything (only the launch latency overlaps);
    // streaming mode: wait only before the first global store (the whole compute phase overlaps)
    if (!a.overlap) pdl_wait();
    using RP = FftPlan<N>;
    using CP = FftPlan<M>;
    using SM = HanserSmem<N, M, T>;
    constexpr int NT = SM::NT;
    constexpr int R = N / M;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* tws = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* U = tws + N;
    cplx<T>* EX = U + (size_t)M * a.KP;
    // SYM: only the columns fx >= 0 (kx >= -f0) exist; the host passes the compact pitch in a.KP and every
    // access below keeps its absolute column index kx through this bias (half the work area: 128 rows per CTA
    // fit two CTAs per SM at 256^2)
    if constexpr (SYM) U += a.f0;

    const int tid = threadIdx.x;
    const int r = blockIdx.x % R;
    const int plane = blockIdx.x / R;
    const int bidx = blockIdx.y;
    // TAB: this plane's defocus table is staged into the (still idle) exchange region: the fold's
    // lookups are shared-memory hits instead of L2 round trips
    cplx<T>* Dq = EX;
    const int K = a.K, KP = a.KP, f0 = a.f0, KK = a.KR * K;
    const double z = a.z_in_params ? a.zv[plane] : a.z[plane];
    const cplx<T>* __restrict__ pupil = a.pupilc ? a.pupilc + (size_t)bidx * a.pupil_stride : nullptr;
    const double* __restrict__ kzc = a.kzc;

    for (int i = tid; i < N; i += NT) tws[i] = a.tw[i];

    // per-thread row-pass twiddles w_N^{b' c} / N^2 (the ifft normalisation rides on them)
    const int g = tid % RP::G;
    const int grp = tid / RP::G;
    const T scale = (T)1.0 / ((T)N * (T)N);
    cplx<T> twr[RP::R1];
#pragma unroll
    for (int c = 0; c < RP::R1; ++c) twr[c] = cscale(a.tw[(g * c) & (N - 1)], scale);

    // column-parallel mapping for fold + column FFT: a warp owns 32 consecutive kx
    const int kx0 = SYM ? -f0 : 0;                 // SYM: only the columns with fx >= 0
    const int ncols = K - kx0;
    const int ncw = (ncols + 31) >> 5;             // warps needed to cover the columns
    const int nrl = (NT / 32) / ncw;               // row lanes (>= 1 because K <= 256)
    const int kx = kx0 + ((tid >> 5) % ncw) * 32 + (tid & 31);
    const int rl = (tid >> 5) / ncw;
    const bool colactive = rl < nrl && kx < K;

    const size_t plane_elems = (size_t)N * N;
    K1_MARK(1);

    for (int v = 0; v < a.nvec; ++v) {
        if constexpr (TABM == 2) {
            // every CTA of the cluster is done with its exchange region before anyone writes into it
            cooperative_groups::this_cluster().sync();
            const bool amp_in = a.amp_in_tab != 0;
            hanser_build_dtab_cluster<T, NT, R>(Dq, a.octkz, a.octamp, a.octij, amp_in, z, f0, tid, r);
            cooperative_groups::this_cluster().sync();
        } else {
            __syncthreads();                                     // previous component's rows are done with EX
            if constexpr (TABM == 3) {
                hanser_build_dtab_cluster<T, NT, 1>(Dq, a.octkz, a.octamp, a.octij, a.amp_in_tab != 0, z, f0, tid, 0);
                __syncthreads();
            }
            if constexpr (TABM == 1) {
                stage_dtab<T, NT>(Dq, a.dtab + (size_t)plane * dtab_size(a.f0), dtab_size(a.f0), tid);
                __syncthreads();
            }
        }
        K1_MARK(2);
        // scalar model + ideal pupil: the (radially symmetric) apodisation rides in the defocus table
        const bool tab_only = TAB && a.amp_in_tab;
        const T* __restrict__ amp = a.ampc + (size_t)v * KK;
        // ---------------- fold + first column stage (registers), in place ----------------
        {
            constexpr int Q1 = CP::R1, L1 = CP::R2;
            if (colactive) {
                if (pupil && a.pupil_has_amp) hanser_fold_cols<TAB, FOLD_PUPIL, T, N, M>(U, tws, Dq, amp, pupil, kzc, z, K, KP, f0, r, kx, rl, nrl);
                else if (pupil) hanser_fold_cols<TAB, FOLD_PUPIL_AMP, T, N, M>(U, tws, Dq, amp, pupil, kzc, z, K, KP, f0, r, kx, rl, nrl);
                else if (TAB && tab_only) hanser_fold_cols<TAB, TAB ? FOLD_TAB_ONLY : FOLD_AMP, T, N, M>(U, tws, Dq, amp, pupil, kzc, z, K, KP, f0, r, kx, rl, nrl);
                else hanser_fold_cols<TAB, FOLD_AMP, T, N, M>(U, tws, Dq, amp, pupil, kzc, z, K, KP, f0, r, kx, rl, nrl);
            }
            K1_MARK(3);
            if constexpr (L1 > 1) {
                __syncthreads();
                if (colactive) {
                    for (int c1 = rl; c1 < Q1; c1 += nrl) {
                        cplx<T> x[L1];
#pragma unroll
                        for (int q = 0; q < L1; ++q) x[q] = U[(c1 * L1 + q) * KP + kx];
                        fft_radix<1, L1>(x);
#pragma unroll
                        for (int q = 0; q < L1; ++q) U[(c1 * L1 + q) * KP + kx] = x[q];
                    }
                }
            }
        }
        __syncthreads();
        // ---------------- rows: FFT_N over kx + epilogue ----------------
        // everything so far read per-handle constants and wrote shared memory only; the stores below must
        // not overtake those of an earlier launch into the same buffers (no-op for ordinary launches)
        K1_MARK(4);
        if (v == 0 && a.overlap) pdl_wait();
        {
            constexpr int R1 = RP::R1, L1 = RP::R2, G = RP::G, NG = SM::NG, PITCH = SM::PITCH;
            constexpr int CL1 = CP::R2, CQ1 = CP::R1;
            cplx<T>* ex = EX + (size_t)grp * SM::EXSZ;
            const int ipos = g - f0;                 // U index of frequency k = g (positive half)
            const int ineg = g - f0 - N;             // U index of k = g when k - N is meant (negative half)
            T* const psfi_plane = a.psfi ? a.psfi + ((size_t)bidx * a.nz + plane) * plane_elems : nullptr;
            cplx<T>* const psfa_plane =
                a.psfa ? a.psfa + (((size_t)bidx * a.nvec + v) * a.nz + plane) * plane_elems : nullptr;
            // SYM: the rows with y = R*y' + r <= N/2 only, i.e. y' < M/2 (plus y' = M/2 when r == 0)
            constexpr int HC = CL1 > 1 ? CL1 / 2 : 1;
            const int nrows = SYM ? M / 2 + (r == 0 ? 1 : 0) : M;
            for (int i = grp; i < nrows; i += NG) {
                int l = i;
                if constexpr (SYM) {
                    if constexpr (CL1 > 1) l = i < M / 2 ? (i / HC) * CL1 + (i % HC) : CL1 / 2;
                    else l = i;                                      // y' = l, rows 0 .. M/2
                }
                const int yp = (l / CL1) + CQ1 * (l % CL1);
                const int y = R * yp + r;
                const int ys = (y + N / 2) & (N - 1);
                const int ym = (N - y) & (N - 1);                    // mirror row (SYM)
                const bool mirror = SYM && ym != y;
                const int ysm = (ym + N / 2) & (N - 1);
                const cplx<T>* Ul = U + (size_t)l * KP;
                cplx<T> x[R1];
                // band-limited rows: a support box narrower than N/2 leaves the middle half of the frequencies
                // (k = L1*q + g with R1/4 <= q < 3 R1/4) empty, and the first radix-16 collapses (CTA-uniform)
                const bool narrow = R1 == 16 && -f0 < N / 4 && K + f0 <= N / 4;
#pragma unroll
                for (int q = 0; q < R1; ++q) {
                    if (R1 == 16 && narrow && q >= R1 / 4 && q < 3 * R1 / 4) continue;
                    // k = L1*q + g ; positive half k < N/2, else frequency k - N
                    if (L1 * q < N / 2) x[q] = (ipos + L1 * q < K) ? Ul[ipos + L1 * q] : mk<T>(0, 0);
                    else if constexpr (SYM) {                        // column |k - N| = N - k
                        const int af = N - L1 * q - g;
                        x[q] = (af <= -f0) ? Ul[af - f0] : mk<T>(0, 0);
                    }
                    else x[q] = (ineg + L1 * q >= 0) ? Ul[ineg + L1 * q] : mk<T>(0, 0);
                }
                if constexpr (R1 == 16) {
                    if (narrow) fft16_mid0<1>(x); else fft_radix<1, R1>(x);
                } else {
                    fft_radix<1, R1>(x);
                }
                __syncwarp();
#pragma unroll
                for (int c = 0; c < R1; ++c) ex[c * PITCH + g] = cmul(x[c], twr[c]);
                __syncwarp();
#pragma unroll
                for (int u = 0; u < R1 / G; ++u) {
                    const int c1 = g + G * u;
                    cplx<T> t[L1];
#pragma unroll
                    for (int q = 0; q < L1; ++q) t[q] = ex[c1 * PITCH + q];
                    fft_radix<1, L1>(t);
                    // output x = c1 + R1*c2, stored fftshift-ed: (x + N/2) mod N
                    if (psfa_plane) {
                        cplx<T>* row = psfa_plane + (size_t)ys * N + c1;
#pragma unroll
                        for (int c2 = 0; c2 < L1; ++c2)
                            row[R1 * c2 < N / 2 ? R1 * c2 + N / 2 : R1 * c2 - N / 2] = t[c2];
                        if (mirror) {
                            row = psfa_plane + (size_t)ysm * N + c1;
#pragma unroll
                            for (int c2 = 0; c2 < L1; ++c2)
                                row[R1 * c2 < N / 2 ? R1 * c2 + N / 2 : R1 * c2 - N / 2] = t[c2];
                        }
                    }
                    if (psfi_plane) {
                        T* row = psfi_plane + (size_t)ys * N + c1;
                        if (v == 0) {
#pragma unroll
                            for (int c2 = 0; c2 < L1; ++c2)
                                row[R1 * c2 < N / 2 ? R1 * c2 + N / 2 : R1 * c2 - N / 2] =
                                    t[c2].x * t[c2].x + t[c2].y * t[c2].y;
                            if (mirror) {
                                row = psfi_plane + (size_t)ysm * N + c1;
#pragma unroll
                                for (int c2 = 0; c2 < L1; ++c2)
                                    row[R1 * c2 < N / 2 ? R1 * c2 + N / 2 : R1 * c2 - N / 2] =
                                        t[c2].x * t[c2].x + t[c2].y * t[c2].y;
                            }
                        } else {
#pragma unroll
                            for (int c2 = 0; c2 < L1; ++c2)
                                row[R1 * c2 < N / 2 ? R1 * c2 + N / 2 : R1 * c2 - N / 2] +=
                                    t[c2].x * t[c2].x + t[c2].y * t[c2].y;
                        }
                    }
                }
            }
        }
    }
    K1_MARK(5); K1_MARK(7);
}

}  // namespace pyotf
