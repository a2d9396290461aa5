// K2: one phase-retrieval iteration (Hanser 2003) = the loop body of retrieve_phase,
// pyotf/phaseretrieval.py:95-130, as two kernels:
//
//   pr_plane_kernel   (one CTA per (z-plane, row residue r), the same decomposition as K1)
//       pupil * defocus -> 2D inverse FFT (fold / columns / rows, all on chip)      :97  (model.PSFi)
//       per output pixel: PSFi = |PSFa|^2, (data - PSFi)^2 accumulated for the MSE   :98,401-403
//                         PSFa <- sqrt(max(data,0)) * PSFa / |PSFa|                  :120-122
//       2D forward FFT of the same rows (rows / columns / unfold: the exact transpose of
//       the inverse decomposition, evaluated only on the pupil support box)          :124
//       * conj(defocus)  (= / defocus, |defocus| = 1 on the box)                     :126
//       -> per-CTA partial pupil [K][K]  (no atomics: the sum is deterministic)
//   pr_finalize_kernel
//       mse[i], mse_diff[i], pupil_diff[i], convergence test                         :99-114
//       new_pupil = mean_z(partials) * mask, optional phase-only projection          :127-130
//       sum |old - new|^2 and sum |old|^2 for the next iteration's pupil_diff        :105
//
// The PSF plane never leaves the SM: HBM/L2 traffic per iteration is one read of `data`,
// one read of the compact pupil per CTA and one write + one read of the partial pupils.
#pragma once
#include "hanser.cuh"

namespace pyotf {

// Phase timestamps (debug builds only: -DPYOTF_PR_TIMING, tools/pr_phases.py): globaltimer marks per CTA
#ifdef PYOTF_PR_TIMING
static __device__ unsigned long long g_pr_marks[256 * 16];
#define PR_MARK(i) do { if (threadIdx.x == 0 && blockIdx.x < 256) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_pr_marks[blockIdx.x * 16 + (i)] = t_; } } while (0)
#else
#define PR_MARK(i) do { } while (0)
#endif

struct PrState {
    int stopped;        // set by the finalize kernel when a tolerance is met
    int iters_run;      // number of entries of mse / mse_diff / pupil_diff that are valid
    int final_is_next;  // 1: the loop ran out, the final pupil is one update past the last applied one
    int pad;
};

struct PrFinalArgs {
    int iter, K, nparts;
    int nmse;               // entries of mse_part (one per CTA of the kernel that compared model and data)
    int cur;                // which pupil buffer holds the pupil this iteration applied
    int phase_only, last_iter;
    double npix_total;      // nz_total * N * N   (mean of the squared error)
    double nz_total;        // planes over all ranks (mean over z of the new pupils)
    double pupil_tol, mse_tol;
    // pinned host words the deciding block writes over PCIe: [0] = PrState::stopped, [1] = iterations finished.  The
    // host follows the loop through them without putting a copy or an event between the launches (a device-to-host
    // copy every 8 iterations left the GPU idle for ~18 us each time: 2.3 us per iteration on the fixture).
    int* host_prog;
};

__device__ __forceinline__ void pr_publish_progress(int* host_prog, int done, int stopped) {
    if (!host_prog) return;
    volatile int* hp = host_prog;
    hp[1] = done;
    if (stopped) hp[0] = stopped;
    __threadfence_system();
}

template <typename T> struct PrPlaneArgs {
    const double* kzc;          // [KR][K]
    const cplx<T>* pupilc;      // [KR][K] current pupil (compact, zero padded rows)
    const cplx<T>* tw;          // [N]
    const cplx<T>* dtab;        // [nz][dtab_size] per-plane defocus tables (TAB variants) or nullptr
    const double* z;            // [nz]
    const T* data;              // [nz][N][N] measured stack, fftshift-ed like PSFi
    cplx<T>* partial;           // [nz * R][K][K]
    double* mse_part;           // [nz * R]
    const PrState* state;
    int K, KR, KP, f0, nz;
    int pdl;                    // launched with programmatic stream serialization (host-side flag only)
    // fused finalize (FUSE variants: every CTA of the grid is resident; see pr_fused_finalize)
    unsigned* arrive;           // grid-arrival counter, monotonic over the iterations of a run (zeroed by pr_run)
    PrFinalArgs fin;            // iter / cur / last_iter describe the FIRST iteration of this launch
    int niter;                  // iterations this launch runs back to back (FUSE: a second grid barrier separates them)
    const unsigned char* mask;
    cplx<T>* pupil0; cplx<T>* pupil1;
    double* diffpart;           // [2][gridDim.x][2]
    double *mse, *mse_diff, *pupil_diff;
    PrState* state_rw;
};

// new_psf = sqrt(max(data, 0)) * exp(i angle(a))  (psqrt, phaseretrieval.py:79; :120-122), I = |a|^2.
// angle(0) = 0, so a vanishing model amplitude gives the real value sqrt(data).
// fp64: exact division / square root.  fp32: sqrt.approx * rsqrt.approx (2^-22 relative error, no
// slow-path subroutine: 78 % of the fixture's pixels are zero after background subtraction and
// 0 / x takes the IEEE division's slow path), with tiny amplitudes rescaled so that I stays normal.
__device__ __forceinline__ cplx<double> replace_magnitude(cplx<double> a, double I, double data) {
    const double m = sqrt(fmax(data, 0.0));
    if (I > 0.0) return cscale(a, m / sqrt(I));
    return mk<double>(m, 0.0);
}
__device__ __forceinline__ cplx<float> replace_magnitude(cplx<float> a, float I, float data) {
    float m, r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(m) : "f"(fmaxf(data, 0.0f)));
    if (I < 1e-30f) {                                  // |a| < 1e-15: rescale (or a == 0)
        a.x *= 1e20f; a.y *= 1e20f;
        I = a.x * a.x + a.y * a.y;
        if (!(I > 0.0f)) return mk<float>(m, 0.0f);
    }
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(I));
    const float s = m * r;
    return mk<float>(a.x * s, a.y * s);
}

// NTH = 256 (two CTAs per SM) or 512 (one wide CTA per SM: short stacks, where there are fewer plane
// quarters than SMs and latency hiding has to come from inside the CTA).
template <typename T, int NT> __device__ int pr_fused_finalize(const PrPlaneArgs<T>& a, int it, double* red);

template <typename T, int N, int M, bool TAB, int NTH, bool FUSE = false>
__global__ void __launch_bounds__(NTH, (sizeof(T) == 4 && NTH == 256) ? 2 : 1) pr_plane_kernel(PrPlaneArgs<T> a) {
    using RP = FftPlan<N>;
    using CP = FftPlan<M>;
    using SM = HanserSmem<N, M, T, NTH>;
    constexpr int NT = SM::NT;
    constexpr int R = N / M;
    PR_MARK(0);
    pdl_launch_dependents();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[NT / 32];
    cplx<T>* tws = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* U = tws + N;
    cplx<T>* EX = U + (size_t)M * a.KP;

    const int tid = threadIdx.x;
    const int r = blockIdx.x % R;
    const int plane = blockIdx.x / R;
    cplx<T>* Dq = EX;                                            // staged defocus table (aliases EX, disjoint phases)
    const cplx<T>* __restrict__ dsrc = TAB ? a.dtab + (size_t)plane * dtab_size(a.f0) : nullptr;
    if constexpr (TAB) stage_dtab<T, NT>(Dq, dsrc, dtab_size(a.f0), tid);
    const int K = a.K, KP = a.KP, f0 = a.f0;
    const double z = a.z[plane];
    const cplx<T>* pupil = a.pupilc;
    const double* __restrict__ kzc = a.kzc;

    for (int i = tid; i < N; i += NT) tws[i] = a.tw[i];

    const int g = tid % RP::G;
    const int grp = tid / RP::G;
    const T scale = (T)1.0 / ((T)N * (T)N);
    cplx<T> twr[RP::R1];
#pragma unroll
    for (int c = 0; c < RP::R1; ++c) twr[c] = cscale(a.tw[(g * c) & (N - 1)], scale);

    const int ncw = (K + 31) >> 5;
    const int nrl = (NT / 32) / ncw;
    const int kx = ((tid >> 5) % ncw) * 32 + (tid & 31);
    const int rl = (tid >> 5) / ncw;
    const bool colactive = rl < nrl && kx < K;

    // everything above reads per-stack constants only; the pupil, the state word and the partial-pupil
    // buffer belong to the previous finalize kernel until it has completed
    pdl_wait();
    PR_MARK(1);
    if (a.state->stopped) return;
    __syncthreads();
    // FUSE: the launch runs a.niter iterations; the in-kernel finalize ends each one behind a grid-wide barrier and a
    // second barrier publishes the new pupil.  tws and the staged defocus table survive an iteration (the forward
    // column phase restages the table once EX is free; the finalize's scratch lies in U).
    for (int it = 0;; ++it) {
    if constexpr (FUSE) { if (it) pupil = ((a.fin.cur + it) & 1) ? a.pupil1 : a.pupil0; }
    // ---------------- inverse: fold + columns ----------------
    {
        constexpr int Q1 = CP::R1, L1 = CP::R2;
        if (colactive) hanser_fold_cols<TAB, FOLD_PUPIL, T, N, M>(U, tws, Dq, nullptr, pupil, kzc, z, K, KP, f0, r, kx, rl, nrl);
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
    PR_MARK(2);
    // ---------------- rows: inverse FFT_N, magnitude replacement, forward FFT_N ----------------
    double mse_acc = 0.0;
    {
        constexpr int R1 = RP::R1, L1 = RP::R2, G = RP::G, NG = SM::NG, PITCH = SM::PITCH;
        constexpr int CL1 = CP::R2, CQ1 = CP::R1;
        cplx<T>* ex = EX + (size_t)grp * SM::EXSZ;
        const int ipos = g - f0;
        const int ineg = g - f0 - N;
        const T* const dplane = a.data + (size_t)plane * N * N;
        for (int l = grp; l < M; l += NG) {
            const int yp = (l / CL1) + CQ1 * (l % CL1);
            const int ys = (R * yp + r + N / 2) & (N - 1);
            cplx<T>* Ul = U + (size_t)l * KP;
            const T* const drow = dplane + (size_t)ys * N;
            cplx<T> x[R1];
#pragma unroll
            for (int q = 0; q < R1; ++q) {
                if (L1 * q < N / 2) x[q] = (ipos + L1 * q < K) ? Ul[ipos + L1 * q] : mk<T>(0, 0);
                else x[q] = (ineg + L1 * q >= 0) ? Ul[ineg + L1 * q] : mk<T>(0, 0);
            }
            fft_radix<1, R1>(x);
            __syncwarp();
#pragma unroll
            for (int c = 0; c < R1; ++c) ex[c * PITCH + g] = cmul(x[c], twr[c]);
            __syncwarp();
            T macc = 0;
            cplx<T> t[R1 / G][L1];
#pragma unroll
            for (int u = 0; u < R1 / G; ++u) {
                const int c1 = g + G * u;
                T dv[L1];
#pragma unroll
                for (int c2 = 0; c2 < L1; ++c2)                  // data is stored fftshift-ed
                    dv[c2] = drow[(c1 + R1 * c2 + N / 2) & (N - 1)];
#pragma unroll
                for (int q = 0; q < L1; ++q) t[u][q] = ex[c1 * PITCH + q];
                fft_radix<1, L1>(t[u]);                          // PSFa at x = c1 + R1*c2 (unshifted)
#pragma unroll
                for (int c2 = 0; c2 < L1; ++c2) {
                    const T I = t[u][c2].x * t[u][c2].x + t[u][c2].y * t[u][c2].y;   // model PSFi
                    const T d = dv[c2] - I;
                    macc += d * d;                               // _calc_mse  phaseretrieval.py:401-403
                    t[u][c2] = replace_magnitude(t[u][c2], I, dv[c2]);   // :79, :120-122
                }
            }
            mse_acc += (double)macc;
            // forward FFT_N over x (fftn(ifftshift(new_psf)): the registers already hold the unshifted order)
            __syncwarp();
#pragma unroll
            for (int u = 0; u < R1 / G; ++u) {
                const int c1 = g + G * u;
                fft_radix<-1, L1>(t[u]);                         // over c2 -> g'
#pragma unroll
                for (int q = 0; q < L1; ++q)
                    ex[c1 * PITCH + q] = q ? cmulc(t[u][q], tws[(c1 * q) & (N - 1)]) : t[u][q];
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < R1; ++c) x[c] = ex[c * PITCH + g];
            fft_radix<-1, R1>(x);                                // over c1 -> q ; k = L1*q + g
#pragma unroll
            for (int q = 0; q < R1; ++q) {
                if (L1 * q < N / 2) { if (ipos + L1 * q < K) Ul[ipos + L1 * q] = x[q]; }
                else { if (ineg + L1 * q >= 0) Ul[ineg + L1 * q] = x[q]; }
            }
        }
    }
    __syncthreads();
    PR_MARK(3);
    // ---------------- forward columns: FFT_M over y' (transpose of the inverse stages) ----------------
    {
        constexpr int Q1 = CP::R1, L1 = CP::R2, JMAX = N / L1;
        if constexpr (TAB) stage_dtab<T, NT>(Dq, dsrc, dtab_size(a.f0), tid);                  // EX is free again
        if constexpr (L1 > 1) {
            if (colactive) {
                for (int c1 = rl; c1 < Q1; c1 += nrl) {
                    cplx<T> x[L1];
#pragma unroll
                    for (int q = 0; q < L1; ++q) x[q] = U[(c1 * L1 + q) * KP + kx];
                    fft_radix<-1, L1>(x);                        // over d -> b'
#pragma unroll
                    for (int q = 0; q < L1; ++q) U[(c1 * L1 + q) * KP + kx] = x[q];
                }
            }
        }
        __syncthreads();
        if (colactive) {
            const int fx = f0 + kx, ax = fx < 0 ? -fx : fx;
            const int HP = dtab_pitch(f0);
            cplx<T>* __restrict__ outp = a.partial + (size_t)blockIdx.x * K * K + kx;
            for (int bp = rl; bp < L1; bp += nrl) {
                const int iy0 = (bp - f0) & (L1 - 1);
                const int m0 = (f0 + iy0 - bp) / L1;
                cplx<T> xs[Q1];
#pragma unroll
                for (int c = 0; c < Q1; ++c) {
                    xs[c] = U[(c * L1 + bp) * KP + kx];
                    // stage twiddle w_M^{-b' c} and the cyclic shift by m0 bins (shift theorem)
                    if (c > 0) xs[c] = cmulc(xs[c], tws[(bp * c * R + m0 * c * (N / Q1)) & (N - 1)]);
                }
                fft_radix<-1, Q1>(xs);                           // over c -> q ; slot u holds bin b' + L1*((m0+u) mod Q1)
#pragma unroll
                for (int jb = 0; jb < JMAX; jb += Q1) {
                    if (iy0 + L1 * jb >= K) break;
#pragma unroll
                    for (int u = 0; u < Q1; ++u) {
                        const int iy = iy0 + L1 * (jb + u);
                        if (iy < K) {
                            const int fy = f0 + iy;
                            cplx<T> d;
                            if constexpr (TAB) {
                                const int ay = fy < 0 ? -fy : fy;
                                d = Dq[ay * HP + ax];
                            } else {
                                cis_cycles(kzc[iy * K + kx] * z, d);
                            }
                            // unfold: residue-r share of bin fy is w_N^{-fy r} G_r[fy mod M]; then / defocus
                            cplx<T> v = xs[u];
                            if constexpr (R > 1) v = cmulc(v, tws[(fy * r) & (N - 1)]);
                            outp[(size_t)iy * K] = cmulc(v, d);
                        }
                    }
                }
            }
        }
    }
    PR_MARK(4);
    // ---------------- block reduction of the squared error ----------------
    for (int o = 16; o; o >>= 1) mse_acc += __shfl_xor_sync(0xffffffffu, mse_acc, o);
    if ((tid & 31) == 0) red[tid >> 5] = mse_acc;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < NT / 32; ++w) s += red[w];
        a.mse_part[blockIdx.x] = s;
    }
    PR_MARK(5);
    if constexpr (FUSE) {
        if (pr_fused_finalize<T, NT>(a, it, reinterpret_cast<double*>(U)) || it + 1 >= a.niter) break;
    } else {
        break;
    }
    }
    PR_MARK(9);
}

// ---------------------------------------------------------------------------------------
// reductions / update
// ---------------------------------------------------------------------------------------
// sums[p] = sum over parts of partial[part][p] (double), sums[KK] = (sum of mse parts, 0):
// the buffer a z-sharded run all-reduces across ranks.
template <typename T>
__global__ void pr_reduce_kernel(const cplx<T>* __restrict__ partial, const double* __restrict__ mse_part, int nparts,
                                 int nmse, int KK, double2* __restrict__ sums, const PrState* __restrict__ state) {
    if (state->stopped) return;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < KK) {
        double sx = 0.0, sy = 0.0;
        for (int q = 0; q < nparts; ++q) {
            cplx<T> v = partial[(size_t)q * KK + p];
            sx += (double)v.x; sy += (double)v.y;
        }
        sums[p] = make_double2(sx, sy);
    }
    if (p == KK) {
        double s = 0.0;
        for (int q = 0; q < nmse; ++q) s += mse_part[q];
        sums[KK] = make_double2(s, 0.0);
    }
}

// Window of one rank (peer-mapped by every other rank), z-sharded runs:
//   [2 slots][world][KK + nblk] records of 2 x 16 bytes; entry [slot][r] is WRITTEN BY RANK r (posted stores over NVLink):
//   its partial pupil sums (one record per pixel: re, im), then one copy of its squared-error sum per finalize CTA.
// Every 16-byte word carries one double as {lo32, flag, hi32, flag} with flag = the iteration's epoch, so each 8-byte
// half validates itself: the reader spins on the word until both flags match -- no fence, no separate ready flag,
// no round trip (the low-latency protocol of collective libraries).  A rank publishes into every window, its own
// included; every rank then polls and reads only its own local window.
__host__ __device__ __forceinline__ size_t pr_win_bytes(int KK, int nblk, int world) { return (size_t)32 * 2 * world * (KK + nblk); }
__host__ __device__ __forceinline__ uint4* pr_win_rec(void* win, int KK, int nblk, int world, int slot, int r) {
    return reinterpret_cast<uint4*>(win) + 2 * ((size_t)slot * world + r) * (KK + nblk);      // two words per record
}
__device__ __forceinline__ void ll_store(uint4* dst, double v, unsigned flag) {
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(dst), "r"(lo), "r"(flag), "r"(hi), "r"(flag) : "memory");
}
// spins until the word carries `flag` (bounded: false after ~20 s, a peer died)
__device__ __forceinline__ bool ll_load(const uint4* src, unsigned flag, double& v) {
    unsigned long long t0 = 0;
    for (unsigned spin = 0;; ++spin) {
        unsigned a, b, c, d;
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(src) : "memory");
        if (b == flag && d == flag) { v = __hiloint2double((int)c, (int)a); return true; }
        if ((spin & 1023u) == 1023u) {
            unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t0 == 0) t0 = t1;
            else if (t1 - t0 > 20000000000ull) return false;
        }
    }
}

struct PrPeers {
    int world;                          // 0: no peers (single GPU or NCCL-reduced sums)
    int rank;
    int slot;
    long long epoch;                    // this iteration's flag value (its low 32 bits travel with every word)
    void* win[PYOTF_MAX_PEERS];         // every rank's window as mapped into this process (fixed rank order)
};

// deterministic warp sum of v[0], v[stride], ... (n terms): same order in every CTA.  The loads of a
// chunk of 16 x 32 terms are all issued before the first add, so a sum costs one L2 round trip.
__device__ __forceinline__ double warp_sum_fixed(const double* __restrict__ v, int n, int stride, int lane) {
    double s = 0.0;
    for (int base = 0; base < n; base += 16 * 32) {
        double t[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int i = base + lane + 32 * k;
            t[k] = i < n ? v[(size_t)i * stride] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) s += t[k];
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

constexpr int PR_FIN_PIX = 32;      // pixels per finalize CTA
constexpr int PR_FIN_LANES = 8;     // threads sharing one pixel's sum over the partials

// pr_finalize_kernel inside the plane kernel (single GPU, every CTA of the grid resident -- the grid has at most one
// CTA per SM): after a grid-wide arrival barrier CTA b reduces the partial pupils of ITS share of the pixels, in the
// same summation order as pr_finalize_kernel (8 lanes per pixel, lane l adds the parts l, l + 8, ... in order, then the
// lanes are added in order), so both paths produce bit-identical pupils.  The scalars (mse, mse_diff, pupil_diff, the
// stop decision) are evaluated redundantly and identically by every CTA.  One kernel per iteration.
template <typename T, int NT> __device__ int pr_fused_finalize(const PrPlaneArgs<T>& a, int it, double* smem_d) {
    const int tid = threadIdx.x, nblk = gridDim.x;
    const PrFinalArgs& f = a.fin;
    const int iter = f.iter + it, cur = (f.cur + it) & 1;             // this iteration of the launch
    const bool last_iter = f.last_iter && it + 1 == a.niter;
    __shared__ int s_stop;
    // A thread owns VL of a pixel's PR_FIN_LANES part-lanes (lane and lane + TL) and keeps one sum per part-lane, so the
    // summation order is pr_finalize_kernel's while NT / TL pixels -- this CTA's whole share on the fixture -- are
    // reduced in ONE pass: every load of the tail shares a single L2 round trip.
    constexpr int VL = 2, TL = PR_FIN_LANES / VL, PIX = NT / TL, NPB = 16;      // NPB: parts per part-lane and batch
    const int KK = f.K * f.K;
    const int per = (KK + nblk - 1) / nblk, p0 = blockIdx.x * per, p1 = min(KK, p0 + per);
    const int px = tid % PIX, lane = tid / PIX;
    const cplx<T>* __restrict__ oldp = cur ? a.pupil1 : a.pupil0;
    cplx<T>* __restrict__ newp = cur ? a.pupil0 : a.pupil1;
    double2* red2 = reinterpret_cast<double2*>(smem_d);                 // [PR_FIN_LANES][PIX]
    double* redn = smem_d + 2 * PR_FIN_LANES * PIX;                     // [NT / 32][2]
    double* scal = redn + 2 * (NT / 32);                                // squared-error sum, pupil-difference numerator / denominator
    // ---- what does not depend on the other CTAs is loaded ahead of the barrier: last iteration's mse (the thread that
    //      takes the stop decision), the mask and the old pupil of this thread's first pixel -------------------------
    const bool decider = tid == NT - 1;                                 // owns no pixel unless per >= PIX
    const double old_mse = (decider && iter > 0) ? a.mse[iter - 1] : 0.0;      // written by the previous iteration
    bool live = p0 + px < p1 && a.mask[p0 + px];                        // pixels outside the pupil are zero whatever the sum
    const cplx<T> old0 = (lane == 0 && p0 + px < p1) ? oldp[p0 + px] : mk<T>(0, 0);
    // ---- grid barrier: this CTA's partial pupil and squared-error sum are visible before it arrives -------------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        atomicAdd(a.arrive, 1u);
        const unsigned target = (unsigned)(2 * iter + 1) * (unsigned)nblk;    // two barriers per iteration
        unsigned seen;
        unsigned long long t0 = 0;
        s_stop = 0;
        for (unsigned spin = 0;; ++spin) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.arrive) : "memory");
            if ((int)(seen - target) >= 0) break;
            if ((spin & 1023u) == 1023u) {                      // bounded: a CTA of this grid never became resident
                unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t0 == 0) t0 = t1;
                else if (t1 - t0 > 2000000000ull) { s_stop = 2; a.state_rw->stopped = 3; pr_publish_progress(f.host_prog, iter, 3); break; }
            }
        }
    }
    __syncthreads();
    PR_MARK(6);
    if (s_stop == 2) return 2;
    // ---- every load is issued before anything waits: the partial pupils of this CTA's pixels (all threads) and the
    //      three scalar sums (warps 0 .. 2, one each) --------------------------------------------------------------
    if (tid < 96) {
        const int w = tid >> 5, l = tid & 31;
        const double* src = w == 0 ? a.mse_part : a.diffpart + (size_t)(iter & 1) * 2 * nblk + (w - 1);
        const int n = w == 0 ? f.nmse : nblk, stride = w == 0 ? 1 : 2;
        double sv = 0.0;
        if (w == 0 || iter > 0) sv = warp_sum_fixed(src, n, stride, l);
        if (l == 0) scal[w] = sv;
    }
    cplx<T> v[VL][NPB];
    auto load_batch = [&](int p, bool lv) {
        const cplx<T>* __restrict__ pp = a.partial + p;
#pragma unroll
        for (int j = 0; j < VL; ++j)
#pragma unroll
            for (int k = 0; k < NPB; ++k) {
                const int q = lane + j * TL + k * PR_FIN_LANES;
                v[j][k] = (lv && q < f.nparts) ? pp[(size_t)q * KK] : mk<T>(0, 0);
            }
    };
    load_batch(p0 + px, live);
    __syncthreads();
    if (decider) {                                                      // the other threads go on to their sums
        const double new_mse = scal[0] / f.npix_total;
        double md = nan(""), pd = nan("");
        if (iter > 0) {
            md = fabs(old_mse - new_mse) / old_mse;                    // phaseretrieval.py:103
            pd = scal[1] / scal[2];                                    // :105
        }
        const int stop = (pd < f.pupil_tol) || (md < f.mse_tol) || (new_mse < f.mse_tol);   // :114
        s_stop = stop;
        if (blockIdx.x == 0) {
            a.mse[iter] = new_mse; a.mse_diff[iter] = md; a.pupil_diff[iter] = pd;
            a.state_rw->iters_run = iter + 1;
            if (stop) { a.state_rw->final_is_next = 0; a.state_rw->stopped = 1; }
            else if (last_iter) a.state_rw->final_is_next = 1;
            pr_publish_progress(f.host_prog, iter + 1, stop);
        }
    }
    double num = 0.0, den = 0.0;
    for (int base = p0; base < p1; base += PIX) {
        const int p = base + px;
        double sx[VL], sy[VL];
#pragma unroll
        for (int j = 0; j < VL; ++j) {
            sx[j] = 0.0; sy[j] = 0.0;
#pragma unroll
            for (int k = 0; k < NPB; ++k) { sx[j] += (double)v[j][k].x; sy[j] += (double)v[j][k].y; }
        }
        if (live && f.nparts > NPB * PR_FIN_LANES) {                    // more than 128 parts
            const cplx<T>* __restrict__ pp = a.partial + p;
#pragma unroll
            for (int j = 0; j < VL; ++j)
                for (int q0 = lane + j * TL + NPB * PR_FIN_LANES; q0 < f.nparts; q0 += NPB * PR_FIN_LANES) {
                    cplx<T> u[NPB];
#pragma unroll
                    for (int k = 0; k < NPB; ++k) {
                        const int q = q0 + k * PR_FIN_LANES;
                        u[k] = q < f.nparts ? pp[(size_t)q * KK] : mk<T>(0, 0);
                    }
#pragma unroll
                    for (int k = 0; k < NPB; ++k) { sx[j] += (double)u[k].x; sy[j] += (double)u[k].y; }
                }
        }
        // a further pass (fewer CTAs than on the fixture): its first batch is in flight while this one is reduced
        bool live_n = false;
        if (base + PIX < p1) {
            live_n = p + PIX < p1 && a.mask[p + PIX];
            load_batch(p + PIX, live_n);
        }
        if (base != p0) __syncthreads();
#pragma unroll
        for (int j = 0; j < VL; ++j) red2[(lane + j * TL) * PIX + px] = make_double2(sx[j], sy[j]);
        __syncthreads();
        if (base == p0) {
            PR_MARK(7);
            if (s_stop) return 1;                                       // written before the barrier above
        }
        if (lane == 0 && p < p1) {
            double tx = 0.0, ty = 0.0;
#pragma unroll
            for (int l = 0; l < PR_FIN_LANES; ++l) { tx += red2[l * PIX + px].x; ty += red2[l * PIX + px].y; }
            const double m = live ? 1.0 : 0.0;
            tx = tx / f.nz_total * m; ty = ty / f.nz_total * m;        // mean(0) * mask   :127
            if (f.phase_only) {                                        // exp(i angle(.)) * mask   :130
                const double mag = hypot(tx, ty);
                if (mag > 0.0) { tx = tx / mag * m; ty = ty / mag * m; }
                else { tx = m; ty = 0.0; }
            }
            const cplx<T> o = base == p0 ? old0 : oldp[p];
            const cplx<T> nv = mk<T>((T)tx, (T)ty);
            newp[p] = nv;
            const double dx = (double)o.x - (double)nv.x, dy = (double)o.y - (double)nv.y;
            num += dx * dx + dy * dy;
            den += (double)o.x * (double)o.x + (double)o.y * (double)o.y;
        }
        live = live_n;
    }
    PR_MARK(8);
    // deterministic block sums of num / den (fixed shuffle tree, warps added in order)
    for (int o = 16; o; o >>= 1) { num += __shfl_xor_sync(0xffffffffu, num, o); den += __shfl_xor_sync(0xffffffffu, den, o); }
    __syncthreads();
    if ((tid & 31) == 0) { redn[2 * (tid >> 5)] = num; redn[2 * (tid >> 5) + 1] = den; }
    __syncthreads();
    if (tid == 0) {
        double n2 = 0.0, d2 = 0.0;
        for (int w = 0; w < NT / 32; ++w) { n2 += redn[2 * w]; d2 += redn[2 * w + 1]; }
        double* dp = a.diffpart + (size_t)((iter + 1) & 1) * 2 * nblk;
        dp[2 * blockIdx.x] = n2; dp[2 * blockIdx.x + 1] = d2;
    }
    // ---- second grid barrier: the new pupil, this CTA's share of the pupil-difference sums and mse[iter] are visible
    //      to every CTA before the next iteration of this launch reads them.  The last iteration of a launch only
    //      arrives (the counter stays two per iteration): the kernel boundary orders the rest. ------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        atomicAdd(a.arrive, 1u);
        if (it + 1 < a.niter) {
            const unsigned target = (unsigned)(2 * iter + 2) * (unsigned)nblk;
            unsigned seen;
            unsigned long long t0 = 0;
            for (unsigned spin = 0;; ++spin) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.arrive) : "memory");
                if ((int)(seen - target) >= 0) break;
                if ((spin & 1023u) == 1023u) {
                    unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t0 == 0) t0 = t1;
                    else if (t1 - t0 > 2000000000ull) { s_stop = 2; a.state_rw->stopped = 3; pr_publish_progress(f.host_prog, iter, 3); break; }
                }
            }
        }
    }
    __syncthreads();
    return s_stop == 2 ? 2 : 0;
}


// TP = cplx<T> (per-CTA partials of this rank) or double2 (all-reduced sums: nparts = 1 and
// the squared-error numerator sits at index KK).  Block = 256 threads = 32 pixels x 8 part lanes.
template <typename T, typename TP>
__global__ void __launch_bounds__(256) pr_finalize_kernel(PrFinalArgs a, const TP* __restrict__ partial,
                                                          const double* __restrict__ mse_part,
                                                          const unsigned char* __restrict__ mask,
                                                          cplx<T>* __restrict__ pupil0, cplx<T>* __restrict__ pupil1,
                                                          double* __restrict__ diffpart /* [2][nblk][2] */,
                                                          double* __restrict__ mse, double* __restrict__ mse_diff,
                                                          double* __restrict__ pupil_diff, PrState* __restrict__ state) {
    __shared__ double2 red[PR_FIN_LANES][PR_FIN_PIX];
    __shared__ int s_stop;
    const int tid = threadIdx.x;
    const int px = tid & (PR_FIN_PIX - 1), lane = tid / PR_FIN_PIX;
    const int KK = a.K * a.K;
    pdl_launch_dependents();
    pdl_wait();
    if (state->stopped) return;
    // every CTA evaluates the scalars in the same fixed order, so all of them take the same branch
    if (tid < 32) {
        double s;
        if (a.nparts == 0) s = reinterpret_cast<const double2*>(partial)[KK].x;
        else s = warp_sum_fixed(mse_part, a.nmse, 1, tid);
        const double new_mse = s / a.npix_total;
        double md = nan(""), pd = nan("");
        if (a.iter > 0) {
            const double old_mse = mse[a.iter - 1];
            md = fabs(old_mse - new_mse) / old_mse;                    // phaseretrieval.py:103
            const double* dp = diffpart + (size_t)(a.iter & 1) * 2 * gridDim.x;
            const double num = warp_sum_fixed(dp, gridDim.x, 2, tid);
            const double den = warp_sum_fixed(dp + 1, gridDim.x, 2, tid);
            pd = num / den;                                            // :105
        }
        const int stop = (pd < a.pupil_tol) || (md < a.mse_tol) || (new_mse < a.mse_tol);   // :114
        if (tid == 0) {
            s_stop = stop;
            if (blockIdx.x == 0) {
                mse[a.iter] = new_mse; mse_diff[a.iter] = md; pupil_diff[a.iter] = pd;
                state->iters_run = a.iter + 1;
                if (stop) { state->final_is_next = 0; state->stopped = 1; }
                else if (a.last_iter) state->final_is_next = 1;
                pr_publish_progress(a.host_prog, a.iter + 1, stop);
            }
        }
    }
    __syncthreads();
    if (s_stop) return;
    const cplx<T>* __restrict__ oldp = a.cur ? pupil1 : pupil0;
    cplx<T>* __restrict__ newp = a.cur ? pupil0 : pupil1;
    const int p = blockIdx.x * PR_FIN_PIX + px;
    double sx = 0.0, sy = 0.0;
    if (p < KK) {
        if (a.nparts == 0) {
            if (lane == 0) { const double2 v = reinterpret_cast<const double2*>(partial)[p]; sx = v.x; sy = v.y; }
        } else {
            const cplx<T>* __restrict__ pp = reinterpret_cast<const cplx<T>*>(partial) + p;
            for (int q0 = lane; q0 < a.nparts; q0 += 16 * PR_FIN_LANES) {    // 16 loads in flight per thread
                cplx<T> v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int q = q0 + k * PR_FIN_LANES;
                    v[k] = q < a.nparts ? pp[(size_t)q * KK] : mk<T>(0, 0);
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) { sx += (double)v[k].x; sy += (double)v[k].y; }
            }
        }
    }
    red[lane][px] = make_double2(sx, sy);
    __syncthreads();
    if (tid < 32) {
        double num = 0.0, den = 0.0;
        if (p < KK) {
            sx = 0.0; sy = 0.0;
#pragma unroll
            for (int l = 0; l < PR_FIN_LANES; ++l) { sx += red[l][px].x; sy += red[l][px].y; }
            const double m = mask[p] ? 1.0 : 0.0;
            sx = sx / a.nz_total * m; sy = sy / a.nz_total * m;        // mean(0) * mask   :127
            if (a.phase_only) {                                        // exp(i angle(.)) * mask   :130
                const double mag = hypot(sx, sy);
                if (mag > 0.0) { sx = sx / mag * m; sy = sy / mag * m; }
                else { sx = m; sy = 0.0; }
            }
            const cplx<T> o = oldp[p];
            const cplx<T> nv = mk<T>((T)sx, (T)sy);
            newp[p] = nv;
            const double dx = (double)o.x - (double)nv.x, dy = (double)o.y - (double)nv.y;
            num = dx * dx + dy * dy;
            den = (double)o.x * (double)o.x + (double)o.y * (double)o.y;
        }
        for (int o = 16; o; o >>= 1) {
            num += __shfl_xor_sync(0xffffffffu, num, o);
            den += __shfl_xor_sync(0xffffffffu, den, o);
        }
        if (tid == 0) {
            double* dp = diffpart + (size_t)((a.iter + 1) & 1) * 2 * gridDim.x;
            dp[2 * blockIdx.x] = num; dp[2 * blockIdx.x + 1] = den;
        }
    }
}


// pr_finalize_kernel for z-sharded runs, fused with the sum over ranks: CTA b sums this rank's partial pupils for
// its 32 pixels and stores the result (and this rank's squared-error sum) into entry [rank] of EVERY rank's window --
// posted writes over NVLink, nobody waits for a round trip.  It then reads the entries of CTA b of every rank from
// its OWN window (spinning on the self-validating words, local memory) and adds them in rank order, so all
// ranks compute bit-identical pupils and scalars and take identical stop decisions: the all-reduce of the
// reference-less multi-GPU loop without a collective call or an extra launch.
template <typename T>
__global__ void __launch_bounds__(256) pr_finalize_peers_kernel(PrFinalArgs a, PrPeers peers, const cplx<T>* __restrict__ partial,
                                                                const double* __restrict__ mse_part,
                                                                const unsigned char* __restrict__ mask,
                                                                cplx<T>* __restrict__ pupil0, cplx<T>* __restrict__ pupil1,
                                                                double* __restrict__ diffpart /* [2][nblk][2] */,
                                                                double* __restrict__ mse, double* __restrict__ mse_diff,
                                                                double* __restrict__ pupil_diff, PrState* __restrict__ state) {
    __shared__ double2 red[PR_FIN_LANES][PR_FIN_PIX];
    const int tid = threadIdx.x;
    const int px = tid & (PR_FIN_PIX - 1), lane = tid / PR_FIN_PIX;
    const int KK = a.K * a.K, nblk = gridDim.x, W = peers.world;
    pdl_launch_dependents();
    pdl_wait();
    if (state->stopped) return;
    const int p = blockIdx.x * PR_FIN_PIX + px;
    // ---- this rank's partial sums for the CTA's pixels (16 loads in flight per thread) ----
    double sx = 0.0, sy = 0.0;
    if (p < KK) {
        const cplx<T>* __restrict__ pp = partial + p;
        for (int q0 = lane; q0 < a.nparts; q0 += 16 * PR_FIN_LANES) {
            cplx<T> v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int q = q0 + k * PR_FIN_LANES;
                v[k] = q < a.nparts ? pp[(size_t)q * KK] : mk<T>(0, 0);
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) { sx += (double)v[k].x; sy += (double)v[k].y; }
        }
    }
    red[lane][px] = make_double2(sx, sy);
    __syncthreads();
    if (tid < 32) {
        const double mse_local = warp_sum_fixed(mse_part, a.nmse, 1, tid);
        sx = 0.0; sy = 0.0;
#pragma unroll
        for (int l = 0; l < PR_FIN_LANES; ++l) { sx += red[l][px].x; sy += red[l][px].y; }
        // ---- publish into every rank's window (self-validating words: no fence, no flag) ----
        const unsigned flag = (unsigned)peers.epoch;
        for (int r = 0; r < W; ++r) {
            uint4* dst = pr_win_rec(peers.win[r], KK, nblk, W, peers.slot, peers.rank);
            if (p < KK) { ll_store(dst + 2 * p, sx, flag); ll_store(dst + 2 * p + 1, sy, flag); }
            if (tid == 0) ll_store(dst + 2 * (KK + blockIdx.x), mse_local, flag);
        }
        // ---- sums over ranks, fixed order; every rank reads the same stored doubles from its own window ----
        void* const mine = peers.win[peers.rank];
        double msum = 0.0;
        bool ok = true;
        sx = 0.0; sy = 0.0;
        for (int r = 0; r < W; ++r) {
            const uint4* src = pr_win_rec(mine, KK, nblk, W, peers.slot, r);
            double m, vx = 0.0, vy = 0.0;
            ok &= ll_load(src + 2 * (KK + blockIdx.x), flag, m);
            if (p < KK) { ok &= ll_load(src + 2 * p, flag, vx); ok &= ll_load(src + 2 * p + 1, flag, vy); }
            msum += m; sx += vx; sy += vy;
        }
        if (!__all_sync(0xffffffffu, ok)) { if (tid == 0) { state->stopped = 2; pr_publish_progress(a.host_prog, a.iter, 2); } return; }
        // ---- scalars and the stop test: identical on every CTA of every rank ----
        const double new_mse = msum / a.npix_total;
        double md = nan(""), pd = nan("");
        if (a.iter > 0) {
            const double old_mse = mse[a.iter - 1];
            md = fabs(old_mse - new_mse) / old_mse;                    // phaseretrieval.py:103
            const double* dp = diffpart + (size_t)(a.iter & 1) * 2 * gridDim.x;
            const double num = warp_sum_fixed(dp, gridDim.x, 2, tid);
            const double den = warp_sum_fixed(dp + 1, gridDim.x, 2, tid);
            pd = num / den;                                            // :105
        }
        const int stop = (pd < a.pupil_tol) || (md < a.mse_tol) || (new_mse < a.mse_tol);   // :114
        if (tid == 0 && blockIdx.x == 0 && *reinterpret_cast<volatile int*>(&state->stopped) == 0) {
            mse[a.iter] = new_mse; mse_diff[a.iter] = md; pupil_diff[a.iter] = pd;
            state->iters_run = a.iter + 1;
            if (stop) { state->final_is_next = 0; state->stopped = 1; }
            else if (a.last_iter) state->final_is_next = 1;
            pr_publish_progress(a.host_prog, a.iter + 1, stop);
        }
        if (!stop) {
            const cplx<T>* __restrict__ oldp = a.cur ? pupil1 : pupil0;
            cplx<T>* __restrict__ newp = a.cur ? pupil0 : pupil1;
            double num = 0.0, den = 0.0;
            if (p < KK) {
                const double m = mask[p] ? 1.0 : 0.0;
                sx = sx / a.nz_total * m; sy = sy / a.nz_total * m;        // mean(0) * mask   :127
                if (a.phase_only) {                                        // exp(i angle(.)) * mask   :130
                    const double mag = hypot(sx, sy);
                    if (mag > 0.0) { sx = sx / mag * m; sy = sy / mag * m; }
                    else { sx = m; sy = 0.0; }
                }
                const cplx<T> o = oldp[p];
                const cplx<T> nv = mk<T>((T)sx, (T)sy);
                newp[p] = nv;
                const double dx = (double)o.x - (double)nv.x, dy = (double)o.y - (double)nv.y;
                num = dx * dx + dy * dy;
                den = (double)o.x * (double)o.x + (double)o.y * (double)o.y;
            }
            for (int o = 16; o; o >>= 1) {
                num += __shfl_xor_sync(0xffffffffu, num, o);
                den += __shfl_xor_sync(0xffffffffu, den, o);
            }
            if (tid == 0) {
                double* dp = diffpart + (size_t)((a.iter + 1) & 1) * 2 * gridDim.x;
                dp[2 * blockIdx.x] = num; dp[2 * blockIdx.x + 1] = den;
            }
        }
    }
}

// initial pupil = ideal mask as complex ones (HanserPSF._gen_pupil, otf.py:346-355), both buffers' padding zeroed
template <typename T>
__global__ void pr_init_pupil_kernel(const unsigned char* __restrict__ mask, int KKP, cplx<T>* __restrict__ p0,
                                     cplx<T>* __restrict__ p1) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= KKP) return;
    p0[i] = mk<T>(mask[i] ? (T)1 : (T)0, 0);
    p1[i] = mk<T>(0, 0);
}

// measured stack (float32 or float64, any of the two) -> T
template <typename TS, typename T>
__global__ void pr_convert_data_kernel(const TS* __restrict__ src, T* __restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (T)src[i];
}

// compact [KR][K] pupil (T) -> unshifted N x N complex128 (zero outside the box)
template <typename T>
__global__ void pr_unpack_pupil_kernel(int N, int K, int f0, const cplx<T>* __restrict__ src, double2* __restrict__ dst) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * N) return;
    int y = idx / N, x = idx - y * N;
    int fy = y < (N + 1) / 2 ? y : y - N, fx = x < (N + 1) / 2 ? x : x - N;   // fftfreq order, any N
    int iy = fy - f0, ix = fx - f0;
    double2 v = make_double2(0.0, 0.0);
    if (iy >= 0 && iy < K && ix >= 0 && ix < K) { cplx<T> s = src[iy * K + ix]; v = make_double2((double)s.x, (double)s.y); }
    dst[idx] = v;
}

}  // namespace pyotf
