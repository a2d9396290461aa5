// K3 fused path: SheppardPSF.PSFa / .PSFi without materialising the amplitude OTF (pyotf/otf.py:535-592,
// :204-207).  easy_ifft(OTFa, axes=(1,2,3)) is evaluated as three line passes:
//   pass 1 (x, contiguous): the shell values of a (kz, ky) row are generated on the fly inside the load
//           (same fp64 predicates as sheppard_otf_kernel); rows that cannot touch the shell are zero-filled
//           without any transform (the shell occupies 17 % of the rows of 34 % of the planes at config 3)
//   pass 2 (y): the generic axis pass, skipping planes that cannot hold shell voxels
//   pass 3 (z): transform + 1/n scaling + |.|^2 accumulated over the polarisation components into PSFi
//           (and / or the PSFa store), so neither OTFa nor a separate |.|^2 pass touches HBM.
#pragma once
#include <cmath>
#include <vector>

#include "common.cuh"
#include "hanser_impl.cuh"
#include "sheppard.cuh"

namespace pyotf {

template <typename T> struct ShepArgs {
    SheppardParams p;
    int v;                      // polarisation component handled by this launch
    int in_x, out_x;            // shifts of pass 1 (see AxisPass)
    int in_z, out_z;            // shifts of pass 3
    cplx<T>* tmp;               // [NZ][N][N] of component v
    cplx<T>* psfa;              // [NZ][N][N] of component v or nullptr
    T* psfi;                    // [NZ][N][N] or nullptr
    const cplx<T>* tw_x;        // [N]
    const cplx<T>* tw_z;        // [NZ]
    const unsigned char* planeflag;   // [NZ][N]: 1 where the (memory-layout) plane zm can hold shell voxels
    int band0[2], bandn[2];           // the same planes as up to two runs [band0, band0 + bandn) of zm (bandn[0] < 0: use planeflag)
    int accumulate;
    double scale;
};

// value of polarisation component v on a shell voxel (otf.py:542-575); scalar model: 1
template <typename T>
__device__ __forceinline__ T sheppard_value(const SheppardParams& p, int v, double kz, double ky, double kx) {
    if (p.vec_mode == 0) return (T)1;
    const double kn = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(kx, kx), __dmul_rn(ky, ky)), __dmul_rn(kz, kz)));
    const double m = kx / kn, n = ky / kn, s = kz / kn;
    double a = 1.0;
    if (p.condition == 1) a = 1.0 / sqrt(s);
    else if (p.condition == 2) a = 1.0 / s;
    // component order: z -> (Pzx, Pzy), y -> (Pyx, Pyy), x -> (Pxx, Pxy); "total" = all six in that order
    int c = v;
    if (p.vec_mode == 2) c = v + 2;
    else if (p.vec_mode == 1) c = v + 4;
    double val;
    switch (c) {
        case 0: val = -m; break;
        case 1: val = -n; break;
        case 2: val = -n * m / (1 + s); break;
        case 3: val = 1 - n * n / (1 + s); break;
        case 4: val = 1 - m * m / (1 + s); break;
        default: val = -m * n / (1 + s); break;
    }
    return (T)(val * a);
}

// value of the amplitude OTF at frequency (kz, ky, kx), r2zy = kz^2 + ky^2: the fp64 predicates of otf.py:503-532 in the
// reference's operation order (the same as sheppard_otf_kernel), 0 off the shell
template <typename T>
__device__ __forceinline__ T sheppard_shell_value(const SheppardParams& p, int v, double kz, double ky, double kx, double r2zy) {
    const double kr2 = __dadd_rn(r2zy, __dmul_rn(kx, kx));         // ((kz^2 + ky^2) + kx^2): numpy's order
    if (kr2 < p.r2_lo || kr2 > p.r2_hi) return (T)0;
    const double kr = sqrt(kr2);
    double dkr;
    if (p.same_dk) dkr = p.dk;
    else if (kr2 == 0.0) dkr = 0.0;
    else {
        const double aa = __dmul_rn(kz, p.dkz), bb = __dmul_rn(ky, p.dk), cc = __dmul_rn(kx, p.dk);
        dkr = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(aa, aa), __dmul_rn(bb, bb)), __dmul_rn(cc, cc))) / kr;
    }
    const double kzz = p.dual ? fabs(kz) : kz;
    const bool valid = (fabs(__dsub_rn(kr, p.kmag)) < dkr) && (kzz > __dadd_rn(p.kz_min, dkr));
    if (!valid) return (T)0;
    return sheppard_value<T>(p, v, kz, ky, kx);
}

// planeflag[zm][x] = 1 iff the plane with memory index zm (fftshift-ed layout) can hold shell voxels:
// kzz > kz_min (valid needs kzz > kz_min + dkr, dkr >= 0) and |kz| <= kmag + max(dk, dkz) (since |kz| <= kr)
static __global__ void sheppard_planeflag_kernel(SheppardParams p, unsigned char* __restrict__ flag) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)p.NZ * p.N) return;
    const int zm = (int)(i / p.N);
    int zu = zm - p.NZ / 2; if (zu < 0) zu += p.NZ;
    const double kz = fftfreq_at(zu, p.NZ, p.kzstep);
    const double kzz = p.dual ? fabs(kz) : kz;
    const double dmax = fmax(fabs(p.dk), fabs(p.dkz));
    flag[i] = (kzz > p.kz_min && fabs(kz) <= (p.kmag + dmax) * (1.0 + 1e-9)) ? 1 : 0;
}

// ---- pass 1: line = (zm, ym) in the fftshift-ed memory layout, axis = x ---------------------------------
template <typename T> struct ShepRow : NoAcc {
    static constexpr bool HAS_EMPTY = true;
    struct Line { double kz, ky, r2zy; cplx<T>* out; int maybe, plane_maybe; };
    ShepArgs<T> a;
    __device__ __forceinline__ cplx<T> twiddle(int k) const { return a.tw_x[k]; }
    __device__ __forceinline__ Line info(long long line) const {
        const SheppardParams& p = a.p;
        const int ym = (int)(line % p.N), zm = (int)(line / p.N);
        int yu = ym - p.N / 2; if (yu < 0) yu += p.N;               // OTFa[j] = otf[(j - n//2) mod n]
        int zu = zm - p.NZ / 2; if (zu < 0) zu += p.NZ;
        Line d;
        d.kz = fftfreq_at(zu, p.NZ, p.kzstep);
        d.ky = fftfreq_at(yu, p.N, p.kstep);
        d.r2zy = __dadd_rn(__dmul_rn(d.kz, d.kz), __dmul_rn(d.ky, d.ky));
        d.out = a.tmp + (size_t)line * p.N;
        // can any voxel of this row lie on the shell?  (conservative: the exact predicate runs per voxel)
        const double kxmax = p.kstep * (double)(p.N / 2);
        const double kzz = p.dual ? fabs(d.kz) : d.kz;
        d.plane_maybe = a.planeflag[(size_t)zm * p.N];
        d.maybe = d.plane_maybe && d.r2zy <= p.r2_hi && d.r2zy + kxmax * kxmax >= p.r2_lo && kzz > p.kz_min;
        return d;
    }
    __device__ __forceinline__ bool empty(const Line& d) const { return !d.maybe; }
    // rows of planes that cannot hold the shell are neither transformed nor zero-filled: passes 2 and 3 skip them
    __device__ __forceinline__ void zero_fill(const Line& d, int o) const { if (d.plane_maybe) d.out[o] = mk<T>(0, 0); }
    __device__ __forceinline__ cplx<T> load(const Line& d, int l) const {
        const SheppardParams& p = a.p;
        int xm = l + a.in_x; if (xm >= p.N) xm -= p.N;             // x_in[l] = OTFa[(l + in_shift) mod n]
        int xu = xm - p.N / 2; if (xu < 0) xu += p.N;
        const double kx = fftfreq_at(xu, p.N, p.kstep);
        return mk<T>(sheppard_shell_value<T>(p, a.v, d.kz, d.ky, kx, d.r2zy), 0);
    }
    __device__ __forceinline__ void store(const Line& d, int o, cplx<T> v, int&) const {
        int xm = o + a.out_x; if (xm >= a.p.N) xm -= a.p.N;
        d.out[xm] = cscale(v, (T)a.scale);
    }
};

// ---- pass 3: line = (ym, xm), axis = z --------------------------------------------------------------------
// FAST: the occupied planes are given as index runs and the only output is a fresh PSFi (the scalar model's usual call)
template <typename T, bool FAST = false> struct ShepCol : NoAcc {
    struct Line { const cplx<T>* in; cplx<T>* psfa; T* psfi; };
    ShepArgs<T> a;
    __device__ __forceinline__ cplx<T> twiddle(int k) const { return a.tw_z[k]; }
    __device__ __forceinline__ Line info(long long line) const {
        Line d;
        d.in = a.tmp + line;
        d.psfa = a.psfa ? a.psfa + line : nullptr;
        d.psfi = a.psfi ? a.psfi + line : nullptr;
        return d;
    }
    // (N * N * NZ < 2^31 is checked by the launcher: plane offsets are 32-bit products)
    __device__ __forceinline__ cplx<T> load(const Line& d, int l) const {
        int zm = l + a.in_z; if (zm >= a.p.NZ) zm -= a.p.NZ;
        // planes that hold no shell voxel were never written
        if constexpr (FAST) {
            const bool on = (unsigned)(zm - a.band0[0]) < (unsigned)a.bandn[0] || (unsigned)(zm - a.band0[1]) < (unsigned)a.bandn[1];
            return on ? d.in[(unsigned)zm * (unsigned)(a.p.N * a.p.N)] : mk<T>(0, 0);
        }
        if (a.bandn[0] >= 0) {
            if ((unsigned)(zm - a.band0[0]) >= (unsigned)a.bandn[0] && (unsigned)(zm - a.band0[1]) >= (unsigned)a.bandn[1]) return mk<T>(0, 0);
        } else if (!a.planeflag[(size_t)zm * a.p.N]) {
            return mk<T>(0, 0);
        }
        return d.in[(unsigned)zm * (unsigned)(a.p.N * a.p.N)];
    }
    __device__ __forceinline__ void store(const Line& d, int o, cplx<T> v, int&) const {
        int zm = o + a.out_z; if (zm >= a.p.NZ) zm -= a.p.NZ;
        const unsigned off = (unsigned)zm * (unsigned)(a.p.N * a.p.N);
        v = cscale(v, (T)a.scale);
        if constexpr (FAST) { d.psfi[off] = v.x * v.x + v.y * v.y; return; }
        if (d.psfa) d.psfa[off] = v;
        if (d.psfi) {
            const T I = v.x * v.x + v.y * v.y;
            if (a.accumulate) d.psfi[off] += I; else d.psfi[off] = I;       // otf.py:204-207
        }
    }
};

// ---- z-first evaluation (scalar model, power-of-two sizes) -----------------------------------------------------
// ifftn(OTFa) = [ifft2 over (ky, kx)] o [ifft over kz].  Along kz a (ky, kx) column of the cap holds a handful of
// voxels (the cap is one or two voxels thick; more only near its rim), so the z transform is a direct sum over them;
// at fixed z the result is non-zero inside the disk kxy < kxy_max only; and the scalar cap is even in ky and in kx
// (its predicates see ky^2 and kx^2 only -- exactly, in floating point too).  That is precisely the input of the
// even-symmetric HanserPSF kernel K1: a per-plane quadrant table P_z(|fy|, |fx|) (for HanserPSF: amp * defocus).
// sheppard_zplanes_kernel writes the tables ([zm][dtab_size], 8.5 MB at config 3) and K1 (hanser_launch_tables)
// turns every plane into PSFa / PSFi on chip: the 134 MB volume is never formed, the only large traffic is the output.
//
// CTA = (8 columns ax0 .. ax0 + 7 of quadrant row ay, one per warp).  Phase A: the shell voxels of each column as a list (freq
// index, value), the predicates evaluated on the occupied kz planes only.  Phase B: lanes = 32 consecutive z, direct
// sum over the column's list into a shared tile [z][col].  Phase C: the tile goes out row by row.
template <typename T> struct ShepZArgs {
    SheppardParams p;
    int H, HP, ntab;            // support box |f| <= H; table pitch and size (dtab_pitch / dtab_size of f0 = -H)
    int band0[2], bandn[2];     // the occupied kz planes as runs of the fftshift-ed index zm
    cplx<T>* tab;               // [NZ][ntab]
    double scale;               // 1 / NZ
    cplx<T>* twn;               // [N] e^{+2 pi i k / N} followed by NZ zero doubles: what K1 needs besides the tables
                                // (hanser_tables_aux_bytes), written by CTA (0, 0) so that no extra launch prepares it
};

constexpr int SHEPZ_COLS = 8, SHEPZ_ZC = 256, SHEPZ_TP = SHEPZ_COLS + 1;   // 8 columns per CTA: one per warp (the phases are latency-bound chains)

template <typename T>
__global__ void __launch_bounds__(256) sheppard_zplanes_kernel(ShepZArgs<T> a) {
    extern __shared__ __align__(16) unsigned char shepz_smem[];
    const SheppardParams& p = a.p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int CPW = SHEPZ_COLS / 8;                                     // columns per warp
    const int NZ = p.NZ, ZC = NZ < SHEPZ_ZC ? NZ : SHEPZ_ZC;
    const int nband = a.bandn[0] + a.bandn[1];
    const int ay = blockIdx.y, ax0 = blockIdx.x * SHEPZ_COLS;
    cplx<T>* tile = reinterpret_cast<cplx<T>*>(shepz_smem);                 // [ZC][SHEPZ_TP]
    cplx<T>* tw = tile + (size_t)ZC * SHEPZ_TP;                             // [NZ] e^{+2 pi i k / NZ}
    T* lval = reinterpret_cast<T*>(tw + NZ);                                // [SHEPZ_COLS][nband]
    int* lidx = reinterpret_cast<int*>(lval + (size_t)SHEPZ_COLS * nband);  // [SHEPZ_COLS][nband]
    __shared__ int nnz[SHEPZ_COLS];
    for (int k = tid; k < NZ; k += 256) {
        double sn, cs;
        sincospi(2.0 * (double)k / (double)NZ, &sn, &cs);
        tw[k] = mk<T>((T)cs, (T)sn);
    }
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        for (int k = tid; k < p.N; k += 256) {
            double sn, cs;
            sincospi(2.0 * (double)k / (double)p.N, &sn, &cs);
            a.twn[k] = mk<T>((T)cs, (T)sn);
        }
        double* zz = reinterpret_cast<double*>(a.twn + p.N);
        for (int k = tid; k < NZ; k += 256) zz[k] = 0.0;
    }
    // ---- phase A: column lists.  A lane evaluates up to NCH band planes of each of the warp's columns as independent
    //      chains (the fp64 predicates are latency-bound), then the survivors are compacted with ballots ----
    const double ky = __dmul_rn((double)ay, p.kstep);
    constexpr int NCH = 4;
    int count[CPW];
#pragma unroll
    for (int j = 0; j < CPW; ++j) count[j] = 0;
    for (int b0 = 0; b0 < nband; b0 += 32 * NCH) {
        T val[CPW][NCH];
        int zu[NCH];
#pragma unroll
        for (int q = 0; q < NCH; ++q) {
            const int b = b0 + 32 * q + lane;
            const bool on = b < nband;
            const int zm = b < a.bandn[0] ? a.band0[0] + b : a.band0[1] + (b - a.bandn[0]);
            zu[q] = zm - NZ / 2; if (zu[q] < 0) zu[q] += NZ;            // unshifted frequency index
            const double kz = fftfreq_at(on ? zu[q] : 0, NZ, p.kzstep);
            const double r2zy = __dadd_rn(__dmul_rn(kz, kz), __dmul_rn(ky, ky));
#pragma unroll
            for (int j = 0; j < CPW; ++j) {
                const int ax = ax0 + warp * CPW + j;
                const double kx = __dmul_rn((double)ax, p.kstep);
                val[j][q] = (on && ax <= a.H) ? sheppard_shell_value<T>(p, 0, kz, ky, kx, r2zy) : (T)0;   // (padding column: zero)
            }
        }
#pragma unroll
        for (int j = 0; j < CPW; ++j) {
            const int c = warp * CPW + j;
#pragma unroll
            for (int q = 0; q < NCH; ++q) {
                const unsigned m = __ballot_sync(0xffffffffu, val[j][q] != (T)0);
                if (val[j][q] != (T)0) {
                    const int pos = count[j] + __popc(m & ((1u << lane) - 1u));
                    lval[c * nband + pos] = val[j][q]; lidx[c * nband + pos] = zu[q];
                }
                count[j] += __popc(m);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CPW; ++j) if (lane == 0) nnz[warp * CPW + j] = count[j];
    __syncthreads();
    const T scale = (T)a.scale;
    for (int zb = 0; zb < NZ; zb += ZC) {
        // ---- phase B: direct z transform of the warp's columns; a lane carries ZU values of z (z, z + 32, ...) so that
        //      the list entries are read once and the table lookups are independent ----
        constexpr int ZU = 4;
        for (int c = warp * CPW; c < warp * CPW + CPW; ++c) {
            const int n = nnz[c];
            const T* __restrict__ lv = lval + c * nband;
            const int* __restrict__ li = lidx + c * nband;
            for (int zg = 0; zg < ZC; zg += 32 * ZU) {
                T sx[ZU], sy[ZU];
#pragma unroll
                for (int u = 0; u < ZU; ++u) { sx[u] = 0; sy[u] = 0; }
                for (int e = 0; e < n; ++e) {
                    const int k = li[e];
                    const T v = lv[e];
#pragma unroll
                    for (int u = 0; u < ZU; ++u) {
                        const cplx<T> t = tw[(k * (zb + zg + 32 * u + lane)) & (NZ - 1)];
                        sx[u] += v * t.x; sy[u] += v * t.y;
                    }
                }
#pragma unroll
                for (int u = 0; u < ZU; ++u)
                    if (zg + 32 * u < ZC) tile[(zg + 32 * u + lane) * SHEPZ_TP + c] = mk<T>(sx[u] * scale, sy[u] * scale);
            }
        }
        __syncthreads();
        // ---- phase C: rows of the tile -> table of plane zm = (z + NZ / 2) mod NZ (fftshift along z) ----
        {
            constexpr int RPW = 32 / SHEPZ_COLS;                                      // z rows per warp pass
            const int c = lane & (SHEPZ_COLS - 1);
            if (ax0 + c < a.HP) {
                for (int zz = warp * RPW + lane / SHEPZ_COLS; zz < ZC; zz += 8 * RPW) {
                    const int zm = (zb + zz + NZ / 2) & (NZ - 1);
                    a.tab[(size_t)zm * a.ntab + ay * a.HP + ax0 + c] = tile[zz * SHEPZ_TP + c];
                }
            }
        }
        __syncthreads();
    }
}

template <typename T> int hanser_launch_tables(int N, int f0, const void* tables, int nz, void* psfa, void* psfi, cudaStream_t s, const void* tw_z);
extern template int hanser_launch_tables<float>(int, int, const void*, int, void*, void*, cudaStream_t, const void*);
extern template int hanser_launch_tables<double>(int, int, const void*, int, void*, void*, cudaStream_t, const void*);

// returns 0 (done), 3 (configuration outside this path: the caller runs the three-pass evaluation) or an error
template <typename T>
int sheppard_zfirst(const SheppardParams& p, const int band0[2], const int bandn[2], void* psfa_out, void* psfi_out,
                    cudaStream_t s) {
    const bool off = getenv("PYOTF_B200_SHEPPARD_3PASS") != nullptr;   // A/B knob, read at every call (tests compare the two)
    const int N = p.N, NZ = p.NZ;
    if (off || p.vec_mode != 0 || p.nvec != 1 || !(N == 32 || N == 64 || N == 128 || N == 256) || !is_pow2(NZ) || NZ < 32 ||
        p.dk <= 0 || p.dkz <= 0 || bandn[0] < 0 || bandn[0] + bandn[1] < 1)
        return 3;
    // lateral extent of the cap: kr < kmag + dkr and kz > kz_min + dkr with dkr <= max(dk, dkz) give
    // kxy^2 = kr^2 - kz^2 < kmag^2 - kz_min^2 + 2 dmax (kmag - kz_min)
    const double dmax = std::fmax(p.dk, p.dkz);
    const double kxy2 = p.kmag * p.kmag - p.kz_min * p.kz_min + 2.0 * dmax * (p.kmag - p.kz_min);
    const int H = (int)std::floor(std::sqrt(kxy2) / p.kstep * (1.0 + 1e-9));
    if (H < 1 || H > N / 4) return 3;        // (wider boxes: K1's staged table / band-limited rows do not apply; three-pass path)
    ShepZArgs<T> a;
    a.p = p; a.H = H; a.HP = dtab_pitch(-H); a.ntab = dtab_size(-H);
    a.band0[0] = band0[0]; a.band0[1] = band0[1]; a.bandn[0] = bandn[0]; a.bandn[1] = bandn[1];
    a.scale = 1.0 / (double)NZ;
    const int nband = bandn[0] + bandn[1], ZC = NZ < SHEPZ_ZC ? NZ : SHEPZ_ZC;
    const size_t smem = sizeof(cplx<T>) * ((size_t)ZC * SHEPZ_TP + NZ) + (sizeof(T) + sizeof(int)) * (size_t)SHEPZ_COLS * nband;
    int dev = 0, smem_limit = 0;
    PYOTF_CUDA_OK(cudaGetDevice(&dev));
    PYOTF_CUDA_OK(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem > (size_t)smem_limit) return 3;
    auto kern = sheppard_zplanes_kernel<T>;
    static thread_local size_t configured[PYOTF_MAX_DEVICES] = {};
    const int dev_slot = current_device_slot();
    if (smem > configured[dev_slot]) {
        PYOTF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev_slot] = smem;
    }
    keep_scratch_pool();
    const size_t tab_bytes = sizeof(cplx<T>) * (size_t)NZ * a.ntab;
    PYOTF_CUDA_OK(cudaMallocAsync(&a.tab, tab_bytes + hanser_tables_aux_bytes(N, NZ, sizeof(cplx<T>)), s));
    a.twn = reinterpret_cast<cplx<T>*>(reinterpret_cast<char*>(a.tab) + tab_bytes);
    dim3 grid((unsigned)((a.HP + SHEPZ_COLS - 1) / SHEPZ_COLS), (unsigned)(H + 1));
    kern<<<grid, 256, smem, s>>>(a);
    cudaError_t e = cudaGetLastError();
    int rc = 0;
    if (e != cudaSuccess) { set_error(std::string("sheppard_zplanes_kernel: ") + cudaGetErrorString(e)); rc = 2; }
    if (!rc) rc = hanser_launch_tables<T>(N, -H, a.tab, NZ, psfa_out, psfi_out, s, a.twn);
    cudaFreeAsync(a.tab, s);
    return rc;
}

int fft_twiddles_f32(int L, const void** out, cudaStream_t s);   // fftnd_f32.cu / fftnd_f64.cu
int fft_twiddles_f64(int L, const void** out, cudaStream_t s);

template <typename T>
int sheppard_fused(const SheppardParams& p, void* psfa_out, void* psfi_out, cudaStream_t s) {
    int dev = 0, smem_limit = 0;
    PYOTF_CUDA_OK(cudaGetDevice(&dev));
    PYOTF_CUDA_OK(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    keep_scratch_pool();
    const size_t nvox = (size_t)p.NZ * p.N * p.N;
    PYOTF_REQUIRE(nvox < ((size_t)1 << 31), "sheppard: volumes of 2^31 voxels or more are not supported");
    cplx<T>* tmp = nullptr;
    unsigned char* planeflag = nullptr;
    ShepArgs<T> a;
    a.p = p;
    {
        // the predicate of sheppard_planeflag_kernel on the host: the occupied planes as runs of the memory index zm
        a.band0[0] = a.band0[1] = 0; a.bandn[0] = a.bandn[1] = 0;
        int runs = 0;
        bool prev = false;
        const double dmax = std::fmax(std::fabs(p.dk), std::fabs(p.dkz));
        for (int zm = 0; zm < p.NZ; ++zm) {
            int zu = zm - p.NZ / 2; if (zu < 0) zu += p.NZ;
            const int f = zu < (p.NZ + 1) / 2 ? zu : zu - p.NZ;
            const double kz = (double)f * p.kzstep, kzz = p.dual ? std::fabs(kz) : kz;
            const bool on = kzz > p.kz_min && std::fabs(kz) <= (p.kmag + dmax) * (1.0 + 1e-9);
            if (on && !prev) { if (runs < 2) a.band0[runs] = zm; ++runs; }
            if (on && runs <= 2) ++a.bandn[runs - 1];
            prev = on;
        }
        if (runs > 2) a.bandn[0] = -1;
    }
    {
        // scalar model, power-of-two sizes: z-first evaluation through the HanserPSF kernel
        const int rcz = sheppard_zfirst<T>(p, a.band0, a.bandn, psfa_out, psfi_out, s);
        if (rcz != 3) return rcz;
    }
    PYOTF_CUDA_OK(cudaMallocAsync(&tmp, nvox * sizeof(cplx<T>), s));
    PYOTF_CUDA_OK(cudaMallocAsync(&planeflag, (size_t)p.NZ * p.N, s));
    // line flags of passes 2 / 3: line (zm, xm) may hold data iff plane zm may hold shell voxels
    sheppard_planeflag_kernel<<<(unsigned)(((size_t)p.NZ * p.N + 255) / 256), 256, 0, s>>>(p, planeflag);
    PYOTF_CUDA_OK(cudaGetLastError());
    a.tmp = tmp; a.planeflag = planeflag;
    const void *twx = nullptr, *twz = nullptr;
    int rc = sizeof(T) == 4 ? fft_twiddles_f32(p.N, &twx, s) : fft_twiddles_f64(p.N, &twx, s);
    if (!rc) rc = sizeof(T) == 4 ? fft_twiddles_f32(p.NZ, &twz, s) : fft_twiddles_f64(p.NZ, &twz, s);
    a.tw_x = (const cplx<T>*)twx; a.tw_z = (const cplx<T>*)twz;
    // easy_ifft = ifftshift(ifftn(fftshift(.))): per axis in_shift = n - n//2 ... see fftn_run (inverse, both shifts)
    auto shifts = [](int L, int& in, int& out) { const int half = L / 2, rest = (L - half) % L; in = rest; out = rest; };
    shifts(p.N, a.in_x, a.out_x);
    shifts(p.NZ, a.in_z, a.out_z);
    for (int v = 0; v < p.nvec && !rc; ++v) {
        a.v = v; a.accumulate = v > 0;
        a.psfa = psfa_out ? (cplx<T>*)psfa_out + (size_t)v * nvox : nullptr;
        a.psfi = (T*)psfi_out;
        a.scale = 1.0 / (double)p.N;
        ShepRow<T> f1; f1.a = a;
        rc = fft_lines_launch<1, T, true>(p.N, (long long)p.NZ * p.N, f1, smem_limit, s);
        if (rc) break;
        const long long shape[3] = {p.NZ, p.N, p.N};
        const int axes[1] = {1};
        const unsigned char* pf[1] = {planeflag};
        rc = fftn_run<T>(tmp, tmp, 3, shape, axes, 1, 1, 0, 1, 1, s, pf);
        if (rc) break;
        a.scale = 1.0 / (double)p.NZ;
        if (a.bandn[0] >= 0 && !a.psfa && a.psfi && !a.accumulate) {
            ShepCol<T, true> f3; f3.a = a;
            rc = fft_lines_launch<1, T, false>(p.NZ, (long long)p.N * p.N, f3, smem_limit, s);
        } else {
            ShepCol<T> f3; f3.a = a;
            rc = fft_lines_launch<1, T, false>(p.NZ, (long long)p.N * p.N, f3, smem_limit, s);
        }
    }
    cudaFreeAsync(planeflag, s);
    cudaFreeAsync(tmp, s);
    return rc;
}

}  // namespace pyotf
