// K1bs: HanserPSF._gen_psf + BasePSF.PSFi for the reference's DEFAULT model -- scalar, ideal pupil
// (pyotf/otf.py:362-428 with pupil_base=None, vec_corr="none", and pyotf/otf.py:204-207) -- at N = 1024
// (BASELINE config 4).  The plane does not fit one SM, so it is two passes over an L2-resident intermediate, like
// K1b (hanser_big.cuh), but with K1s's symmetries (hanser_sym.cuh) and its one-lane-per-line mapping:
//
//   pupil * apodisation * defocus depends on (|ky|, |kx|) only and is symmetric under ky <-> kx, so the field is
//   even in y and in x: only the columns kx = 0 .. H and the rows y = 0 .. 511 are transformed, every computed
//   value lands at (+-y, +-x); output row y = 512 is the transpose of output column x = 512.
//   A 1024-point transform X[c1 + 32 c2] = sum_g w32^{g c2} [ w1024^{g c1} sum_q x[32 q + g] w32^{q c1} ] of an
//   even sequence needs 17 + 17 of its 64 radix-32 units: the first-stage result of residue 32 - g is
//       Z_{32-g}[c1] = conj(w1024^{g c1}) * Y_g[-c1],
//   and the outputs of second-stage unit 32 - c1 mirror those of unit c1.  The band limit (H < 256) leaves 16 of
//   the 32 first-stage inputs non-zero (fft32_mid0).
//
//   pass A (columns): CTA = (plane, 16 columns kx).  F(|ky|, kx) = apodisation * exp(2 pi i kz z) / N^2 is evaluated
//                     inside the first-stage load (each (|ky|, kx) exactly once per plane), two radix-32 stages in
//                     registers with one shared-memory exchange, the 16 x 513 result tile is transposed through
//                     shared memory and written as full lines  U[plane][kx][y], y = 0 .. 512.
//   pass B (rows)   : CTA = (plane, 16 rows y).  First stage straight from U (coalesced: lane = y), exchange,
//                     second stage, |.|^2 into a staging tile, then full 4 KB output lines (STG.128) to rows
//                     512 + y and 512 - y (fftshift-ed positions), both mirrored in x from the same tile.
//
// A half-warp executes ONE unit index (g / c1) for 16 lines, one per lane: twiddles and branch conditions are
// uniform per half-warp, shared-memory accesses are unit-stride or odd-stride.  U per plane is 0.9 MB (K1b's
// intermediate: 3.6 MB) and stays in L2 between the passes.  Both kernels are launched with programmatic stream
// serialization: each waits for its predecessor only before its first dependent global access.
#pragma once
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "hanser.cuh"

namespace pyotf {

// radix-32 whose inputs x[8] .. x[23] are zero (not read): even / odd halves are radix-16 transforms with inputs
// 4 .. 11 zero
template <int DIR, typename T> __device__ __forceinline__ void fft32_mid0(cplx<T>* x) {
    cplx<T> e[16], o[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) { e[i] = x[2 * i]; o[i] = x[2 * i + 1]; e[12 + i] = x[24 + 2 * i]; o[12 + i] = x[25 + 2 * i]; }
    fft16_mid0<DIR>(e);
    fft16_mid0<DIR>(o);
#define PYOTF_W32(K) { const cplx<T> t = mul_w32<DIR, K>(o[K]); x[K] = e[K] + t; x[K + 16] = e[K] - t; }
    PYOTF_W32(0) PYOTF_W32(1) PYOTF_W32(2) PYOTF_W32(3) PYOTF_W32(4) PYOTF_W32(5) PYOTF_W32(6) PYOTF_W32(7)
    PYOTF_W32(8) PYOTF_W32(9) PYOTF_W32(10) PYOTF_W32(11) PYOTF_W32(12) PYOTF_W32(13) PYOTF_W32(14) PYOTF_W32(15)
#undef PYOTF_W32
}

template <typename T> struct BigSymCfg {
    static constexpr int N = 1024, R = 32, NU = 17;          // 17 unit indices per stage
    static constexpr int LANES = 16;                         // lines per CTA (one per lane of a half-warp)
    static constexpr int NW = 9, NT = 32 * NW;               // warp w: units w (lower half-warp) and 16 - w (upper)
    static constexpr int NTW = 16 * 16 + 1;                  // w1024^k, k <= 256
    static constexpr int EPC = 16 / (int)sizeof(T);          // reals per 16-byte store
    static constexpr int OQ = 512 / EPC + 1;                 // staging: x = EPC * m + j lives at [j][m], m <= 512 / EPC
    static constexpr int OP = EPC * OQ + 1;                  // odd row pitch (reals)
    static constexpr int SP = 513;                           // odd pitch of the pass-A transpose tile [kx][y]
    static constexpr int UPITCH = 516;                       // U[plane][kx][UPITCH]: y = 0 .. 512, 16-byte aligned rows
    static constexpr size_t e_bytes() { return sizeof(cplx<T>) * (size_t)NU * R * LANES; }
    static constexpr size_t tw_bytes() { return sizeof(cplx<T>) * (size_t)(NTW + 1); }
    static constexpr size_t o_bytes() { return sizeof(T) * (size_t)LANES * OP + 16; }
    static constexpr size_t smem_a() { return tw_bytes() + e_bytes(); }
    static constexpr size_t smem_b() { return tw_bytes() + e_bytes() + o_bytes(); }
    static_assert(sizeof(cplx<T>) * (size_t)LANES * SP <= sizeof(cplx<T>) * (size_t)NU * R * LANES, "transpose tile must fit the exchange");
};

template <typename T> struct BigSymArgs {
    const double* kzc;        // [KR][K]
    const T* ampc;            // [KR][K] apodisation * ideal mask
    const cplx<T>* tw;        // [N] e^{+2 pi i k / N}
    const double* z;          // [nzc] (already offset to the chunk)
    cplx<T>* U;               // [nzc][H + 1][UPITCH]
    T* psfi;                  // [nzc][N][N] of this chunk
    int K, H;
};

// first stage of an even 1024-point inverse transform for residue g, one transform per lane.
// bigsym_load: x[q] = sample of frequency 32 q + g (q < 8) / 32 q + g - 1024 (q >= 24) through ld(|f|), 0 <= |f| <= H < 256.
// bigsym_stage1: dst = E + lane receives Z_g[c1] and, for 0 < g < 16, Z_{32-g}[c1] (c1 = 0 .. 16) at [(c1 * 32 + g) * LANES].
template <typename T, class LD> __device__ __forceinline__ void bigsym_load(LD ld, int H, int g, cplx<T>* x) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int f = 32 * q + g;                       // frequency 32 q + g
        x[q] = f <= H ? ld(f) : mk<T>(0, 0);
    }
#pragma unroll
    for (int q = 24; q < 32; ++q) {
        const int f = 1024 - 32 * q - g;                // frequency 32 q + g - 1024 = -f
        x[q] = f <= H ? ld(f) : mk<T>(0, 0);
    }
}
template <typename T, int LANES>
__device__ __forceinline__ void bigsym_stage1(cplx<T>* x, int g, const cplx<T>* __restrict__ tws, cplx<T>* __restrict__ dst) {
    fft32_mid0<1>(x);
    dst[(size_t)g * LANES] = x[0];
    const bool mir = g > 0 && g < 16;
    if (mir) dst[(size_t)(32 - g) * LANES] = x[0];
#pragma unroll
    for (int c = 1; c <= 16; ++c) {
        const cplx<T> w = tws[g * c];
        dst[(size_t)(c * 32 + g) * LANES] = cmul(x[c], w);
        if (mir) dst[(size_t)(c * 32 + 32 - g) * LANES] = cmulc(x[32 - c], w);
    }
}

// pupil-plane samples of one column for the first-stage loads: every load is issued before the first use
// (kz * z -> e^{2 pi i kz z} -> * apodisation; the branch on the apodisation comes after all loads)
template <typename T, bool EVEN, class PF>
__device__ __forceinline__ void pupil_column_load(const double* __restrict__ kzc, PF pf, int K, int H, int g, bool col, double z, cplx<T>* x) {
    double kz[16];
    cplx<T> pv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int q = i < 8 ? i : i + 16;
        const int fs = i < 8 ? 32 * q + g : 32 * q + g - 1024;        // signed frequency
        const int f = EVEN && fs < 0 ? -fs : fs;
        const bool ok = col && (fs >= 0 ? fs <= H : -fs <= H);
        kz[i] = ok ? kzc[(ptrdiff_t)f * K] : 0.0;
        pv[i] = ok ? pf((ptrdiff_t)f * K) : mk<T>(0, 0);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int q = i < 8 ? i : i + 16;
        cplx<T> d = mk<T>(0, 0);
        if (pv[i].x != (T)0 || pv[i].y != (T)0) {
            cis_cycles(kz[i] * z, d);                                   // _calc_defocus   otf.py:357-360
            d = cmul(d, pv[i]);
        }
        x[q] = d;
    }
}

// ---- pass A: columns ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BigSymCfg<T>::NT, sizeof(T) == 4 ? 2 : 1) hanser_bigsym_cols_kernel(BigSymArgs<T> a) {
    using C = BigSymCfg<T>;
    constexpr int LANES = C::LANES, NT = C::NT;
    pdl_launch_dependents();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* const tws = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* const E = reinterpret_cast<cplx<T>*>(smem_raw + C::tw_bytes());
    const int tid = threadIdx.x, l16 = tid & 15, warp = tid >> 5, half = (tid >> 4) & 1;
    const int unit = half ? 16 - warp : warp;
    const bool active = !(half && warp == 8);
    const int H = a.H, NC = H + 1, K = a.K;
    const int plane = blockIdx.y, kx0 = blockIdx.x * LANES, kx = kx0 + l16;
    const bool col = kx < NC;
    const double z = a.z[plane];
    cplx<T> x[32];
    if (active) {
        const T scale = (T)1.0 / ((T)C::N * (T)C::N);
        const T* __restrict__ amp = a.ampc + (size_t)H * K + H + kx;
        auto pf = [&](ptrdiff_t o) { return mk<T>(amp[o] * scale, (T)0); };
        pupil_column_load<T, true>(a.kzc + (size_t)H * K + H + kx, pf, K, H, unit, col, z, x);
    }
    for (int k = tid; k < C::NTW; k += NT) tws[k] = a.tw[k];
    __syncthreads();
    if (active) bigsym_stage1<T, LANES>(x, unit, tws, E + l16);
    __syncthreads();
    if (active) {
#pragma unroll
        for (int g = 0; g < 32; ++g) x[g] = E[(size_t)(unit * 32 + g) * LANES + l16];
        fft32<1>(x);
    }
    __syncthreads();                                               // E is free: it becomes the transpose tile S[kx][y]
    cplx<T>* const S = E;
    if (active) {
        cplx<T>* const s = S + (size_t)l16 * C::SP;
        if (unit == 0) {
#pragma unroll
            for (int c2 = 0; c2 <= 16; ++c2) s[32 * c2] = x[c2];
        } else {
#pragma unroll
            for (int c2 = 0; c2 < 16; ++c2) s[unit + 32 * c2] = x[c2];
            if (unit < 16) {
#pragma unroll
                for (int c2 = 16; c2 < 32; ++c2) s[1024 - unit - 32 * c2] = x[c2];
            }
        }
    }
    __syncthreads();
    pdl_wait();                                                    // U may still be read by the previous chunk's pass B
    const int ncol = min(LANES, NC - kx0);
    cplx<T>* const Ub = a.U + ((size_t)plane * NC + kx0) * C::UPITCH;
    for (int i = tid; i < ncol * C::SP; i += NT) {
        const int c = i / C::SP, y = i - c * C::SP;
        Ub[(size_t)c * C::UPITCH + y] = S[i];
    }
}

// ---- pass B: rows ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BigSymCfg<T>::NT, sizeof(T) == 4 ? 2 : 1) hanser_bigsym_rows_kernel(BigSymArgs<T> a) {
    using C = BigSymCfg<T>;
    constexpr int LANES = C::LANES, NT = C::NT, N = C::N, EPC = C::EPC, OQ = C::OQ, OP = C::OP;
    pdl_launch_dependents();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* const tws = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* const E = reinterpret_cast<cplx<T>*>(smem_raw + C::tw_bytes());
    T* const O = reinterpret_cast<T*>(smem_raw + C::tw_bytes() + C::e_bytes());
    const int tid = threadIdx.x, lane = tid & 31, l16 = tid & 15, warp = tid >> 5, half = (tid >> 4) & 1;
    const int unit = half ? 16 - warp : warp;
    const bool active = !(half && warp == 8);
    const int H = a.H, NC = H + 1;
    const int plane = blockIdx.y, y0 = blockIdx.x * LANES, y = y0 + l16;
    T* const psfi = a.psfi + (size_t)plane * N * N;
    pdl_wait();                                                    // U comes from pass A
    const cplx<T>* __restrict__ Up = a.U + (size_t)plane * NC * C::UPITCH;
    cplx<T> x[32];
    if (active) {
        const cplx<T>* __restrict__ u = Up + y;
        auto ld = [&](int f) { return u[(size_t)f * C::UPITCH]; };
        bigsym_load<T>(ld, H, unit, x);
    }
    for (int k = tid; k < C::NTW; k += NT) tws[k] = a.tw[k];
    __syncthreads();
    if (active) bigsym_stage1<T, LANES>(x, unit, tws, E + l16);
    __syncthreads();
    if (active) {
#pragma unroll
        for (int g = 0; g < 32; ++g) x[g] = E[(size_t)(unit * 32 + g) * LANES + l16];
        fft32<1>(x);
        // |.|^2 at x = unit + 32 c2 (and, mirrored, at 1024 - x); staged as [x % EPC][x / EPC]
        T* const o = O + (size_t)l16 * OP;
        const int lo = (unit % EPC) * OQ + unit / EPC;
        const int hi = ((32 - unit) % EPC) * OQ + (32 - unit) / EPC;     // x = 1024 - unit - 32 c2 = (32 - unit) + 32 (31 - c2)
#pragma unroll
        for (int c2 = 0; c2 < 16; ++c2) o[lo + (32 / EPC) * c2] = x[c2].x * x[c2].x + x[c2].y * x[c2].y;
        if (unit == 0) {
            const T pw = x[16].x * x[16].x + x[16].y * x[16].y;         // x = 512
            o[512 / EPC] = pw;
            // output row y = 512 (stored at row 0) is the transpose of output column x = 512
            psfi[512 + y] = pw;
            if (y != 0) psfi[512 - y] = pw;
        } else if (unit < 16) {
#pragma unroll
            for (int c2 = 16; c2 < 32; ++c2) o[hi + (32 / EPC) * (31 - c2)] = x[c2].x * x[c2].x + x[c2].y * x[c2].y;
        }
    }
    // corner: out[512][512] (stored at [0][0]) = | sum_kx (-1)^kx U[|kx|][512] |^2
    if (blockIdx.x == 0 && warp == 8 && half) {
        cplx<T> s = mk<T>(0, 0);
        for (int kx = l16; kx < NC; kx += 16) {
            const cplx<T> u = Up[(size_t)kx * C::UPITCH + 512];
            const T wgt = (kx == 0 ? (T)1 : (T)2) * ((kx & 1) ? (T)-1 : (T)1);
            s.x += wgt * u.x; s.y += wgt * u.y;
        }
#pragma unroll
        for (int of = 8; of; of >>= 1) { s.x += __shfl_xor_sync(0xffff0000u, s.x, of); s.y += __shfl_xor_sync(0xffff0000u, s.y, of); }
        if (l16 == 0) psfi[0] = s.x * s.x + s.y * s.y;
    }
    __syncthreads();
    // ---- staged rows -> output rows 512 + y and 512 - y, 16 bytes per lane and instruction ----------------------
    constexpr int NCH = N / EPC / 32;                   // 32-lane groups of 16-byte chunks per row
    constexpr int NITEM = LANES * NCH, IU = 4;          // items in flight per warp
    for (int it0 = warp; it0 < NITEM; it0 += IU * C::NW) {
        T v[IU][EPC];
#pragma unroll
        for (int u = 0; u < IU; ++u) {
            const int it = it0 + u * C::NW;
            if (it < NITEM) {
                const int row = it / NCH, k = it - row * NCH;
                const T* const o = O + (size_t)row * OP;
#pragma unroll
                for (int j = 0; j < EPC; ++j) {
                    const int xs = (lane + 32 * k) * EPC + j;
                    const int ax = xs < 512 ? 512 - xs : xs - 512;
                    v[u][j] = o[(ax % EPC) * OQ + ax / EPC];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < IU; ++u) {
            const int it = it0 + u * C::NW;
            if (it < NITEM) {
                const int row = it / NCH, k = it - row * NCH;
                const int yy = y0 + row, xs = (lane + 32 * k) * EPC;
                T* const r0 = psfi + (size_t)(512 + yy) * N + xs;
                T* const r1 = psfi + (size_t)(512 - yy) * N + xs;
                if constexpr (EPC == 4) {
                    const float4 q = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
                    *reinterpret_cast<float4*>(r0) = q;
                    if (yy != 0) *reinterpret_cast<float4*>(r1) = q;
                } else {
                    const double2 q = make_double2(v[u][0], v[u][1]);
                    *reinterpret_cast<double2*>(r0) = q;
                    if (yy != 0) *reinterpret_cast<double2*>(r1) = q;
                }
            }
        }
    }
}

// =================================================================================================================
// K1br: the same two register-resident passes WITHOUT the symmetry -- any pupil (aberrated, retrieved, user-supplied),
// any apodisation / polarisation component, at N = 1024 (BASELINE config 4: 1024^2 planes with Zernike aberrations).
// A quarter-warp executes one of the 32 unit indices of a stage for 8 lines.
//   pass A: CTA = (plane, 8 columns kx): pupil * apodisation * exp(2 pi i kz z) / N^2 built inside the first-stage
//           load, W[plane][kx][y] written as full 8 KB lines after a shared-memory transpose
//   pass B: CTA = (plane, 8 rows y): first stage straight from W (lane = y), |.|^2, full fftshift-ed output lines
// =================================================================================================================
template <typename T> struct BigGenCfg {
    static constexpr int N = 1024, R = 32;
    static constexpr int LANES = 8;                          // lines per CTA (one per lane of a quarter-warp)
    static constexpr int NT = 256;                           // 32 quarter-warps = the 32 unit indices of a stage
    static constexpr int EPC = 16 / (int)sizeof(T);
    static constexpr int OP = N + EPC;                       // staging O[row][x], rows 16 bytes (4 banks) apart: the per-lane writes of a
                                                             // warp (8 rows x 4 consecutive x) hit 32 different banks
    static constexpr int SP = N + 1;                         // odd pitch of the pass-A transpose tile [kx][y]
    static constexpr int EC = R + 1;                         // exchange E[c1][EC][LANES]: units 4 w .. 4 w + 3 of a warp read disjoint banks
    static constexpr size_t tw_bytes() { return sizeof(cplx<T>) * (size_t)N; }
    static constexpr size_t e_bytes() { return sizeof(cplx<T>) * (size_t)R * EC * LANES; }   // >= LANES * SP: exchange, then transpose tile
    static_assert((size_t)R * EC * LANES >= (size_t)LANES * SP, "transpose tile must fit the exchange");
    static constexpr size_t o_bytes() { return sizeof(T) * (size_t)LANES * OP + 16; }
    static constexpr size_t smem_a() { return tw_bytes() + e_bytes(); }
    static constexpr size_t smem_b() { return tw_bytes() + e_bytes() + o_bytes(); }
};

template <typename T> struct BigGenArgs {
    const double* kzc;        // [KR][K]
    const T* ampc;            // [KR][K] of the current polarisation component
    const cplx<T>* pupilc;    // [KR][K] compact pupil of this batch entry or nullptr (ideal: the mask is in ampc)
    const cplx<T>* tw;        // [N]
    const double* z;          // [nzc]
    cplx<T>* W;               // [nzc][K][N]
    T* psfi;                  // [nzc][N][N]
    int K, H, accumulate;
    int bulk;                 // output lines leave through bulk asynchronous copies (shared -> global) instead of STG.128
};

// first stage of a band-limited (|f| <= H < 256) 1024-point inverse transform for residue g
// biggen_load: x[q] = sample of the signed frequency 32 q + g (q < 8) / 32 q + g - 1024 (q >= 24) through ld(f)
template <typename T, class LD> __device__ __forceinline__ void biggen_load(LD ld, int H, int g, cplx<T>* x) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int f = 32 * q + g;
        x[q] = f <= H ? ld(f) : mk<T>(0, 0);
    }
#pragma unroll
    for (int q = 24; q < 32; ++q) {
        const int f = 32 * q + g - 1024;
        x[q] = f >= -H ? ld(f) : mk<T>(0, 0);
    }
}
template <typename T, int LANES, int EC>
__device__ __forceinline__ void biggen_stage1(cplx<T>* x, int g, const cplx<T>* __restrict__ tws, cplx<T>* __restrict__ dst) {
    fft32_mid0<1>(x);
    dst[(size_t)g * LANES] = x[0];
#pragma unroll
    for (int c = 1; c < 32; ++c) dst[(size_t)(c * EC + g) * LANES] = cmul(x[c], tws[(g * c) & 1023]);
}

template <typename T>
__global__ void __launch_bounds__(BigGenCfg<T>::NT, sizeof(T) == 4 ? 2 : 1) hanser_biggen_cols_kernel(BigGenArgs<T> a) {
    using C = BigGenCfg<T>;
    constexpr int LANES = C::LANES, NT = C::NT;
    pdl_launch_dependents();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* const tws = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* const E = reinterpret_cast<cplx<T>*>(smem_raw + C::tw_bytes());
    const int tid = threadIdx.x, l8 = tid & 7, unit = tid >> 3;
    const int H = a.H, K = a.K;
    const int plane = blockIdx.y, ix0 = blockIdx.x * LANES, ix = ix0 + l8;
    const bool col = ix < K;
    const double z = a.z[plane];
    cplx<T> x[32];
    {
        const T scale = (T)1.0 / ((T)C::N * (T)C::N);
        const T* __restrict__ amp = a.ampc + (size_t)H * K + ix;
        if (a.pupilc) {
            const cplx<T>* __restrict__ pup = a.pupilc + (size_t)H * K + ix;
            auto pf = [&](ptrdiff_t o) { return cscale(pup[o], amp[o] * scale); };
            pupil_column_load<T, false>(a.kzc + (size_t)H * K + ix, pf, K, H, unit, col, z, x);
        } else {
            auto pf = [&](ptrdiff_t o) { return mk<T>(amp[o] * scale, (T)0); };
            pupil_column_load<T, false>(a.kzc + (size_t)H * K + ix, pf, K, H, unit, col, z, x);
        }
    }
    for (int k = tid; k < C::N; k += NT) tws[k] = a.tw[k];
    __syncthreads();
    biggen_stage1<T, LANES, C::EC>(x, unit, tws, E + l8);
    __syncthreads();
#pragma unroll
    for (int g = 0; g < 32; ++g) x[g] = E[(size_t)(unit * C::EC + g) * LANES + l8];
    fft32<1>(x);
    __syncthreads();                                               // the exchange becomes the transpose tile S[kx][y]
    cplx<T>* const S = E;
    {
        cplx<T>* const s = S + (size_t)l8 * C::SP + unit;
#pragma unroll
        for (int c2 = 0; c2 < 32; ++c2) s[32 * c2] = x[c2];
    }
    __syncthreads();
    pdl_wait();                                                    // W may still be read by the previous chunk's pass B
    const int ncol = min(LANES, K - ix0);
    cplx<T>* const Wb = a.W + ((size_t)plane * K + ix0) * C::N;
    for (int i = tid; i < ncol * C::N; i += NT) {
        const int c = i >> 10, y = i & 1023;
        Wb[i] = S[(size_t)c * C::SP + y];
    }
}

template <typename T>
__global__ void __launch_bounds__(BigGenCfg<T>::NT, sizeof(T) == 4 ? 2 : 1) hanser_biggen_rows_kernel(BigGenArgs<T> a) {
    using C = BigGenCfg<T>;
    constexpr int LANES = C::LANES, NT = C::NT, N = C::N, EPC = C::EPC, OP = C::OP;
    pdl_launch_dependents();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* const tws = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* const E = reinterpret_cast<cplx<T>*>(smem_raw + C::tw_bytes());
    T* const O = reinterpret_cast<T*>(smem_raw + C::tw_bytes() + C::e_bytes());
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, l8 = tid & 7, unit = tid >> 3;
    const int H = a.H, K = a.K;
    const int plane = blockIdx.y, y0 = blockIdx.x * LANES, y = y0 + l8;
    T* const psfi = a.psfi + (size_t)plane * N * N;
    pdl_wait();                                                    // W comes from pass A
    cplx<T> x[32];
    {
        const cplx<T>* __restrict__ w = a.W + ((size_t)plane * K + H) * N + y;
        auto ld = [&](int f) { return w[(ptrdiff_t)f * N]; };
        biggen_load<T>(ld, H, unit, x);
    }
    for (int k = tid; k < N; k += NT) tws[k] = a.tw[k];
    __syncthreads();
    biggen_stage1<T, LANES, C::EC>(x, unit, tws, E + l8);
    __syncthreads();
    {
#pragma unroll
        for (int g = 0; g < 32; ++g) x[g] = E[(size_t)(unit * C::EC + g) * LANES + l8];
        fft32<1>(x);
        T* const o = O + (size_t)l8 * OP + unit;                  // x = unit + 32 c2
#pragma unroll
        for (int c2 = 0; c2 < 32; ++c2) o[32 * c2] = x[c2].x * x[c2].x + x[c2].y * x[c2].y;
    }
    if (!a.accumulate && a.bulk) {
        // ---- staged rows -> fftshift-ed output lines by the copy engine: two bulk asynchronous copies per row
        //      (x >= 512 | x < 512 swap halves), shared -> global, SASS UBLKCP ------------------------------------
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the staged rows are read by the async proxy next
        __syncthreads();
        if (tid < 2 * LANES) {
            const int row = tid >> 1, half = tid & 1;
            const int ys = (y0 + row + N / 2) & (N - 1);
            bulk_s2g(psfi + (size_t)ys * N + (half ? 0 : N / 2), O + (size_t)row * OP + (half ? N / 2 : 0), (unsigned)(N / 2 * sizeof(T)));
            bulk_commit_and_wait_read();                                   // the staging tile must outlive the engine's reads
        }
        return;
    }
    __syncthreads();
    // ---- staged rows -> fftshift-ed output lines, 16 bytes per lane and instruction (further polarisation components:
    //      read-modify-write) ------------------------------------------------------------------------------------------
    constexpr int NCH = N / EPC / 32, NITEM = LANES * NCH, NW = NT / 32;
    for (int it = warp; it < NITEM; it += NW) {
        const int row = it / NCH, k = it - row * NCH;
        const int xs = (lane + 32 * k) * EPC, xu = (xs + N / 2) & (N - 1);   // output column xs holds x = (xs + 512) mod 1024
        const int ys = (y0 + row + N / 2) & (N - 1);
        const T* const o = O + (size_t)row * OP + xu;
        T* const r0 = psfi + (size_t)ys * N + xs;
        if constexpr (EPC == 4) {
            float4 q = *reinterpret_cast<const float4*>(o);
            if (a.accumulate) { const float4 p = *reinterpret_cast<const float4*>(r0); q.x += p.x; q.y += p.y; q.z += p.z; q.w += p.w; }
            *reinterpret_cast<float4*>(r0) = q;
        } else {
            double2 q = *reinterpret_cast<const double2*>(o);
            if (a.accumulate) { const double2 p = *reinterpret_cast<const double2*>(r0); q.x += p.x; q.y += p.y; }
            *reinterpret_cast<double2*>(r0) = q;
        }
    }
}

template <class K, class A>
inline int bigsym_launch_pdl(K kern, dim3 grid, int nt, size_t smem, cudaStream_t s, const A& args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3((unsigned)nt); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PYOTF_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, args));
    return 0;
}

// K1bs: scalar model with the ideal pupil at N = 1024, PSFi only.  *done stays false when the configuration does
// not fit (the caller falls back to the generic two-pass path).
template <typename T>
int hanser_bigsym_launch(const pyotf_hanser* h, const double* z, int nz, int batch, void* psfi, int smem_limit, cudaStream_t s,
                         bool* done) {
    using C = BigSymCfg<T>;
    *done = false;
    const int H = -h->f0;
    if (h->N != C::N || h->K != 2 * H + 1 || H > 255 || H < 1 || h->nvec != 1 || !psfi) return 0;
    if (C::smem_b() > (size_t)smem_limit) return 0;
    auto ka = hanser_bigsym_cols_kernel<T>;
    auto kb = hanser_bigsym_rows_kernel<T>;
    static thread_local bool configured[PYOTF_MAX_DEVICES] = {};
    const int dev_slot = current_device_slot();
    if (!configured[dev_slot]) {
        PYOTF_CUDA_OK(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_a()));
        PYOTF_CUDA_OK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_b()));
        configured[dev_slot] = true;
    }
    const int NC = H + 1, nkb = (NC + C::LANES - 1) / C::LANES;
    // planes per chunk: the intermediate of a chunk stays L2-resident next to the output stream
    static const int chunk_env = [] { const char* e = getenv("PYOTF_B200_K1BS_CHUNK"); return e ? atoi(e) : 0; }();
    const size_t uplane = sizeof(cplx<T>) * (size_t)NC * C::UPITCH;
    int nzc = chunk_env > 0 ? chunk_env : (int)std::max<size_t>(1, ((size_t)40 << 20) / uplane);
    nzc = std::min(nzc, nz);
    cplx<T>* U = nullptr;
    keep_scratch_pool();
    PYOTF_CUDA_OK(cudaMallocAsync(&U, uplane * nzc, s));
    int rc = 0;
    for (int b = 0; b < batch && !rc; ++b)
        for (int z0 = 0; z0 < nz && !rc; z0 += nzc) {
            const int nc = std::min(nzc, nz - z0);
            BigSymArgs<T> a;
            a.kzc = h->kzc; a.ampc = (const T*)h->ampc; a.tw = (const cplx<T>*)h->tw; a.z = z + z0; a.U = U;
            a.psfi = (T*)psfi + ((size_t)b * nz + z0) * (size_t)C::N * C::N;
            a.K = h->K; a.H = H;
            rc = bigsym_launch_pdl(ka, dim3((unsigned)nkb, (unsigned)nc), C::NT, C::smem_a(), s, a);
            if (!rc) rc = bigsym_launch_pdl(kb, dim3((unsigned)(512 / C::LANES), (unsigned)nc), C::NT, C::smem_b(), s, a);
        }
    cudaFreeAsync(U, s);
    if (!rc) { PYOTF_CUDA_OK(cudaGetLastError()); *done = true; }
    return rc;
}

// K1br: any pupil / any polarisation component at N = 1024, PSFi only.  *done stays false when it does not apply.
template <typename T>
int hanser_biggen_launch(const pyotf_hanser* h, const double* z, int nz, const void* pupilc, int batch, long long pupil_stride,
                         void* psfi, int smem_limit, cudaStream_t s, bool* done) {
    using C = BigGenCfg<T>;
    *done = false;
    const int H = -h->f0;
    if (h->N != C::N || h->K != 2 * H + 1 || H > 255 || H < 1 || !psfi) return 0;
    if (C::smem_b() > (size_t)smem_limit) return 0;
    auto ka = hanser_biggen_cols_kernel<T>;
    auto kb = hanser_biggen_rows_kernel<T>;
    static thread_local bool configured[PYOTF_MAX_DEVICES] = {};
    const int dev_slot = current_device_slot();
    if (!configured[dev_slot]) {
        PYOTF_CUDA_OK(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_a()));
        PYOTF_CUDA_OK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_b()));
        configured[dev_slot] = true;
    }
    const int K = h->K, nkb = (K + C::LANES - 1) / C::LANES;
    const size_t KKP = (size_t)h->KR * K;
    static const int chunk_env = [] { const char* e = getenv("PYOTF_B200_K1BR_CHUNK"); return e ? atoi(e) : 0; }();
    const size_t wplane = sizeof(cplx<T>) * (size_t)K * C::N;
    int nzc = chunk_env > 0 ? chunk_env : (int)std::max<size_t>(1, ((size_t)48 << 20) / wplane);
    nzc = std::min(nzc, nz);
    cplx<T>* W = nullptr;
    keep_scratch_pool();
    PYOTF_CUDA_OK(cudaMallocAsync(&W, wplane * nzc, s));
    int rc = 0;
    for (int b = 0; b < batch && !rc; ++b)
        for (int z0 = 0; z0 < nz && !rc; z0 += nzc) {
            const int nc = std::min(nzc, nz - z0);
            for (int v = 0; v < h->nvec && !rc; ++v) {
                BigGenArgs<T> a;
                a.kzc = h->kzc; a.ampc = (const T*)h->ampc + (size_t)v * KKP;
                a.pupilc = pupilc ? (const cplx<T>*)pupilc + (size_t)b * pupil_stride : nullptr;
                a.tw = (const cplx<T>*)h->tw; a.z = z + z0; a.W = W;
                a.psfi = (T*)psfi + ((size_t)b * nz + z0) * (size_t)C::N * C::N;
                a.K = K; a.H = H; a.accumulate = v > 0;
                static const bool bulk = getenv("PYOTF_B200_K1BR_BULK") != nullptr;      // A/B knob (measured 2 % slower than STG.128)
                a.bulk = bulk && ((size_t)psfi % 16 == 0) ? 1 : 0;
                rc = bigsym_launch_pdl(ka, dim3((unsigned)nkb, (unsigned)nc), C::NT, C::smem_a(), s, a);
                if (!rc) rc = bigsym_launch_pdl(kb, dim3((unsigned)(C::N / C::LANES), (unsigned)nc), C::NT, C::smem_b(), s, a);
            }
        }
    cudaFreeAsync(W, s);
    if (!rc) { PYOTF_CUDA_OK(cudaGetLastError()); *done = true; }
    return rc;
}

}  // namespace pyotf
