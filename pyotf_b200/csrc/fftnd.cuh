// K3a: batched 1D FFT passes along one axis of a strided N-d array, with the reference's
// fftshift / ifftshift folded into the load / store indices.  Composed per axis they give
//   easy_fft  = fftshift(fftn(ifftshift(x)))   pyotf/utils.py:15-17  (OTFi, Hanser OTFa)
//   easy_ifft = ifftshift(ifftn(fftshift(x)))  pyotf/utils.py:20-22  (Sheppard PSFa)
//
// One CTA stages a tile of TI lines in shared memory ([TI][L+1] complex, conflict-free for
// both the coalesced global side and the per-line butterflies), runs an in-place
// mixed-radix (16/8/4/2) FFT with digit-reversed output positions, and stores the tile
// back coalesced.  Non power-of-two lengths take a direct DFT on the same staging.
// HBM traffic per pass = one read + one write of the array (the roofline for this kernel).
#pragma once
#include "fft_regs.cuh"

namespace pyotf {

struct AxisPass {
    long long n_outer;     // number of slabs before the axis
    long long n_inner;     // number of elements after the axis (1 => axis is contiguous)
    int L;                 // axis length
    int in_shift;          // x_in[i] = src[(i + in_shift) % L]
    int out_shift;         // dst[(o + out_shift) % L] = X[o]
    double scale;          // applied to outputs (1/L for inverse)
    const void* tw;        // [L] e^{+2 pi i k / L} in the transform's dtype (device; cached per length)
    const unsigned char* line_flags;   // optional: one byte per line (outer * n_inner + inner); 0 = the line is
                                       // entirely zero on input, so (in place) it is left untouched
    int herm_Z = 0, herm_Y = 0, herm_Yshift = 0;   // last pass of the real 3D transform (contiguous axis only): the source holds
                                       // the planes kz = 0 .. Z/2 of [kz][y-shifted][x]; line (kz, ys) is stored to the
                                       // fftshift-ed plane (kz + Z/2) % Z and, for 0 < kz < Z/2, its complex conjugate
                                       // to the Hermitian partner ((-zs) % Z, (-ys) % Y, (-xs) % X)
};

// radix sequence for power-of-two L (product == L): up to three stages of radix <= 16
__host__ __device__ constexpr int radix_count(int L) { return L <= 16 ? 1 : (L <= 256 ? 2 : 3); }
__host__ __device__ constexpr int radix_at(int L, int s) {
    switch (L) {
        case 2: return s == 0 ? 2 : 1;
        case 4: return s == 0 ? 4 : 1;
        case 8: return s == 0 ? 8 : 1;
        case 16: return s == 0 ? 16 : 1;
        case 32: return s == 0 ? 8 : (s == 1 ? 4 : 1);
        case 64: return s <= 1 ? 8 : 1;
        case 128: return s == 0 ? 16 : (s == 1 ? 8 : 1);
        case 256: return s <= 1 ? 16 : 1;
        case 512: return 8;
        case 1024: return s == 0 ? 16 : 8;
        case 2048: return s <= 1 ? 16 : 8;
        case 4096: return 16;
        default: return 1;
    }
}

// Shared-memory layout of a tile: line t starts at t * fft_pitch(L); element e of a line sits at
// fft_pad(e) = e + e/16 (power-of-two lengths).  With the pitch odd (mod 16 elements) the 16 lines of a
// tile fall into 16 different bank groups for any fixed element index, and the +e/16 skew spreads the
// stride-16 / stride-8 accesses of the late stages and of the digit-reversed read-out.  Without it 87 % of
// the shared-memory wavefronts of these kernels were bank-conflict replays (profiles/README.md).
template <bool POW2> __host__ __device__ __forceinline__ constexpr int fft_pad(int e) { return POW2 ? e + (e >> 4) : e; }
template <bool POW2> __host__ __device__ __forceinline__ constexpr int fft_pitch(int L) {
    return POW2 ? ((L + (L >> 4)) | 1) : L + 1;
}

// one in-place stage over all lines of the tile: blocks of length LB, radix R.  Consecutive threads take
// the SAME butterfly of consecutive LINES (conflict-free: the pitch is odd), then the next butterfly.
template <int DIR, typename T, int L, int LB, int R, int TI>
__device__ __forceinline__ void smem_stage(cplx<T>* tile, int pitch, int nlines, const cplx<T>* tw) {
    constexpr int LS = LB / R;            // sub-block length after this stage
    constexpr int NB = L / R;             // butterflies per line
    auto butterfly = [&](int line, int t) {
        const int blk = t / LS, bp = t - blk * LS;
        cplx<T>* base = tile + (size_t)line * pitch;
        const int e0 = blk * LB + bp;
        cplx<T> x[R];
#pragma unroll
        for (int a = 0; a < R; ++a) x[a] = base[fft_pad<true>(e0 + LS * a)];
        fft_radix<DIR, R>(x);
#pragma unroll
        for (int c = 0; c < R; ++c) {
            if (LS > 1 && c > 0) {
                cplx<T> w0 = tw[(bp * c * (L / LB)) & (L - 1)];
                x[c] = DIR > 0 ? cmul(x[c], w0) : cmulc(x[c], w0);
            }
            base[fft_pad<true>(e0 + LS * c)] = x[c];
        }
    };
    if (nlines == TI) {
        for (int w = threadIdx.x; w < TI * NB; w += blockDim.x) butterfly(w % TI, w / TI);
    } else {
        for (int w = threadIdx.x; w < nlines * NB; w += blockDim.x) butterfly(w % nlines, w / nlines);
    }
}

// position (within a line) holding natural-order output index o after the staged FFT
template <int L> __device__ __forceinline__ int digit_reversed_pos(int o) {
    int pos = 0, lb = L;
#pragma unroll
    for (int s = 0; s < radix_count(L); ++s) {
        const int r = radix_at(L, s);
        lb /= r;
        pos += (o % r) * lb;
        o /= r;
    }
    return pos;
}

template <typename TIN, typename T> __device__ __forceinline__ cplx<T> load_as_cplx(const TIN* src, size_t i);
template <> __device__ __forceinline__ cplx<float> load_as_cplx<float, float>(const float* s, size_t i) { return mk<float>(s[i], 0.f); }
template <> __device__ __forceinline__ cplx<double> load_as_cplx<double, double>(const double* s, size_t i) { return mk<double>(s[i], 0.0); }
template <> __device__ __forceinline__ cplx<float> load_as_cplx<cplx<float>, float>(const cplx<float>* s, size_t i) { return s[i]; }
template <> __device__ __forceinline__ cplx<double> load_as_cplx<cplx<double>, double>(const cplx<double>* s, size_t i) { return s[i]; }

// TIN = T (real input) or cplx<T>.  POW2 selects the staged FFT, otherwise a direct DFT.
template <int DIR, typename TIN, typename T, int L, int TI, bool POW2>
__global__ void __launch_bounds__(256) fft_axis_kernel(AxisPass p, const TIN* __restrict__ src, cplx<T>* __restrict__ dst) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Lr = POW2 ? L : p.L;
    const int pitch = fft_pitch<POW2>(Lr);
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);        // [Lr]  e^{+2 pi i k / L}
    cplx<T>* tile = tw + Lr;                                    // [TI][pitch]
    cplx<T>* tile2 = tile + (size_t)TI * pitch;                 // DFT output (non-pow2 only)

    {
        const cplx<T>* __restrict__ twg = reinterpret_cast<const cplx<T>*>(p.tw);
        for (int k = threadIdx.x; k < Lr; k += blockDim.x) tw[k] = twg[k];
    }
    // tile -> (outer, inner0)
    const long long tiles_per_outer = (p.n_inner + TI - 1) / TI;
    const bool contiguous = p.n_inner == 1;
    long long outer, inner0;
    int nlines;
    if (contiguous) {
        // lines are consecutive "outer" slabs
        outer = (long long)blockIdx.x * TI;
        inner0 = 0;
        nlines = (int)min((long long)TI, p.n_outer - outer);
    } else {
        outer = blockIdx.x / tiles_per_outer;
        inner0 = (blockIdx.x % tiles_per_outer) * TI;
        nlines = (int)min((long long)TI, p.n_inner - inner0);
    }
    const size_t base = contiguous ? (size_t)outer * Lr : (size_t)outer * Lr * p.n_inner + inner0;
    const size_t stride_t = contiguous ? (size_t)Lr : 1, stride_l = contiguous ? 1 : (size_t)p.n_inner;
    if (p.line_flags) {
        // lines known to be all-zero on input transform to zero: skip the tile (zero-fill when out of place)
        const size_t line0 = contiguous ? (size_t)outer : (size_t)outer * p.n_inner + inner0;
        bool any = false;
        for (int t = 0; t < nlines; ++t) any |= p.line_flags[line0 + t] != 0;     // CTA-uniform
        if (!any) {
            if ((const void*)src != (const void*)dst) {
                if (contiguous) {
                    for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) dst[base + w] = mk<T>(0, 0);
                } else {
                    for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
                        int l = w / nlines, t = w - l * nlines;
                        dst[base + t + l * stride_l] = mk<T>(0, 0);
                    }
                }
            }
            return;
        }
    }

    // ---- load (coalesced along the contiguous direction), input shift folded in ----
    // full power-of-two tiles: every thread issues all of its loads before the first shared-memory
    // store, so TI*L/256 (16 at L = 256) requests per thread are in flight -- the passes are bound by
    // memory-level parallelism, not by bandwidth
    constexpr bool FAST = POW2 && (TI * (POW2 ? L : 1)) % 256 == 0 && (TI * (POW2 ? L : 1)) / 256 <= 32;
    bool fast = false;
    if constexpr (FAST) fast = nlines == TI && blockDim.x == 256;
    if constexpr (FAST) {
        if (fast) {
            constexpr int PER = TI * L / 256;
            cplx<T> v[PER];
            if (contiguous) {
#pragma unroll
                for (int k = 0; k < PER; ++k) v[k] = load_as_cplx<TIN, T>(src, base + threadIdx.x + 256 * k);   // lines are back to back
#pragma unroll
                for (int k = 0; k < PER; ++k) {
                    const int w = threadIdx.x + 256 * k;
                    const int t = w / L, l = w % L;
                    int i = l - p.in_shift; if (i < 0) i += L;
                    tile[(size_t)t * pitch + fft_pad<POW2>(i)] = v[k];
                }
            } else {
#pragma unroll
                for (int k = 0; k < PER; ++k) {
                    const int w = threadIdx.x + 256 * k;
                    const int l = w / TI, t = w % TI;
                    v[k] = load_as_cplx<TIN, T>(src, base + t + l * stride_l);
                }
#pragma unroll
                for (int k = 0; k < PER; ++k) {
                    const int w = threadIdx.x + 256 * k;
                    const int l = w / TI, t = w % TI;
                    int i = l - p.in_shift; if (i < 0) i += L;
                    tile[(size_t)t * pitch + fft_pad<POW2>(i)] = v[k];
                }
            }
        }
    }
    if (fast) {
    } else if (contiguous) {
        for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
            int t = w / Lr, l = w - t * Lr;
            int i = l - p.in_shift; if (i < 0) i += Lr;       // src index l lands at x_in[i]
            tile[(size_t)t * pitch + fft_pad<POW2>(i)] = load_as_cplx<TIN, T>(src, base + t * stride_t + l);
        }
    } else {
        for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
            int l = w / nlines, t = w - l * nlines;
            int i = l - p.in_shift; if (i < 0) i += Lr;
            tile[(size_t)t * pitch + fft_pad<POW2>(i)] = load_as_cplx<TIN, T>(src, base + t * stride_t + l * stride_l);
        }
    }
    __syncthreads();

    const T scale = (T)p.scale;
    if constexpr (POW2) {
        constexpr int r0 = radix_at(L, 0), r1 = radix_at(L, 1), r2 = radix_at(L, 2), ns = radix_count(L);
        smem_stage<DIR, T, L, L, r0, TI>(tile, pitch, nlines, tw);
        if constexpr (ns > 1) { __syncthreads(); smem_stage<DIR, T, L, L / r0, r1, TI>(tile, pitch, nlines, tw); }
        if constexpr (ns > 2) { __syncthreads(); smem_stage<DIR, T, L, L / r0 / r1, r2, TI>(tile, pitch, nlines, tw); }
        __syncthreads();
    } else {
        for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
            int t = w / Lr, o = w - t * Lr;
            const cplx<T>* x = tile + (size_t)t * pitch;
            cplx<T> acc = mk<T>(0, 0);
            int idx = 0;
            for (int i = 0; i < Lr; ++i) {
                cplx<T> w0 = tw[idx];
                acc = acc + (DIR > 0 ? cmul(x[i], w0) : cmulc(x[i], w0));
                idx += o; if (idx >= Lr) idx -= Lr;
            }
            tile2[(size_t)t * pitch + o] = acc;
        }
        __syncthreads();
    }
    // ---- store, output shift folded in ----
    const cplx<T>* res = POW2 ? tile : tile2;
    if constexpr (FAST) {
        if (fast) {
            constexpr int PER = TI * L / 256;
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int w = threadIdx.x + 256 * k;
                const int t = contiguous ? w / L : w % TI, l = contiguous ? w % L : w / TI;
                int o = l - p.out_shift; if (o < 0) o += L;
                const int pos = digit_reversed_pos<L>(o);
                const cplx<T> v = cscale(res[(size_t)t * pitch + fft_pad<POW2>(pos)], scale);
                if (p.herm_Z) {
                    const int line = (int)outer + t;                    // herm_Y is a power of two
                    const int kz = line >> p.herm_Yshift, ys = line & (p.herm_Y - 1), hz = p.herm_Z / 2;
                    const int zs = kz + hz == p.herm_Z ? 0 : kz + hz;
                    dst[((size_t)zs * p.herm_Y + ys) * L + l] = v;
                    if (kz > 0 && kz < hz) {
                        const int ym = ys ? p.herm_Y - ys : 0, lm = l ? L - l : 0;
                        dst[((size_t)(p.herm_Z - zs) * p.herm_Y + ym) * L + lm] = cconj(v);
                    }
                } else {
                    dst[base + t * stride_t + l * stride_l] = v;
                }
            }
            return;
        }
    }
    if (contiguous) {
        for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
            int t = w / Lr, l = w - t * Lr;                    // l = destination index
            int o = l - p.out_shift; if (o < 0) o += Lr;
            int pos = o;
            if constexpr (POW2) pos = digit_reversed_pos<L>(o);
            const cplx<T> v = cscale(res[(size_t)t * pitch + fft_pad<POW2>(pos)], scale);
            if (p.herm_Z) {
                const int line = (int)outer + t;                        // herm_Y is a power of two
                const int kz = line >> p.herm_Yshift, ys = line & (p.herm_Y - 1), hz = p.herm_Z / 2;
                const int zs = kz + hz == p.herm_Z ? 0 : kz + hz;
                dst[((size_t)zs * p.herm_Y + ys) * Lr + l] = v;
                if (kz > 0 && kz < hz) {
                    const int ym = ys ? p.herm_Y - ys : 0, lm = l ? Lr - l : 0;
                    dst[((size_t)(p.herm_Z - zs) * p.herm_Y + ym) * Lr + lm] = cconj(v);
                }
            } else {
                dst[base + t * stride_t + l] = v;
            }
        }
    } else {
        for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
            int l = w / nlines, t = w - l * nlines;
            int o = l - p.out_shift; if (o < 0) o += Lr;
            int pos = o;
            if constexpr (POW2) pos = digit_reversed_pos<L>(o);
            dst[base + t * stride_t + l * stride_l] = cscale(res[(size_t)t * pitch + fft_pad<POW2>(pos)], scale);
        }
    }
}

// |a|^2 summed over the leading axis: PSFi = sum_v |PSFa_v|^2  (pyotf/otf.py:204-207)
template <typename T>
__global__ void abs2_sum0_kernel(const cplx<T>* __restrict__ a, int nvec, size_t n, T* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T acc = 0;
    for (int v = 0; v < nvec; ++v) { cplx<T> x = a[(size_t)v * n + i]; acc += x.x * x.x + x.y * x.y; }
    out[i] = acc;
}

}  // namespace pyotf
