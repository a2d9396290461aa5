// K1b: HanserPSF._gen_psf (pyotf/otf.py:362-428) for sizes the single-CTA fused kernel K1 does
// not cover (N > 256, e.g. the 1024^2 planes of BASELINE config 4, or N not a power of two).
//
// Two passes per plane, both on the shared-memory line-FFT of fftnd.cuh:
//   pass A (rows)    : for each of the K pupil rows inside the support box, build
//                      pupil * amp * exp(2 pi i kz z) on the fly (nothing is read but the
//                      compact K x K tables) and transform along kx -> W[plane][iy][x]
//   pass B (columns) : for each x, gather the K non-zero ky entries of W, transform along
//                      ky, scale by 1/N^2, store fftshift-ed PSFa and/or accumulate |.|^2.
// W holds only K of N rows (43 % at the reference optics) and planes are processed in chunks
// sized so that W stays L2-resident; HBM sees the compact tables and the final PSFi/PSFa.
#pragma once
#include "fftnd.cuh"

namespace pyotf {

template <typename T> struct BigArgs {
    const double* kzc;        // [KR][K]
    const T* ampc;            // [KR][K] of the current polarisation component
    const cplx<T>* pupilc;    // [B][..] compact pupils or nullptr
    long long pupil_stride;
    const double* z;          // [nz] (already offset to the chunk)
    cplx<T>* W;               // [B][nzc][K][N]
    cplx<T>* psfa;            // [B][nvec][nz][N][N] offset to (component, chunk) or nullptr
    T* psfi;                  // [B][nz][N][N] offset to the chunk or nullptr
    long long psfa_bstride, psfi_bstride;   // elements between batch entries
    int N, K, f0, nzc, accumulate;
    double scale;
};

// generic line FFT: line index -> functor-defined loads / stores.  LINES are processed TI per CTA.
// LD(line, l) -> cplx<T> (l = input index along the transformed axis, natural order)
// ST(line, l, value)      (l = output index, natural order)
template <int DIR, typename T, int L, int TI, bool POW2, bool LINE_MAJOR, class LD, class ST>
__global__ void __launch_bounds__(256) fft_lines_kernel(int Lrt, long long nlines_total, LD ld, ST st) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Lr = POW2 ? L : Lrt;
    const int pitch = Lr + 1;
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* tile = tw + Lr;
    cplx<T>* tile2 = tile + (size_t)TI * pitch;
    for (int k = threadIdx.x; k < Lr; k += blockDim.x) {
        double s, c;
        sincospi(2.0 * (double)k / (double)Lr, &s, &c);
        tw[k] = mk<T>((T)c, (T)s);
    }
    const long long line0 = (long long)blockIdx.x * TI;
    const int nlines = (int)min((long long)TI, nlines_total - line0);
    // LINE_MAJOR: consecutive threads walk along one line (the axis is contiguous in memory);
    // otherwise consecutive threads take consecutive lines (the lines are contiguous).
    if (LINE_MAJOR) {
        for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
            int t = w / Lr, l = w - t * Lr;
            tile[(size_t)t * pitch + l] = ld(line0 + t, l);
        }
    } else {
        for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
            int l = w / nlines, t = w - l * nlines;
            tile[(size_t)t * pitch + l] = ld(line0 + t, l);
        }
    }
    __syncthreads();
    if constexpr (POW2) {
        constexpr int r0 = radix_at(L, 0), r1 = radix_at(L, 1), r2 = radix_at(L, 2), ns = radix_count(L);
        smem_stage<DIR, T, L, L, r0>(tile, pitch, nlines, tw);
        if constexpr (ns > 1) { __syncthreads(); smem_stage<DIR, T, L, L / r0, r1>(tile, pitch, nlines, tw); }
        if constexpr (ns > 2) { __syncthreads(); smem_stage<DIR, T, L, L / r0 / r1, r2>(tile, pitch, nlines, tw); }
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
    const cplx<T>* res = POW2 ? tile : tile2;
    if (LINE_MAJOR) {
        for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
            int t = w / Lr, o = w - t * Lr;
            int pos = o;
            if constexpr (POW2) pos = digit_reversed_pos<L>(o);
            st(line0 + t, o, res[(size_t)t * pitch + pos]);
        }
    } else {
        for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
            int o = w / nlines, t = w - o * nlines;
            int pos = o;
            if constexpr (POW2) pos = digit_reversed_pos<L>(o);
            st(line0 + t, o, res[(size_t)t * pitch + pos]);
        }
    }
}

// ---- pass A: line = (b, plane, iy), axis = kx ---------------------------------------------
template <typename T> struct BigRowLoad {
    BigArgs<T> a;
    __device__ __forceinline__ cplx<T> operator()(long long line, int l) const {
        const int iy = (int)(line % a.K);
        const long long bp = line / a.K;
        const int plane = (int)(bp % a.nzc), b = (int)(bp / a.nzc);
        const int f = l < (a.N + 1) / 2 ? l : l - a.N;
        const int ix = f - a.f0;
        if (ix < 0 || ix >= a.K) return mk<T>(0, 0);
        const int off = iy * a.K + ix;
        const T am = a.ampc[off];
        if (am == (T)0 && !a.pupilc) return mk<T>(0, 0);
        cplx<T> d;
        cis_cycles(a.kzc[off] * a.z[plane], d);                  // _calc_defocus   otf.py:357-360
        cplx<T> v = cscale(d, am);
        if (a.pupilc) v = cmul(v, a.pupilc[(size_t)b * a.pupil_stride + off]);
        return v;
    }
};
template <typename T> struct BigRowStore {
    BigArgs<T> a;
    __device__ __forceinline__ void operator()(long long line, int o, cplx<T> v) const {
        a.W[(size_t)line * a.N + o] = v;                           // x stays unshifted here
    }
};
// ---- pass B: line = (b, plane, x), axis = ky ----------------------------------------------
template <typename T> struct BigColLoad {
    BigArgs<T> a;
    __device__ __forceinline__ cplx<T> operator()(long long line, int l) const {
        const int x = (int)(line % a.N);
        const long long bp = line / a.N;
        const int f = l < (a.N + 1) / 2 ? l : l - a.N;
        const int iy = f - a.f0;
        if (iy < 0 || iy >= a.K) return mk<T>(0, 0);
        return a.W[((size_t)bp * a.K + iy) * a.N + x];
    }
};
template <typename T> struct BigColStore {
    BigArgs<T> a;
    __device__ __forceinline__ void operator()(long long line, int o, cplx<T> v) const {
        const int x = (int)(line % a.N);
        const long long bp = line / a.N;
        const int plane = (int)(bp % a.nzc), b = (int)(bp / a.nzc);
        const int xs = (x + a.N / 2) % a.N, ys = (o + a.N / 2) % a.N;     // fftshift   otf.py:426
        const T sc = (T)a.scale;
        v = cscale(v, sc);
        const size_t pix = ((size_t)plane * a.N + ys) * a.N + xs;
        if (a.psfa) a.psfa[(size_t)b * a.psfa_bstride + pix] = v;
        if (a.psfi) {
            T* p = a.psfi + (size_t)b * a.psfi_bstride + pix;
            const T I = v.x * v.x + v.y * v.y;
            if (a.accumulate) *p += I; else *p = I;                        // otf.py:204-207
        }
    }
};

}  // namespace pyotf
