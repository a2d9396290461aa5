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
    const cplx<T>* pupilc;    // [KR][K] compact pupil of this batch entry or nullptr
    const cplx<T>* tw;        // [N] e^{+2 pi i k / N} (the handle's table)
    const double* z;          // [nzc] (already offset to the chunk)
    cplx<T>* W;               // [nzc][K][N]
    cplx<T>* psfa;            // [nzc][N][N] of this (entry, component, chunk) or nullptr
    T* psfi;                  // [nzc][N][N] of this (entry, chunk) or nullptr
    int N, K, f0, nzc, accumulate;
    double scale;
};

// generic line FFT: the functor F maps a line index to a small descriptor once per line
// (F::Line info(line)), then loads / stores elements of that line:
//   F::load(const Line&, l) -> cplx<T>   (l = input index along the transformed axis, natural order)
//   F::store(const Line&, o, value)      (o = output index, natural order)
// TI lines per CTA.  LINE_MAJOR: consecutive threads walk along one line (the axis is contiguous in
// memory); otherwise consecutive threads take consecutive lines (the lines are contiguous).
template <int DIR, typename T, int L, int TI, bool POW2, bool LINE_MAJOR, class F>
__global__ void __launch_bounds__(256) fft_lines_kernel(int Lrt, long long nlines_total, F f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ typename F::Line sline[TI];
    if (f.skip()) return;                                     // e.g. phase retrieval already converged
    typename F::Acc acc = F::acc_init();                      // per-thread accumulator of the store functor
    const int Lr = POW2 ? L : Lrt;
    const int pitch = fft_pitch<POW2>(Lr);
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* tile = tw + Lr;
    cplx<T>* tile2 = tile + (size_t)TI * pitch;
    const long long line0 = (long long)blockIdx.x * TI;
    const int nlines = (int)min((long long)TI, nlines_total - line0);
    if (threadIdx.x < nlines) sline[threadIdx.x] = f.info(line0 + threadIdx.x);
    for (int k = threadIdx.x; k < Lr; k += blockDim.x) tw[k] = f.twiddle(k);   // e^{+2 pi i k / L}, precomputed
    __syncthreads();
    if constexpr (F::HAS_EMPTY) {                      // all lines of the tile known to be zero: no transform
        bool any = false;
        for (int t = 0; t < nlines; ++t) any |= !f.empty(sline[t]);     // CTA-uniform
        if (!any) {
            for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
                const int t = LINE_MAJOR ? w / Lr : w % nlines, o = LINE_MAJOR ? w % Lr : w / nlines;
                f.zero_fill(sline[t], o);
            }
            return;
        }
    }
    // full power-of-two tiles: a thread issues a batch of its loads before the first shared-memory store of the
    // batch (the element loads are generic-pointer loads the compiler will not move across shared stores), so up
    // to UB requests per thread are in flight instead of one -- the passes are bound by memory-level parallelism
    constexpr bool FAST = POW2 && (TI * (POW2 ? L : 1)) % 256 == 0;
    bool fast = false;
    if constexpr (FAST) fast = nlines == TI && blockDim.x == 256;
    if constexpr (FAST) {
        if (fast) {
            constexpr int PER = TI * L / 256, UB = PER < F::LOAD_BATCH ? PER : F::LOAD_BATCH;
            static_assert(PER % UB == 0, "load batch must divide the per-thread element count");
            for (int k0 = 0; k0 < PER; k0 += UB) {
                cplx<T> v[UB];
#pragma unroll
                for (int k = 0; k < UB; ++k) {
                    const int w = threadIdx.x + 256 * (k0 + k);
                    const int t = LINE_MAJOR ? w / L : w % TI, l = LINE_MAJOR ? w % L : w / TI;
                    v[k] = f.load(sline[t], l);
                }
#pragma unroll
                for (int k = 0; k < UB; ++k) {
                    const int w = threadIdx.x + 256 * (k0 + k);
                    const int t = LINE_MAJOR ? w / L : w % TI, l = LINE_MAJOR ? w % L : w / TI;
                    tile[(size_t)t * pitch + fft_pad<POW2>(l)] = v[k];
                }
            }
        }
    }
    if (fast) {
    } else if (LINE_MAJOR) {
        for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
            int t = w / Lr, l = w - t * Lr;
            tile[(size_t)t * pitch + fft_pad<POW2>(l)] = f.load(sline[t], l);
        }
    } else if (nlines == TI) {                       // full tile: compile-time divisor
        for (int w = threadIdx.x; w < TI * Lr; w += blockDim.x) {
            int l = w / TI, t = w - l * TI;
            tile[(size_t)t * pitch + fft_pad<POW2>(l)] = f.load(sline[t], l);
        }
    } else {
        for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
            int l = w / nlines, t = w - l * nlines;
            tile[(size_t)t * pitch + fft_pad<POW2>(l)] = f.load(sline[t], l);
        }
    }
    __syncthreads();
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
    const cplx<T>* res = POW2 ? tile : tile2;
    if (LINE_MAJOR) {
        for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
            int t = w / Lr, o = w - t * Lr;
            int pos = o;
            if constexpr (POW2) pos = digit_reversed_pos<L>(o);
            f.store(sline[t], o, res[(size_t)t * pitch + fft_pad<POW2>(pos)], acc);
        }
    } else if (nlines == TI) {
        for (int w = threadIdx.x; w < TI * Lr; w += blockDim.x) {
            int o = w / TI, t = w - o * TI;
            int pos = o;
            if constexpr (POW2) pos = digit_reversed_pos<L>(o);
            f.store(sline[t], o, res[(size_t)t * pitch + fft_pad<POW2>(pos)], acc);
        }
    } else {
        for (int w = threadIdx.x; w < nlines * Lr; w += blockDim.x) {
            int o = w / nlines, t = w - o * nlines;
            int pos = o;
            if constexpr (POW2) pos = digit_reversed_pos<L>(o);
            f.store(sline[t], o, res[(size_t)t * pitch + fft_pad<POW2>(pos)], acc);
        }
    }
    f.finish(acc);
}

// functors without a reduction / early exit
struct NoAcc {
    static constexpr int LOAD_BATCH = 8;              // element loads a thread keeps in flight (power of two)
    static constexpr bool HAS_EMPTY = false;          // functors that can prove a line is all-zero override this
    using Acc = int;
    __device__ __forceinline__ static Acc acc_init() { return 0; }
    __device__ __forceinline__ void finish(Acc) const {}
    __device__ __forceinline__ bool skip() const { return false; }
};

// ---- pass A: line = (plane, iy), axis = kx --------------------------------------------------
template <typename T> struct BigRow : NoAcc {
    struct Line { int iy; double z; cplx<T>* w; };
    BigArgs<T> a;
    __device__ __forceinline__ Line info(long long line) const {
        Line d;
        d.iy = (int)(line % a.K);
        d.z = a.z[(int)(line / a.K)];
        d.w = a.W + (size_t)line * a.N;
        return d;
    }
    __device__ __forceinline__ cplx<T> load(const Line& d, int l) const {
        const int f = l < (a.N + 1) / 2 ? l : l - a.N;
        const int ix = f - a.f0;
        if (ix < 0 || ix >= a.K) return mk<T>(0, 0);
        const int off = d.iy * a.K + ix;
        const T am = a.ampc[off];
        if (am == (T)0 && !a.pupilc) return mk<T>(0, 0);
        cplx<T> v;
        cis_cycles(a.kzc[off] * d.z, v);                         // _calc_defocus   otf.py:357-360
        v = cscale(v, am);
        if (a.pupilc) v = cmul(v, a.pupilc[off]);
        return v;
    }
    __device__ __forceinline__ void store(const Line& d, int o, cplx<T> v, int&) const { d.w[o] = v; }   // x stays unshifted
    __device__ __forceinline__ cplx<T> twiddle(int k) const { return a.tw[k]; }
};
// ---- pass B: line = (plane, x), axis = ky ---------------------------------------------------
template <typename T> struct BigCol : NoAcc {
    struct Line { const cplx<T>* w; cplx<T>* psfa; T* psfi; };
    BigArgs<T> a;
    __device__ __forceinline__ Line info(long long line) const {
        const int x = (int)(line % a.N), plane = (int)(line / a.N);
        const int xs = (x + a.N / 2) % a.N;                       // fftshift   otf.py:426
        const size_t base = (size_t)plane * a.N * a.N + xs;
        Line d;
        d.w = a.W + (size_t)plane * a.K * a.N + x;
        d.psfa = a.psfa ? a.psfa + base : nullptr;
        d.psfi = a.psfi ? a.psfi + base : nullptr;
        return d;
    }
    __device__ __forceinline__ cplx<T> load(const Line& d, int l) const {
        const int f = l < (a.N + 1) / 2 ? l : l - a.N;
        const int iy = f - a.f0;
        if (iy < 0 || iy >= a.K) return mk<T>(0, 0);
        return d.w[(size_t)iy * a.N];
    }
    __device__ __forceinline__ void store(const Line& d, int o, cplx<T> v, int&) const {
        int ys = o + a.N / 2; if (ys >= a.N) ys -= a.N;            // fftshift
        const T sc = (T)a.scale;
        v = cscale(v, sc);
        const size_t off = (size_t)ys * a.N;
        if (d.psfa) d.psfa[off] = v;
        if (d.psfi) {
            const T I = v.x * v.x + v.y * v.y;
            if (a.accumulate) d.psfi[off] += I; else d.psfi[off] = I;      // otf.py:204-207
        }
    }
    __device__ __forceinline__ cplx<T> twiddle(int k) const { return a.tw[k]; }
};

}  // namespace pyotf
