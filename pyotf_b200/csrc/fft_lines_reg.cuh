// Register-resident line FFT for lines that are CONTIGUOUS ACROSS lines (the transformed axis is strided: the y / z
// passes of the 3D transforms, pass B of K1b, passes B / C of K2b).  Same functor interface as fft_lines_kernel
// (hanser_big.cuh), different mapping: one LANE per line, so every element access of a warp is a coalesced run of
// LANES consecutive lines, and a warp executes one radix unit for all of its lines:
//   X[c1 + R1 c2] = sum_g w_R2^{g c2} [ w_L^{g c1} sum_q x[R2 q + g] w_R1^{q c1} ],   L = R1 * R2
//   stage 1: unit g  (R2 of them): R1 loads straight from global memory (all in flight), radix-R1 in registers,
//            twiddle, one shared-memory store per value       E[c1][g][lane]
//   stage 2: unit c1 (R1 of them): R2 loads from E, radix-R2 in registers, the functor's store (coalesced).
// No staging tile, no digit reversal, one barrier; per element one global load, one global store, one 8-byte
// shared-memory write and read.  Inputs the functor knows to be zero (band limits, empty planes) cost no load.
#pragma once
#include "fft_regs.cuh"

namespace pyotf {

// NT threads per CTA: 256 (several CTAs per SM) or, where the exchange buffer allows one CTA per SM only, 512
template <int L> struct RegPlan;
template <> struct RegPlan<64>   { static constexpr int R1 = 8,  R2 = 8,  LANES = 32, NT = 256; };
template <> struct RegPlan<128>  { static constexpr int R1 = 16, R2 = 8,  LANES = 32, NT = 256; };
template <> struct RegPlan<256>  { static constexpr int R1 = 16, R2 = 16, LANES = 32, NT = 256; };
template <> struct RegPlan<512>  { static constexpr int R1 = 32, R2 = 16, LANES = 16, NT = 256; };
template <> struct RegPlan<1024> { static constexpr int R1 = 32, R2 = 32, LANES = 16, NT = 512; };

template <typename T, int L> constexpr size_t fft_lines_reg_smem() {
    return sizeof(cplx<T>) * ((size_t)L + (size_t)L * RegPlan<L>::LANES);
}

template <int DIR, typename T, int L, class F>
__global__ void __launch_bounds__(RegPlan<L>::NT) fft_lines_reg_kernel(long long nlines_total, F f) {
    using P = RegPlan<L>;
    constexpr int R1 = P::R1, R2 = P::R2, LANES = P::LANES, UPW = 32 / LANES, NT = P::NT, NU = NT / 32 * UPW;   // units in flight per CTA
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (f.skip()) return;
    typename F::Acc acc = F::acc_init();
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);            // [L] e^{+2 pi i k / L}
    cplx<T>* E = tw + L;                                            // [R1][R2][LANES]
    const int lane = threadIdx.x & (LANES - 1);
    const int u0 = (threadIdx.x >> 5) * UPW + ((threadIdx.x & 31) / LANES);     // first unit of this thread
    const long long line = (long long)blockIdx.x * LANES + lane;
    const typename F::Line d = f.info(line);
    for (int k = threadIdx.x; k < L; k += NT) tw[k] = f.twiddle(k);
    if constexpr (F::HAS_EMPTY) {
        if (!__syncthreads_or(!f.empty(d))) {                      // every line of the block is zero: no transform
            for (int o = u0; o < L; o += NU) f.zero_fill(d, o);
            return;
        }
    } else {
        __syncthreads();
    }
    for (int g = u0; g < R2; g += NU) {
        cplx<T> x[R1];
#pragma unroll
        for (int q = 0; q < R1; ++q) x[q] = f.load(d, R2 * q + g);
        fft_radix<DIR, R1>(x);
        cplx<T>* e = E + (size_t)g * LANES + lane;
        e[0] = x[0];
#pragma unroll
        for (int c = 1; c < R1; ++c) {
            const cplx<T> w = tw[(g * c) & (L - 1)];
            e[(size_t)c * R2 * LANES] = DIR > 0 ? cmul(x[c], w) : cmulc(x[c], w);
        }
    }
    __syncthreads();
    for (int c1 = u0; c1 < R1; c1 += NU) {
        cplx<T> x[R2];
        const cplx<T>* e = E + (size_t)c1 * R2 * LANES + lane;
#pragma unroll
        for (int g = 0; g < R2; ++g) x[g] = e[(size_t)g * LANES];
        fft_radix<DIR, R2>(x);
#pragma unroll
        for (int c2 = 0; c2 < R2; ++c2) f.store(d, c1 + R1 * c2, x[c2], acc);
    }
    f.finish(acc);
}

}  // namespace pyotf
