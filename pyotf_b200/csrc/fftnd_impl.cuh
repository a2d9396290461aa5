// Host launchers for the axis-pass FFT (instantiated per dtype in fftnd_f32.cu / fftnd_f64.cu).
#pragma once
#include <cstdlib>
#include <map>
#include <mutex>

#include "common.cuh"
#include "fftnd.cuh"
#include "fft_lines_reg.cuh"

namespace pyotf {

template <int DIR, typename TIN, typename T, int L, bool POW2, int TI>
int fft_axis_launch_ti(const AxisPass& p, const void* src, void* dst, size_t smem, cudaStream_t s) {
    auto kern = fft_axis_kernel<DIR, TIN, T, L, TI, POW2>;
    static thread_local size_t configured[PYOTF_MAX_DEVICES] = {};      // the attribute is per device
    const int dev_slot = current_device_slot();
    if (smem > configured[dev_slot]) {
        PYOTF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev_slot] = smem;
    }
    long long nblocks = p.n_inner == 1 ? (p.n_outer + TI - 1) / TI : p.n_outer * ((p.n_inner + TI - 1) / TI);
    PYOTF_REQUIRE(nblocks > 0 && nblocks < (1ll << 31), "fft: bad grid");
    kern<<<(unsigned)nblocks, 256, smem, s>>>(p, (const TIN*)src, (cplx<T>*)dst);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int DIR, typename TIN, typename T, int L, bool POW2>
int fft_axis_launch_l(const AxisPass& p, const void* src, void* dst, int smem_limit, cudaStream_t s) {
    // lines per tile: 16 keeps 128 B (fp32) / 256 B (fp64) global segments on strided axes;
    // long axes drop to 8 / 4 lines so the tile still fits (and two CTAs share an SM when possible)
    const int Lr = POW2 ? L : p.L;
    auto need = [&](int ti) { return sizeof(cplx<T>) * ((size_t)Lr + (size_t)ti * fft_pitch<POW2>(Lr) * (POW2 ? 1 : 2)); };
    const size_t soft = (size_t)smem_limit / 2 - 1024;
    if (need(16) <= soft || (need(8) > (size_t)smem_limit && need(16) <= (size_t)smem_limit))
        return fft_axis_launch_ti<DIR, TIN, T, L, POW2, 16>(p, src, dst, need(16), s);
    if (need(8) <= soft || (need(4) > (size_t)smem_limit && need(8) <= (size_t)smem_limit))
        return fft_axis_launch_ti<DIR, TIN, T, L, POW2, 8>(p, src, dst, need(8), s);
    PYOTF_REQUIRE(need(4) <= (size_t)smem_limit, "fft: axis too long for the shared-memory staged pass");
    return fft_axis_launch_ti<DIR, TIN, T, L, POW2, 4>(p, src, dst, need(4), s);
}

// the generic strided pass as a line functor of the register-resident kernel (fft_lines_reg.cuh):
// line = outer * n_inner + inner, element l of a line at ((outer * L + l) * n_inner + inner).
// LC = the (power-of-two) axis length at compile time; SMALL = the array has fewer than 2^31 elements: element offsets
// inside a line are 32-bit products and the shifts are masks (the index arithmetic was half of the pass's instructions).
template <typename TIN, typename T, int LC = 0, bool SMALL = false> struct AxisFn {
    struct Line { const TIN* in; cplx<T>* out; bool has; };
    using Acc = int;
    static constexpr int LOAD_BATCH = 8;
    static constexpr bool HAS_EMPTY = true;
    AxisPass p;
    const TIN* src;
    cplx<T>* dst;
    __device__ __forceinline__ static Acc acc_init() { return 0; }
    __device__ __forceinline__ void finish(Acc) const {}
    __device__ __forceinline__ bool skip() const { return false; }
    __device__ __forceinline__ cplx<T> twiddle(int k) const { return reinterpret_cast<const cplx<T>*>(p.tw)[k]; }
    __device__ __forceinline__ Line info(long long line) const {
        size_t base;
        if constexpr (SMALL) {
            const unsigned ln = (unsigned)line, ni = (unsigned)p.n_inner;
            const unsigned outer = ln / ni, inner = ln - outer * ni;
            base = (size_t)(outer * (unsigned)p.L * ni + inner);
        } else {
            const long long outer = line / p.n_inner, inner = line - outer * p.n_inner;
            base = (size_t)outer * p.L * p.n_inner + inner;
        }
        Line d;
        d.in = src + base; d.out = dst + base;
        d.has = p.line_flags ? p.line_flags[line] != 0 : true;
        return d;
    }
    __device__ __forceinline__ bool empty(const Line& d) const { return !d.has; }
    __device__ __forceinline__ void zero_fill(const Line& d, int o) const {
        if ((const void*)src != (const void*)dst) d.out[(size_t)o * p.n_inner] = mk<T>(0, 0);
    }
    __device__ __forceinline__ cplx<T> load(const Line& d, int i) const {
        if (!d.has) return mk<T>(0, 0);
        if constexpr (SMALL && LC > 0) {
            const unsigned l = (unsigned)(i + p.in_shift) & (unsigned)(LC - 1);
            return load_as_cplx<TIN, T>(d.in, l * (unsigned)p.n_inner);
        } else {
            int l = i + p.in_shift; if (l >= p.L) l -= p.L;              // x_in[i] = src[(i + in_shift) % L]
            return load_as_cplx<TIN, T>(d.in, (size_t)l * p.n_inner);
        }
    }
    __device__ __forceinline__ void store(const Line& d, int o, cplx<T> v, Acc&) const {
        if constexpr (SMALL && LC > 0) {
            const unsigned l = (unsigned)(o + p.out_shift) & (unsigned)(LC - 1);
            d.out[l * (unsigned)p.n_inner] = cscale(v, (T)p.scale);
        } else {
            int l = o + p.out_shift; if (l >= p.L) l -= p.L;             // dst[(o + out_shift) % L] = X[o]
            d.out[(size_t)l * p.n_inner] = cscale(v, (T)p.scale);
        }
    }
};

template <int DIR, typename TIN, typename T, int L, bool SMALL>
int fft_axis_reg_launch_s(const AxisPass& p, const void* src, void* dst, cudaStream_t s) {
    using F = AxisFn<TIN, T, L, SMALL>;
    F f;
    f.p = p; f.src = (const TIN*)src; f.dst = (cplx<T>*)dst;
    constexpr size_t smem = fft_lines_reg_smem<T, L>();
    auto kern = fft_lines_reg_kernel<DIR, T, L, F>;
    static thread_local bool configured[PYOTF_MAX_DEVICES] = {};
    const int dev_slot = current_device_slot();
    if (!configured[dev_slot]) {
        PYOTF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev_slot] = true;
    }
    const long long nlines = p.n_outer * p.n_inner, nblocks = nlines / RegPlan<L>::LANES;
    PYOTF_REQUIRE(nblocks > 0 && nblocks < (1ll << 31), "fft: bad grid");
    kern<<<(unsigned)nblocks, RegPlan<L>::NT, smem, s>>>(nlines, f);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int DIR, typename TIN, typename T, int L>
int fft_axis_reg_launch(const AxisPass& p, const void* src, void* dst, cudaStream_t s) {
    const bool small_arr = p.n_outer * (long long)p.L * p.n_inner < (1ll << 31);
    return small_arr ? fft_axis_reg_launch_s<DIR, TIN, T, L, true>(p, src, dst, s)
                     : fft_axis_reg_launch_s<DIR, TIN, T, L, false>(p, src, dst, s);
}

template <int DIR, typename TIN, typename T>
int fft_axis_launch(const AxisPass& p, const void* src, void* dst, int smem_limit, cudaStream_t s) {
    // strided axes: one lane per line, registers + one exchange (in place is safe: a block reads all of its lines
    // before it writes any)
    static const bool no_reg = getenv("PYOTF_B200_NO_REG_LINES") != nullptr;
    if (!no_reg && p.n_inner > 1) {
        if (p.L == 64 && p.n_inner % 32 == 0) return fft_axis_reg_launch<DIR, TIN, T, 64>(p, src, dst, s);
        if (p.L == 128 && p.n_inner % 32 == 0) return fft_axis_reg_launch<DIR, TIN, T, 128>(p, src, dst, s);
        if (p.L == 256 && p.n_inner % 32 == 0 && fft_lines_reg_smem<T, 256>() <= (size_t)smem_limit)
            return fft_axis_reg_launch<DIR, TIN, T, 256>(p, src, dst, s);
        if (p.L == 512 && p.n_inner % 16 == 0 && fft_lines_reg_smem<T, 512>() <= (size_t)smem_limit)
            return fft_axis_reg_launch<DIR, TIN, T, 512>(p, src, dst, s);
        if (p.L == 1024 && p.n_inner % 16 == 0 && fft_lines_reg_smem<T, 1024>() <= (size_t)smem_limit)
            return fft_axis_reg_launch<DIR, TIN, T, 1024>(p, src, dst, s);
    }
    switch (p.L) {
#define PYOTF_CASE(LL) case LL: return fft_axis_launch_l<DIR, TIN, T, LL, true>(p, src, dst, smem_limit, s);
        PYOTF_CASE(2) PYOTF_CASE(4) PYOTF_CASE(8) PYOTF_CASE(16) PYOTF_CASE(32) PYOTF_CASE(64) PYOTF_CASE(128)
        PYOTF_CASE(256) PYOTF_CASE(512) PYOTF_CASE(1024) PYOTF_CASE(2048)
#undef PYOTF_CASE
        default: return fft_axis_launch_l<DIR, TIN, T, 0, false>(p, src, dst, smem_limit, s);
    }
}

template <typename T> __global__ void fft_twiddle_kernel(int L, cplx<T>* __restrict__ tw) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L) return;
    double sn, cs;
    sincospi(2.0 * (double)k / (double)L, &sn, &cs);
    tw[k] = mk<T>((T)cs, (T)sn);
}

// e^{+2 pi i k / L} for k < L, computed in fp64 once per (device, length, dtype) and kept for the process
template <typename T> int fft_twiddles(int L, const void** out, cudaStream_t s) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, void*> cache;
    int dev = 0;
    PYOTF_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find({dev, L});
    if (it == cache.end()) {
        void* p = nullptr;
        PYOTF_CUDA_OK(cudaMalloc(&p, sizeof(cplx<T>) * (size_t)L));
        fft_twiddle_kernel<T><<<(L + 255) / 256, 256, 0, s>>>(L, (cplx<T>*)p);
        PYOTF_CUDA_OK(cudaGetLastError());
        PYOTF_CUDA_OK(cudaStreamSynchronize(s));      // other streams may use the table right away
        it = cache.emplace(std::make_pair(dev, L), p).first;
    }
    *out = it->second;
    return 0;
}

// ---- real 3D forward transform (OTFi = easy_fft(PSFi), pyotf/otf.py:209-212): first pass ---------------------------
// z-lines of a REAL [Z][Y][X] array -> the planes kz = 0 .. Z/2 of its z-transform (the rest is the complex
// conjugate mirror, filled in by the last pass).  One lane per PAIR of adjacent columns (x, x + 1): the two real lines
// travel as one complex line a + i b (one float2 load per sample), and are separated after the transform,
//     A[kz] = (C[kz] + conj C[-kz]) / 2,   B[kz] = (C[kz] - conj C[-kz]) / 2i,
// which needs C[kz] and C[Z - kz] in the same thread: second-stage unit c1 is executed together with unit R1 - c1
// (units 0 and R1/2 pair with themselves).  Same mapping as fft_lines_reg.cuh otherwise: a warp executes one unit
// for 32 column pairs, all first-stage loads in flight, one shared-memory exchange, 16-byte stores.
template <typename T, int Z>
__global__ void __launch_bounds__(16 * RegPlan<Z>::R1) rfft_z_half_kernel(const T* __restrict__ src, cplx<T>* __restrict__ dst,
                                                                          long long plane, int in_shift, const cplx<T>* __restrict__ twg) {
    using P = RegPlan<Z>;
    constexpr int R1 = P::R1, R2 = P::R2, NW = R1 / 2, NT = 32 * NW;
    using V2 = typename vec2<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);            // [Z]
    cplx<T>* E = tw + Z;                                            // [R1][R2][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t j = 2 * ((size_t)blockIdx.x * 32 + lane);          // first column of this lane's pair, within a plane
    for (int k = threadIdx.x; k < Z; k += NT) tw[k] = twg[k];
    __syncthreads();
    for (int g = warp; g < R2; g += NW) {
        cplx<T> x[R1];
#pragma unroll
        for (int q = 0; q < R1; ++q) {
            int zz = R2 * q + g + in_shift; if (zz >= Z) zz -= Z;   // x_in[i] = src[(i + in_shift) % Z]
            const V2 v = *reinterpret_cast<const V2*>(src + (size_t)zz * plane + j);
            x[q] = mk<T>(v.x, v.y);
        }
        fft_radix<-1, R1>(x);
        cplx<T>* e = E + (size_t)g * 32 + lane;
        e[0] = x[0];
#pragma unroll
        for (int c = 1; c < R1; ++c) e[(size_t)c * R2 * 32] = cmulc(x[c], tw[(g * c) & (Z - 1)]);
    }
    __syncthreads();
    // A / B at kz from C[kz] = c and C[-kz] = m
    auto put = [&](int kz, cplx<T> c, cplx<T> m) {
        if (kz > Z / 2) return;
        const T h = (T)0.5;
        const cplx<T> a = mk<T>((c.x + m.x) * h, (c.y - m.y) * h);
        const cplx<T> b = mk<T>((c.y + m.y) * h, (m.x - c.x) * h);
        cplx<T>* o = dst + (size_t)kz * plane + j;
        if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(o) = make_float4(a.x, a.y, b.x, b.y);
        else { o[0] = a; o[1] = b; }
    };
    if (warp == 0) {
        // units 0 and R1/2 are their own partners
        cplx<T> xa[R2];
#pragma unroll
        for (int g = 0; g < R2; ++g) xa[g] = E[(size_t)g * 32 + lane];
        fft_radix<-1, R2>(xa);
#pragma unroll
        for (int c2 = 0; c2 < R2; ++c2) put(R1 * c2, xa[c2], xa[(R2 - c2) % R2]);
#pragma unroll
        for (int g = 0; g < R2; ++g) xa[g] = E[(size_t)((R1 / 2) * R2 + g) * 32 + lane];
        fft_radix<-1, R2>(xa);
#pragma unroll
        for (int c2 = 0; c2 < R2; ++c2) put(R1 / 2 + R1 * c2, xa[c2], xa[R2 - 1 - c2]);
    } else {
        const int c1 = warp, c1m = R1 - warp;
        cplx<T> xa[R2], xb[R2];
#pragma unroll
        for (int g = 0; g < R2; ++g) { xa[g] = E[(size_t)(c1 * R2 + g) * 32 + lane]; xb[g] = E[(size_t)(c1m * R2 + g) * 32 + lane]; }
        fft_radix<-1, R2>(xa);
        fft_radix<-1, R2>(xb);
#pragma unroll
        for (int c2 = 0; c2 < R2; ++c2) {
            put(c1 + R1 * c2, xa[c2], xb[R2 - 1 - c2]);             // -kz = (R1 - c1) + R1 (R2 - 1 - c2)
            put(c1m + R1 * c2, xb[c2], xa[R2 - 1 - c2]);
        }
    }
}

template <typename T, int Z>
int rfft_z_half_launch(const void* src, void* dst, long long plane, int in_shift, const void* tw, cudaStream_t s) {
    constexpr size_t smem = sizeof(cplx<T>) * ((size_t)Z + (size_t)Z * 32);
    auto kern = rfft_z_half_kernel<T, Z>;
    static thread_local bool configured[PYOTF_MAX_DEVICES] = {};
    const int dev_slot = current_device_slot();
    if (!configured[dev_slot]) {
        PYOTF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev_slot] = true;
    }
    kern<<<(unsigned)(plane / 64), 16 * RegPlan<Z>::R1, smem, s>>>((const T*)src, (cplx<T>*)dst, plane, in_shift, (const cplx<T>*)tw);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- contiguous-axis pass in the register mapping (the x pass of the real 3D transform) -----------------------------
// A CTA brings LANES consecutive lines (rows of L complex values, contiguous in memory) into a shared-memory tile with an
// odd pitch (coalesced loads, all in flight), transforms them one LANE per line exactly like fft_lines_reg_kernel
// (radix-R1 units from the tile, one exchange, radix-R2 units back into the tile) and writes the tile out coalesced --
// with HERM every line goes to its fftshift-ed plane and, conjugated and reversed, to its Hermitian partner (AxisPass).
// 45 instead of 105 instructions per point of the staged generic kernel (whose index arithmetic was 58 % of its code).
// BULK: the tile is brought in by one bulk asynchronous copy per line (2 KB rows, signalled through an mbarrier) and
// the lines that go to their own position leave the same way; the tile pitch is then L + 2 (16-byte aligned rows:
// two-way bank conflicts on the per-lane accesses instead of none) instead of L + 1.
template <int DIR, typename T, int L, int LANES, bool HERM, bool BULK>
__global__ void __launch_bounds__(LANES * 16) cfft_rows_kernel(AxisPass p, const cplx<T>* __restrict__ src, cplx<T>* __restrict__ dst) {
    using P = RegPlan<L>;
    constexpr int R1 = P::R1, R2 = P::R2, NU = 16, NT = LANES * NU, TP = BULK ? L + 16 / (int)sizeof(cplx<T>) : L + 1, PER = LANES * L / NT;
    static_assert(R1 <= NU && R2 <= NU && (LANES * L) % NT == 0, "plan");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    cplx<T>* tw = reinterpret_cast<cplx<T>*>(smem_raw);            // [L]
    cplx<T>* Tt = tw + L;                                           // [LANES][TP]
    cplx<T>* E = Tt + (size_t)LANES * TP;                           // [R1][R2][LANES]
    const int tid = threadIdx.x, lane = tid % LANES, u = tid / LANES;
    const size_t row0 = (size_t)blockIdx.x * LANES;
    if constexpr (BULK) {
        if (tid == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0) mbar_expect_tx(&bar, (unsigned)(LANES * L * sizeof(cplx<T>)));
        if (tid < LANES) bulk_g2s(Tt + (size_t)tid * TP, src + (row0 + tid) * L, (unsigned)(L * sizeof(cplx<T>)), &bar);
        const cplx<T>* __restrict__ twg = reinterpret_cast<const cplx<T>*>(p.tw);
        for (int k = tid; k < L; k += NT) tw[k] = twg[k];
        __syncthreads();                                            // twiddles (and: the expect_tx precedes every wait)
        mbar_wait(&bar, 0);
    } else {
        const cplx<T>* __restrict__ s0 = src + row0 * L;
        cplx<T> v[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) v[k] = s0[k * NT + tid];
        const cplx<T>* __restrict__ twg = reinterpret_cast<const cplx<T>*>(p.tw);
        for (int k = tid; k < L; k += NT) tw[k] = twg[k];
#pragma unroll
        for (int k = 0; k < PER; ++k) { const int e = k * NT + tid; Tt[(e / L) * TP + (e & (L - 1))] = v[k]; }
        __syncthreads();
    }
    const cplx<T>* trow = Tt + (size_t)lane * TP;
    for (int g = u; g < R2; g += NU) {
        cplx<T> x[R1];
        const int gs = g + p.in_shift;
#pragma unroll
        for (int q = 0; q < R1; ++q) x[q] = trow[(R2 * q + gs) & (L - 1)];      // x_in[i] = src[(i + in_shift) % L]
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
    const T scale = (T)p.scale;
    for (int c1 = u; c1 < R1; c1 += NU) {
        cplx<T> x[R2];
        const cplx<T>* e = E + (size_t)c1 * R2 * LANES + lane;
#pragma unroll
        for (int g = 0; g < R2; ++g) x[g] = e[(size_t)g * LANES];
        fft_radix<DIR, R2>(x);
        cplx<T>* orow = Tt + (size_t)lane * TP;
        const int os = c1 + p.out_shift;
#pragma unroll
        for (int c2 = 0; c2 < R2; ++c2) orow[(os + R1 * c2) & (L - 1)] = cscale(x[c2], scale);   // dst[(o + out_shift) % L] = X[o]
    }
    if constexpr (BULK) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile is read by the copy engine next
    __syncthreads();
    // (the launcher checks that the output has fewer than 2^31 elements: 32-bit offsets)
    auto out_row = [&](unsigned line, unsigned& mirror_row) -> unsigned {
        if constexpr (HERM) {
            const unsigned kz = line >> p.herm_Yshift, ys = line & (unsigned)(p.herm_Y - 1), hz = (unsigned)p.herm_Z / 2;
            const unsigned zs = (kz + hz) & (unsigned)(p.herm_Z - 1);
            mirror_row = (kz > 0 && kz < hz) ? ((((unsigned)p.herm_Z - zs) << p.herm_Yshift) + ((0u - ys) & (unsigned)(p.herm_Y - 1))) : 0xffffffffu;
            return (zs << p.herm_Yshift) + ys;
        } else {
            mirror_row = 0xffffffffu;
            return line;
        }
    };
    if constexpr (BULK) {
        if (tid < LANES) {
            unsigned mr;
            const unsigned orow = out_row((unsigned)row0 + tid, mr);
            bulk_s2g(dst + (size_t)orow * L, Tt + (size_t)tid * TP, (unsigned)(L * sizeof(cplx<T>)));
        }
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int e = k * NT + tid, r = e / L, c = e & (L - 1);
        unsigned mr;
        const unsigned orow = out_row((unsigned)row0 + r, mr);
        const cplx<T> v = Tt[r * TP + c];
        if constexpr (!BULK) dst[orow * L + c] = v;
        if (HERM && mr != 0xffffffffu) dst[mr * L + ((0u - c) & (unsigned)(L - 1))] = cconj(v);
    }
    if constexpr (BULK) { if (tid < LANES) bulk_commit_and_wait_read(); }      // the tile must outlive the copy engine's reads
}

template <int DIR, typename T, int L, int LANES, bool HERM, bool BULK>
int cfft_rows_launch_b(const AxisPass& p, const void* src, void* dst, cudaStream_t s) {
    constexpr size_t smem = sizeof(cplx<T>) * ((size_t)L + (size_t)LANES * (L + 2) + (size_t)L * LANES);
    auto kern = cfft_rows_kernel<DIR, T, L, LANES, HERM, BULK>;
    static thread_local bool configured[PYOTF_MAX_DEVICES] = {};
    const int dev_slot = current_device_slot();
    if (!configured[dev_slot]) {
        PYOTF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev_slot] = true;
    }
    kern<<<(unsigned)(p.n_outer / LANES), LANES * 16, smem, s>>>(p, (const cplx<T>*)src, (cplx<T>*)dst);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int DIR, typename T, int L, int LANES, bool HERM>
int cfft_rows_launch(const AxisPass& p, const void* src, void* dst, cudaStream_t s) {
    // bulk asynchronous copies for the tile (measured 4 % slower on cfg1 than LDG / STG through registers: the 16-byte
    // aligned tile pitch costs two-way bank conflicts; kept selectable)
    static const bool bulk = getenv("PYOTF_B200_BULK_COPY") != nullptr;
    const bool aligned = ((size_t)src | (size_t)dst) % 16 == 0;
    return (bulk && aligned) ? cfft_rows_launch_b<DIR, T, L, LANES, HERM, true>(p, src, dst, s)
                                 : cfft_rows_launch_b<DIR, T, L, LANES, HERM, false>(p, src, dst, s);
}

template <typename T> int fft_twiddles(int L, const void** out, cudaStream_t s);
template <int DIR, typename TIN, typename T>
int fft_axis_launch(const AxisPass& p, const void* src, void* dst, int smem_limit, cudaStream_t s);

// fftshift(fftn(ifftshift(x))) of a real [Z][Y][X] array using the Hermitian symmetry of its spectrum: the z pass
// produces the planes kz = 0 .. Z/2 only (half-size scratch), the y pass runs in place on those, and the x pass writes
// every line twice -- to its fftshift-ed position and, conjugated and reversed, to its Hermitian partner.  HBM / L2
// traffic: 1 + 1 + 1 + 1 + 1 + 2 half-volumes (complex) instead of 0.5 + 2 + 2 + 2 + 2 + 2.
template <typename T>
int rfft3_shifted_run(const void* in, void* out, long long Z, long long Y, long long X, int smem_limit, cudaStream_t s, bool* done) {
    *done = false;
    static const bool off = getenv("PYOTF_B200_NO_RFFT3") != nullptr;      // A/B knob: the generic three full passes
    if (off || !(Z == 64 || Z == 128 || Z == 256) || !(Y == 64 || Y == 128 || Y == 256 || Y == 512 || Y == 1024) ||
        X % 32 != 0 || !is_pow2((int)X) || X < 64 || X > 1024 || (Y * X) % 64 != 0)
        return 0;
    if (fft_lines_reg_smem<T, 256>() > (size_t)smem_limit && Y >= 256) return 0;
    const long long plane = Y * X, hz = Z / 2 + 1;
    cplx<T>* W = nullptr;
    keep_scratch_pool();
    PYOTF_CUDA_OK(cudaMallocAsync(&W, sizeof(cplx<T>) * (size_t)hz * plane, s));
    const void *twz = nullptr, *twy = nullptr, *twx = nullptr;
    int rc = fft_twiddles<T>((int)Z, &twz, s);
    if (!rc) rc = fft_twiddles<T>((int)Y, &twy, s);
    if (!rc) rc = fft_twiddles<T>((int)X, &twx, s);
    if (!rc) {
        const int zshift = (int)((Z - Z / 2) % Z);
        if (Z == 64) rc = rfft_z_half_launch<T, 64>(in, W, plane, zshift, twz, s);
        else if (Z == 128) rc = rfft_z_half_launch<T, 128>(in, W, plane, zshift, twz, s);
        else rc = rfft_z_half_launch<T, 256>(in, W, plane, zshift, twz, s);
    }
    if (!rc) {
        AxisPass p;
        p.L = (int)Y; p.n_outer = hz; p.n_inner = X;
        p.in_shift = (int)((Y - Y / 2) % Y); p.out_shift = (int)(Y / 2); p.scale = 1.0; p.tw = twy; p.line_flags = nullptr;
        rc = fft_axis_launch<-1, cplx<T>, T>(p, W, W, smem_limit, s);
    }
    if (!rc) {
        AxisPass p;
        p.L = (int)X; p.n_outer = hz * Y; p.n_inner = 1;
        p.in_shift = (int)((X - X / 2) % X); p.out_shift = (int)(X / 2); p.scale = 1.0; p.tw = twx; p.line_flags = nullptr;
        p.herm_Z = (int)Z; p.herm_Y = (int)Y; p.herm_Yshift = 0;
        while ((1 << p.herm_Yshift) < (int)Y) ++p.herm_Yshift;
        constexpr int LN = sizeof(T) == 4 ? 16 : 8;             // lines per CTA: tile + exchange = 65 KB, three CTAs per SM
        static const bool generic_x = getenv("PYOTF_B200_RFFT3_GENERIC_X") != nullptr;      // A/B knob: the staged kernel
        const bool small_out = Z * Y * X < (1ll << 31) && is_pow2((int)Z);
        if (X == 256 && !generic_x && small_out) rc = cfft_rows_launch<-1, T, 256, LN, true>(p, W, out, s);
        else if (X == 128 && !generic_x && small_out) rc = cfft_rows_launch<-1, T, 128, LN, true>(p, W, out, s);
        else rc = fft_axis_launch<-1, cplx<T>, T>(p, W, out, smem_limit, s);
    }
    cudaFreeAsync(W, s);
    if (!rc) *done = true;
    return rc;
}

// out = [shift](fftn / ifftn)([shift] in) over `axes` of a C-contiguous array.
// pass_flags: optional, one pointer per pass in execution order (the LAST listed axis runs first): per-line
// "has data" bytes, see AxisPass::line_flags.
template <typename T>
int fftn_run(const void* in, void* out, int ndim, const long long* shape, const int* axes, int naxes, int inverse,
             int real_in, int shift_in, int shift_out, cudaStream_t s, const unsigned char* const* pass_flags) {
    int dev = 0, smem_limit = 0;
    PYOTF_CUDA_OK(cudaGetDevice(&dev));
    PYOTF_CUDA_OK(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (real_in && !inverse && ndim == 3 && naxes == 3 && shift_in && shift_out && !pass_flags) {
        bool done = false;
        int rc = rfft3_shifted_run<T>(in, out, shape[0], shape[1], shape[2], smem_limit, s, &done);
        if (rc || done) return rc;
    }
    bool first = true;
    // contiguous axis last-to-first keeps the real->complex pass on the fastest axis when present
    for (int k = naxes - 1; k >= 0; --k) {
        int ax = axes[k];
        AxisPass p;
        p.L = (int)shape[ax];
        p.n_outer = 1; p.n_inner = 1;
        for (int d = 0; d < ax; ++d) p.n_outer *= shape[d];
        for (int d = ax + 1; d < ndim; ++d) p.n_inner *= shape[d];
        int half = p.L / 2, rest = (p.L - half) % p.L;
        p.in_shift = shift_in ? (inverse ? rest : half) : 0;
        p.out_shift = shift_out ? (inverse ? rest : half) : 0;
        p.scale = inverse ? 1.0 / (double)p.L : 1.0;
        p.line_flags = pass_flags ? pass_flags[naxes - 1 - k] : nullptr;
        {
            int rc = fft_twiddles<T>(p.L, &p.tw, s);
            if (rc) return rc;
        }
        const void* src = first ? in : out;
        int rc;
        if (first && real_in) rc = inverse ? fft_axis_launch<1, T, T>(p, src, out, smem_limit, s)
                                           : fft_axis_launch<-1, T, T>(p, src, out, smem_limit, s);
        else rc = inverse ? fft_axis_launch<1, cplx<T>, T>(p, src, out, smem_limit, s)
                          : fft_axis_launch<-1, cplx<T>, T>(p, src, out, smem_limit, s);
        if (rc) return rc;
        first = false;
    }
    return 0;
}

template <typename T> int abs2_sum0_run(const void* a, int nvec, long long n, void* out, cudaStream_t s) {
    abs2_sum0_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const cplx<T>*)a, nvec, (size_t)n, (T*)out);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace pyotf
