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
// line = outer * n_inner + inner, element l of a line at ((outer * L + l) * n_inner + inner)
template <typename TIN, typename T> struct AxisFn {
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
        const long long outer = line / p.n_inner, inner = line - outer * p.n_inner;
        const size_t base = (size_t)outer * p.L * p.n_inner + inner;
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
        int l = i + p.in_shift; if (l >= p.L) l -= p.L;              // x_in[i] = src[(i + in_shift) % L]
        return load_as_cplx<TIN, T>(d.in, (size_t)l * p.n_inner);
    }
    __device__ __forceinline__ void store(const Line& d, int o, cplx<T> v, Acc&) const {
        int l = o + p.out_shift; if (l >= p.L) l -= p.L;             // dst[(o + out_shift) % L] = X[o]
        d.out[(size_t)l * p.n_inner] = cscale(v, (T)p.scale);
    }
};

template <int DIR, typename TIN, typename T, int L>
int fft_axis_reg_launch(const AxisPass& p, const void* src, void* dst, cudaStream_t s) {
    AxisFn<TIN, T> f;
    f.p = p; f.src = (const TIN*)src; f.dst = (cplx<T>*)dst;
    constexpr size_t smem = fft_lines_reg_smem<T, L>();
    auto kern = fft_lines_reg_kernel<DIR, T, L, AxisFn<TIN, T>>;
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

// out = [shift](fftn / ifftn)([shift] in) over `axes` of a C-contiguous array.
// pass_flags: optional, one pointer per pass in execution order (the LAST listed axis runs first): per-line
// "has data" bytes, see AxisPass::line_flags.
template <typename T>
int fftn_run(const void* in, void* out, int ndim, const long long* shape, const int* axes, int naxes, int inverse,
             int real_in, int shift_in, int shift_out, cudaStream_t s, const unsigned char* const* pass_flags) {
    int dev = 0, smem_limit = 0;
    PYOTF_CUDA_OK(cudaGetDevice(&dev));
    PYOTF_CUDA_OK(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
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
