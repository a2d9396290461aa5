// Per-dtype host launchers for K1 (instantiated in hanser_f32.cu / hanser_f64.cu).
#pragma once
#include <cstring>
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "hanser.cuh"
#include "hanser_big.cuh"
#include "fft_lines_reg.cuh"

namespace pyotf {

template <typename T> int hanser_build_tables(pyotf_hanser* h, cudaStream_t s) {
    HanserGeomParams g;
    g.N = h->N; g.K = h->K; g.KR = h->KR; g.f0 = h->f0; g.nvec = h->nvec;
    g.vec_mode = h->p.vec_corr; g.condition = h->p.condition;
    g.apply_mask = h->p.support < 0 ? 1 : 0;
    g.kstep = 1.0 / ((double)h->N * h->p.res);      // numpy.fft.fftfreq: val = 1.0 / (n * d)
    g.kmag = h->p.ni / h->p.wl;                      // otf.py:341
    g.diff_limit = h->p.na / h->p.wl;                // otf.py:352
    int KK = h->KR * h->K;
    hanser_geometry_kernel<T><<<(KK + 255) / 256, 256, 0, s>>>(g, h->kzc, (T*)h->ampc, h->maskc);
    PYOTF_CUDA_OK(cudaGetLastError());
    twiddle_kernel<T><<<(h->N + 255) / 256, 256, 0, s>>>(h->N, (cplx<T>*)h->tw);
    PYOTF_CUDA_OK(cudaGetLastError());
    if (h->K == 2 * (-h->f0) + 1 && h->K < h->N) {      // symmetric support box: octant lists for the table builders
        const int H = -h->f0, ntot = (H / 2 + 1) * (H + 2);
        PYOTF_CUDA_OK(cudaMalloc(&h->octkz, sizeof(double) * ntot));
        PYOTF_CUDA_OK(cudaMalloc(&h->octamp, sizeof(T) * ntot));
        PYOTF_CUDA_OK(cudaMalloc(&h->octij, sizeof(int) * ntot));
        hanser_oct_kernel<T><<<(ntot + 255) / 256, 256, 0, s>>>(h->kzc, (const T*)h->ampc, h->f0, h->K, h->octkz,
                                                              (T*)h->octamp, h->octij);
        PYOTF_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

template <typename T>
int hanser_pack(const pyotf_hanser* h, const void* src, int batch, void* dst, cudaStream_t s) {
    long long tot = (long long)h->KR * h->K * batch;
    pack_pupil_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(h->N, h->K, h->KR, h->f0, (const double2*)src,
                                                                      (cplx<T>*)dst, batch);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

// entries of one per-plane defocus table when the support box is symmetric about DC and the table
// fits the kernels' exchange region (where it is staged), else 0 (defocus evaluated on the fly)
template <typename T> int hanser_ntab(const pyotf_hanser* h, int nthreads = 256) {
    if (h->K != 2 * (-h->f0) + 1) return 0;
    int cap = 0;
    switch (h->N) {       // EXTOT does not depend on M
        case 32:  cap = nthreads == 512 ? HanserSmem<32, 16, T, 512>::EXTOT : HanserSmem<32, 16, T>::EXTOT; break;
        case 64:  cap = nthreads == 512 ? HanserSmem<64, 16, T, 512>::EXTOT : HanserSmem<64, 16, T>::EXTOT; break;
        case 128: cap = nthreads == 512 ? HanserSmem<128, 16, T, 512>::EXTOT : HanserSmem<128, 16, T>::EXTOT; break;
        case 256: cap = nthreads == 512 ? HanserSmem<256, 16, T, 512>::EXTOT : HanserSmem<256, 16, T>::EXTOT; break;
        default: return 0;
    }
    return dtab_size(h->f0) <= cap ? dtab_size(h->f0) : 0;
}

// (re)build the per-plane defocus tables of a stack into `dst` ([nz][ntab])
template <typename T>
int hanser_build_dtabs(const pyotf_hanser* h, const double* z, int nz, bool with_amp, cplx<T>* dst, cudaStream_t s) {
    const int H = -h->f0, ntot = (H / 2 + 1) * (H + 2);
    PYOTF_REQUIRE(nz <= 65535, "hanser_psf: too many planes for one launch");
    dim3 grid((unsigned)((ntot + 255) / 256), (unsigned)nz);
    hanser_dtab_kernel<T><<<grid, 256, 0, s>>>(h->kzc, with_amp ? (const T*)h->ampc : nullptr, z, h->f0, h->K, dst);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T, int N, int M, int TABM, bool SYM = false>
int hanser_launch_nmt(const HanserArgs<T>& a, int batch, size_t smem, cudaStream_t s) {
    using SM = HanserSmem<N, M, T>;
    auto kern = hanser_psf_kernel<T, N, M, TABM, SYM>;
    static thread_local size_t configured[PYOTF_MAX_DEVICES] = {};      // the attribute is per device
    const int dev_slot = current_device_slot();
    if (smem > configured[dev_slot]) {
        PYOTF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev_slot] = smem;
    }
    dim3 grid((unsigned)((N / M) * a.nz), (unsigned)batch);
    if constexpr (TABM == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(SM::NT); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = N / M; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        PYOTF_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, a));
    } else if (TABM == 3 && a.pupilc == nullptr) {
        // programmatic dependent launch: back-to-back stacks overlap.  Default: the kernel waits for its
        // predecessor before it reads anything, so only the launch latency is hidden and no assumption is made.
        // Streaming mode (pyotf_b200_hanser_set_overlap): the caller promises that nothing this launch reads
        // (z_dev; the handle's tables are immutable) is written by the kernel that precedes it in the stream; the
        // kernel then waits only right before its first global store, so the whole compute phase overlaps too.
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(SM::NT); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        PYOTF_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, a));
    } else {
        kern<<<grid, SM::NT, smem, s>>>(a);
    }
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

// pitch of the compact work area of the even-symmetric variant (columns fx = 0 .. -f0 only), same residue
// rule as the full pitch (cabi.cu)
inline int hanser_sym_pitch(int f0) { const int nc = 1 - f0; return nc + ((8 - nc % 16) + 16) % 16; }

template <typename T> bool hanser_use_sym(const HanserArgs<T>& a) {
    // ideal pupil of the scalar model: the even-symmetric variant (half the columns, half the rows)
    static const bool no_sym = getenv("PYOTF_B200_K1_NOSYM") != nullptr;
    return a.amp_in_tab && !no_sym;
}

// 128 rows per CTA (256^2 only, even-symmetric variant only: its compact work area lets two CTAs share an SM):
// the fold and the per-CTA defocus table are paid twice per plane instead of four times
template <typename T, int N>
int hanser_launch_sym128(const HanserArgs<T>& a, int tabm, int batch, cudaStream_t s) {
    const size_t smem = HanserSmem<N, 128, T>::bytes(a.KP);
    if (tabm == 3) return hanser_launch_nmt<T, N, 128, 3, true>(a, batch, smem, s);
    return hanser_launch_nmt<T, N, 128, 1, true>(a, batch, smem, s);
}

template <typename T, int N, int M>
int hanser_launch_nm(const pyotf_hanser* h, const HanserArgs<T>& a, int tabm, int batch, cudaStream_t s) {
    size_t smem = HanserSmem<N, M, T>::bytes(a.KP);
#ifdef PYOTF_ENABLE_CLUSTER_TABLE   // experiment kept for the record (measured slower): -DPYOTF_ENABLE_CLUSTER_TABLE
    if (tabm == 2) {
        if constexpr (N / M <= 8) return hanser_launch_nmt<T, N, M, 2>(a, batch, smem, s);
    }
#endif
    const bool sym = hanser_use_sym(a);
    if (tabm == 3) return sym ? hanser_launch_nmt<T, N, M, 3, true>(a, batch, smem, s) : hanser_launch_nmt<T, N, M, 3>(a, batch, smem, s);
    if (tabm == 1) return sym ? hanser_launch_nmt<T, N, M, 1, true>(a, batch, smem, s) : hanser_launch_nmt<T, N, M, 1>(a, batch, smem, s);
    return hanser_launch_nmt<T, N, M, 0>(a, batch, smem, s);
}

template <typename T, int N> size_t hanser_smem_for(int M, int KP) {
    if (M == 16) return N <= 64 ? HanserSmem<N, 16, T>::bytes(KP) : (size_t)-1;
    if (M == 32) return HanserSmem<N, (N >= 32 ? 32 : 16), T>::bytes(KP);
    return HanserSmem<N, (N >= 64 ? 64 : 16), T>::bytes(KP);
}

// picks rows-per-CTA M; *m_out receives it when only the choice is wanted (launch = false)
template <typename T, int N>
int hanser_launch_n(const pyotf_hanser* h, const HanserArgs<T>& a_in, int tabm, int batch, int smem_limit, cudaStream_t s,
                    int* m_out = nullptr) {
    const HanserArgs<T>& a = a_in;
    // largest rows-per-CTA M whose working set allows two CTAs per SM, else the largest that fits
    int pick = 0;
    const int cand[3] = {64, 32, 16};
    static const int forced = [] { const char* e = getenv("PYOTF_B200_K1_ROWS"); return e ? atoi(e) : 0; }();   // tuning knob
    if (forced && forced <= N && forced != 128 && (forced != 16 || N <= 64) && hanser_smem_for<T, N>(forced, a.KP) <= (size_t)smem_limit) pick = forced;
    // 64 or 32 rows per CTA, two CTAs per SM if the working set allows, else one; 16 rows only as a last
    // resort (its radix-16 fold spills registers badly: measured 20x slower)
    for (int pass = 0; pass < 3 && !pick; ++pass)
        for (int i = 0; i < 3 && !pick; ++i) {
            int M = cand[i];
            if (M > N || (M == 16 && (pass < 2 || N > 64))) continue;   // 16 rows: small planes only, last resort
            size_t need = hanser_smem_for<T, N>(M, a.KP);
            size_t lim = pass == 0 ? (size_t)(smem_limit / 2 - 1024) : (size_t)smem_limit;
            if (need <= lim) pick = M;
        }
    PYOTF_REQUIRE(pick != 0, "hanser_psf: pupil support too large for the fused on-chip path at this size/dtype");
    // even-symmetric variant: compact work area (the choice above is made on the full pitch, as before)
    const bool sym = hanser_use_sym(a);
    HanserArgs<T> as = a;
    if (sym) as.KP = hanser_sym_pitch(a.f0);
    if constexpr (N == 256) {
        // 128 rows per CTA when as many such CTAs share an SM as 64-row ones would (two in fp32 mode, one in fp64
        // mode) and the 64-row grid would not fit one wave anyway
        const size_t half = (size_t)(smem_limit / 2 - 1024);
        const size_t need128 = HanserSmem<N, 128, T>::bytes(as.KP), need64 = HanserSmem<N, 64, T>::bytes(as.KP);
        const int per_sm64 = need64 <= half ? 2 : 1, per_sm128 = need128 <= half ? 2 : (need128 <= (size_t)smem_limit ? 1 : 0);
        const bool fits = sym && per_sm128 >= per_sm64;
        int sms = 0;
        PYOTF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
        const long long ctas64 = (long long)(N / 64) * a.nz * batch;
        if (fits && (forced == 128 || (!forced && pick == 64 && ctas64 > (long long)per_sm64 * sms))) pick = 128;
    }
    if (m_out) { *m_out = pick; return 0; }
    if (pick == 128) { if constexpr (N == 256) return hanser_launch_sym128<T, N>(as, tabm, batch, s); }
    if (sym) {
        if (pick == 64) { if constexpr (N >= 64) return hanser_launch_nm<T, N, 64>(h, as, tabm, batch, s); }
        if (pick == 32) { if constexpr (N >= 32) return hanser_launch_nm<T, N, 32>(h, as, tabm, batch, s); }
        if constexpr (N <= 64) return hanser_launch_nm<T, N, 16>(h, as, tabm, batch, s);
    }
    if (pick == 64) { if constexpr (N >= 64) return hanser_launch_nm<T, N, 64>(h, a, tabm, batch, s); }
    if (pick == 32) { if constexpr (N >= 32) return hanser_launch_nm<T, N, 32>(h, a, tabm, batch, s); }
    if constexpr (N <= 64) return hanser_launch_nm<T, N, 16>(h, a, tabm, batch, s);
    set_error("hanser_psf: internal error (rows per CTA)");
    return 1;
}

// K1s (hanser_sym.cuh, its own translation units hanser_sym_f32.cu / _f64.cu): one CTA per plane, scalar model with
// the ideal pupil at N = 256, PSFi only.  *done stays false when the configuration does not fit (caller falls back).
template <typename T> int hanser_sym_launch(const HanserArgs<T>& a, int batch, int smem_limit, cudaStream_t s, bool* done);

// K1bs (hanser_bigsym.cuh, its own translation units): the same model at N = 1024 (BASELINE config 4), two passes
// over an L2-resident half-spectrum intermediate.  *done stays false when the configuration does not fit.
template <typename T>
int hanser_bigsym_launch(const pyotf_hanser* h, const double* z, int nz, int batch, void* psfi, int smem_limit, cudaStream_t s,
                         bool* done);
// K1br: any pupil at N = 1024, PSFi only (same file)
template <typename T>
int hanser_biggen_launch(const pyotf_hanser* h, const double* z, int nz, const void* pupilc, int batch, long long pupil_stride,
                         void* psfi, int smem_limit, cudaStream_t s, bool* done);

// ---------------------------------------------------------------------------------------
// K1b launchers (hanser_big.cuh): any size, two line-FFT passes per plane chunk
// ---------------------------------------------------------------------------------------
template <int DIR, typename T, int L, bool POW2, bool LM, int TI, class F>
int fft_lines_launch_ti(int Lr, long long nlines, F f, size_t smem, cudaStream_t s, int* nblocks_out) {
    auto kern = fft_lines_kernel<DIR, T, L, TI, POW2, LM, F>;
    static thread_local size_t configured[PYOTF_MAX_DEVICES] = {};      // the attribute is per device
    const int dev_slot = current_device_slot();
    if (smem > configured[dev_slot]) {
        PYOTF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev_slot] = smem;
    }
    long long nblocks = (nlines + TI - 1) / TI;
    PYOTF_REQUIRE(nblocks > 0 && nblocks < (1ll << 31), "hanser_psf: bad grid");
    if (nblocks_out) *nblocks_out = (int)nblocks;
    kern<<<(unsigned)nblocks, 256, smem, s>>>(Lr, nlines, f);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int DIR, typename T, int L, bool POW2, bool LM, class F>
int fft_lines_launch_l(int Lr, long long nlines, F f, int smem_limit, cudaStream_t s, int* nblocks_out) {
    auto need = [&](int ti) { return sizeof(cplx<T>) * ((size_t)Lr + (size_t)ti * fft_pitch<POW2>(Lr) * (POW2 ? 1 : 2)); };
    const size_t soft = (size_t)smem_limit / 2 - 1024;
    if (need(16) <= soft || (need(8) > (size_t)smem_limit && need(16) <= (size_t)smem_limit))
        return fft_lines_launch_ti<DIR, T, L, POW2, LM, 16>(Lr, nlines, f, need(16), s, nblocks_out);
    if (need(8) <= soft || (need(4) > (size_t)smem_limit && need(8) <= (size_t)smem_limit))
        return fft_lines_launch_ti<DIR, T, L, POW2, LM, 8>(Lr, nlines, f, need(8), s, nblocks_out);
    PYOTF_REQUIRE(need(4) <= (size_t)smem_limit, "hanser_psf: size too large for the shared-memory line FFT");
    return fft_lines_launch_ti<DIR, T, L, POW2, LM, 4>(Lr, nlines, f, need(4), s, nblocks_out);
}

// strided lines (the axis is not contiguous): the register-resident kernel of fft_lines_reg.cuh where a plan exists
template <int DIR, typename T, int L, class F>
int fft_lines_reg_launch(long long nlines, F f, cudaStream_t s, int* nblocks_out) {
    constexpr size_t smem = fft_lines_reg_smem<T, L>();
    auto kern = fft_lines_reg_kernel<DIR, T, L, F>;
    static thread_local bool configured[PYOTF_MAX_DEVICES] = {};
    const int dev_slot = current_device_slot();
    if (!configured[dev_slot]) {
        PYOTF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev_slot] = true;
    }
    const long long nblocks = nlines / RegPlan<L>::LANES;
    PYOTF_REQUIRE(nblocks > 0 && nblocks < (1ll << 31), "fft: bad grid");
    if (nblocks_out) *nblocks_out = (int)nblocks;
    kern<<<(unsigned)nblocks, RegPlan<L>::NT, smem, s>>>(nlines, f);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int DIR, typename T, bool LM, class F>
int fft_lines_launch(int L, long long nlines, F f, int smem_limit, cudaStream_t s, int* nblocks_out = nullptr) {
    if constexpr (!LM) {
        static const bool no_reg = getenv("PYOTF_B200_NO_REG_LINES") != nullptr;     // A/B knob: the staged kernel
        if (!no_reg) {
            if (L == 64 && nlines % 32 == 0) return fft_lines_reg_launch<DIR, T, 64>(nlines, f, s, nblocks_out);
            if (L == 128 && nlines % 32 == 0) return fft_lines_reg_launch<DIR, T, 128>(nlines, f, s, nblocks_out);
            if (L == 256 && nlines % 32 == 0 && fft_lines_reg_smem<T, 256>() <= (size_t)smem_limit)
                return fft_lines_reg_launch<DIR, T, 256>(nlines, f, s, nblocks_out);
            if (L == 512 && nlines % 16 == 0 && fft_lines_reg_smem<T, 512>() <= (size_t)smem_limit)
                return fft_lines_reg_launch<DIR, T, 512>(nlines, f, s, nblocks_out);
            if (L == 1024 && nlines % 16 == 0 && fft_lines_reg_smem<T, 1024>() <= (size_t)smem_limit)
                return fft_lines_reg_launch<DIR, T, 1024>(nlines, f, s, nblocks_out);
        }
    }
    switch (L) {
#define PYOTF_CASE(LL) case LL: return fft_lines_launch_l<DIR, T, LL, true, LM>(L, nlines, f, smem_limit, s, nblocks_out);
        PYOTF_CASE(2) PYOTF_CASE(4) PYOTF_CASE(8) PYOTF_CASE(16) PYOTF_CASE(32) PYOTF_CASE(64) PYOTF_CASE(128)
        PYOTF_CASE(256) PYOTF_CASE(512) PYOTF_CASE(1024) PYOTF_CASE(2048)
#undef PYOTF_CASE
        default: return fft_lines_launch_l<DIR, T, 0, false, LM>(L, nlines, f, smem_limit, s, nblocks_out);
    }
}

template <typename T>
int hanser_launch_big(const pyotf_hanser* h, const double* z, int nz, const void* pupilc, int batch,
                      long long pupil_stride, void* psfa, void* psfi, int smem_limit, cudaStream_t s) {
    const int N = h->N, K = h->K;
    const size_t plane = (size_t)N * N, KKP = (size_t)h->KR * K;
    // planes per chunk: keep the K x N intermediate of a chunk L2-resident (<= 48 MiB)
    const size_t wplane = (size_t)K * N * sizeof(cplx<T>);
    int nzc = (int)std::max<size_t>(1, std::min<size_t>((size_t)nz, ((size_t)48 << 20) / wplane));
    cplx<T>* W = nullptr;
    keep_scratch_pool();
    PYOTF_CUDA_OK(cudaMallocAsync(&W, wplane * nzc, s));
    int rc = 0;
    for (int b = 0; b < batch && !rc; ++b)
        for (int z0 = 0; z0 < nz && !rc; z0 += nzc) {
            const int nc = std::min(nzc, nz - z0);
            for (int v = 0; v < h->nvec && !rc; ++v) {
                BigArgs<T> a;
                a.kzc = h->kzc; a.ampc = (const T*)h->ampc + (size_t)v * KKP;
                a.pupilc = pupilc ? (const cplx<T>*)pupilc + (size_t)b * pupil_stride : nullptr;
                a.z = z + z0; a.W = W; a.tw = (const cplx<T>*)h->tw;
                a.psfa = psfa ? (cplx<T>*)psfa + (((size_t)b * h->nvec + v) * nz + z0) * plane : nullptr;
                a.psfi = psfi ? (T*)psfi + ((size_t)b * nz + z0) * plane : nullptr;
                a.N = N; a.K = K; a.f0 = h->f0; a.nzc = nc; a.accumulate = v > 0;
                a.scale = 1.0 / ((double)N * (double)N);           // numpy ifftn "backward" normalisation
                { BigRow<T> fa; fa.a = a; rc = fft_lines_launch<1, T, true>(N, (long long)nc * K, fa, smem_limit, s); }
                if (!rc) { BigCol<T> fb; fb.a = a; rc = fft_lines_launch<1, T, false>(N, (long long)nc * N, fb, smem_limit, s); }
            }
        }
    cudaFreeAsync(W, s);
    return rc;
}

template <typename T>
int hanser_launch(const pyotf_hanser* h, const double* z, int nz, const void* pupilc, int batch,
                  long long pupil_stride, void* psfa, void* psfi, cudaStream_t s, const double* z_host) {
    HanserArgs<T> a;
    a.z_in_params = 0;
    const bool fused_n = h->N == 32 || h->N == 64 || h->N == 128 || h->N == 256;
    // host-side zrange: short stacks of the fused path travel in the kernel arguments, anything else is uploaded
    // into a buffer the handle owns (stream-ordered: the copy queues behind the previous launch that reads it)
    const bool zparams = z_host && nz <= PYOTF_Z_PARAM_MAX && fused_n;
    auto upload_z = [&]() -> int {
        pyotf_hanser* hm = const_cast<pyotf_hanser*>(h);
        const size_t need = sizeof(double) * (size_t)nz;
        if (need > hm->zbuf_cap) {
            PYOTF_CUDA_OK(cudaStreamSynchronize(s));
            cudaFree(hm->zbuf); hm->zbuf = nullptr; hm->zbuf_cap = 0;
            PYOTF_CUDA_OK(cudaMalloc(&hm->zbuf, need));
            hm->zbuf_cap = need;
        }
        PYOTF_CUDA_OK(cudaMemcpyAsync(hm->zbuf, z_host, need, cudaMemcpyHostToDevice, s));
        z = hm->zbuf;
        return 0;
    };
    if (z_host && !zparams) { if (int rc = upload_z()) return rc; }
    if (zparams) {
        for (int i = 0; i < nz; ++i) a.zv[i] = z_host[i];
        a.z_in_params = 1;
    }
    a.kzc = h->kzc; a.ampc = (const T*)h->ampc; a.pupilc = (const cplx<T>*)pupilc;
    a.tw = (const cplx<T>*)h->tw; a.z = z; a.psfi = (T*)psfi; a.psfa = (cplx<T>*)psfa;
    a.K = h->K; a.KR = h->KR; a.KP = h->KP; a.f0 = h->f0; a.nz = nz; a.nvec = h->nvec; a.pupil_stride = pupil_stride;
    // z by value: the launch reads nothing but the handle's immutable tables before its first store, so
    // back-to-back single stacks may overlap without any promise from the caller (only the PDL launch path of
    // hanser_launch_nmt looks at this flag)
    a.overlap = (h->overlap || zparams) ? 1 : 0;
    int smem_limit = 0;
    PYOTF_CUDA_OK(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    a.pupil_has_amp = 0;
    const bool fused_size = h->N == 32 || h->N == 64 || h->N == 128 || h->N == 256;
    if (pupilc && h->nvec == 1 && fused_size) {
        // scalar model: fold the apodisation into the pupil(s) once per launch (scratch owned by the handle;
        // like the table scratch it makes a handle single-stream)
        pyotf_hanser* hm = const_cast<pyotf_hanser*>(h);
        const size_t KKP = (size_t)h->KR * h->K;
        const long long total = (long long)KKP * (pupil_stride ? batch : 1);
        const size_t need = (size_t)total * sizeof(cplx<T>);
        if (need > hm->pscratch_cap) {
            PYOTF_CUDA_OK(cudaStreamSynchronize(s));
            cudaFree(hm->pscratch); hm->pscratch = nullptr; hm->pscratch_cap = 0;
            PYOTF_CUDA_OK(cudaMalloc(&hm->pscratch, need));
            hm->pscratch_cap = need;
        }
        PYOTF_REQUIRE(pupil_stride == 0 || pupil_stride == (long long)KKP, "hanser_psf: pupil_stride must be 0 or rows*K");
        pupil_premul_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>((const cplx<T>*)pupilc, (const T*)h->ampc,
                                                                              (cplx<T>*)hm->pscratch, (int)KKP, total);
        PYOTF_CUDA_OK(cudaGetLastError());
        a.pupilc = (const cplx<T>*)hm->pscratch;
        a.pupil_has_amp = 1;
    }
    a.dtab = nullptr; a.amp_in_tab = 0;
    a.octkz = h->octkz; a.octamp = (const T*)h->octamp; a.octij = h->octij;
    const int ntab = h->octkz ? hanser_ntab<T>(h) : 0;
    const bool fused = h->N == 32 || h->N == 64 || h->N == 128 || h->N == 256;
    int tabm = 0;
    if (ntab && fused) {
        a.amp_in_tab = (!pupilc && h->nvec == 1) ? 1 : 0;
        int M = 0, rc = 0;
        switch (h->N) {
            case 32:  rc = hanser_launch_n<T, 32>(h, a, 0, batch, smem_limit, s, &M); break;
            case 64:  rc = hanser_launch_n<T, 64>(h, a, 0, batch, smem_limit, s, &M); break;
            case 128: rc = hanser_launch_n<T, 128>(h, a, 0, batch, smem_limit, s, &M); break;
            default:  rc = hanser_launch_n<T, 256>(h, a, 0, batch, smem_limit, s, &M); break;
        }
        if (rc) return rc;
        static const int k1_tab = [] { const char* e = getenv("PYOTF_B200_K1_TABLE"); return e ? atoi(e) : 0; }();
#ifdef PYOTF_ENABLE_CLUSTER_TABLE
        const bool cluster_tab = batch == 1 && k1_tab == 2 && h->N / M <= 8;
#else
        const bool cluster_tab = false;
#endif
        if (cluster_tab) {
            tabm = 2;       // experiment: the plane's CTA cluster builds the table through DSMEM (measured slower)
        } else if (batch == 1 && k1_tab != 1) {
            tabm = 3;       // one stack: every CTA builds its own table (no separate launch)
        } else {
            // a batch shares the per-plane tables: one small launch per stack; the handle owns the scratch.
            // NOTE: a handle must not be used from two streams at once (the scratch is shared).
            tabm = 1;
            pyotf_hanser* hm = const_cast<pyotf_hanser*>(h);
            const size_t need = (size_t)nz * ntab * sizeof(cplx<T>);
            if (need > hm->dtab_cap) {
                PYOTF_CUDA_OK(cudaStreamSynchronize(s));
                cudaFree(hm->dtab); hm->dtab = nullptr; hm->dtab_cap = 0;
                PYOTF_CUDA_OK(cudaMalloc(&hm->dtab, need));
                hm->dtab_cap = need;
            }
            if (zparams) { if (int rcz = upload_z()) return rcz; a.z = z; }   // the table kernel reads z from memory
            rc = hanser_build_dtabs<T>(h, z, nz, a.amp_in_tab != 0, (cplx<T>*)hm->dtab, s);
            if (rc) return rc;
            a.dtab = (const cplx<T>*)hm->dtab;
        }
    }
    if (h->N == 256 && tabm == 3 && a.amp_in_tab && !psfa && psfi && -h->f0 <= 63 && h->octkz && hanser_use_sym(a)) {
        // the reference's default model: whole-plane CTAs exploiting the even / transpose symmetry (hanser_sym.cuh)
        static const bool old_sym = getenv("PYOTF_B200_K1_SYM_OLD") != nullptr;
        if (!old_sym) {
            bool done = false;
            int rc = hanser_sym_launch<T>(a, batch, smem_limit, s, &done);
            if (rc || done) return rc;
        }
    }
    switch (h->N) {
        case 32:  return hanser_launch_n<T, 32>(h, a, tabm, batch, smem_limit, s);
        case 64:  return hanser_launch_n<T, 64>(h, a, tabm, batch, smem_limit, s);
        case 128: return hanser_launch_n<T, 128>(h, a, tabm, batch, smem_limit, s);
        case 256: return hanser_launch_n<T, 256>(h, a, tabm, batch, smem_limit, s);
        default: break;
    }
    if (h->N == 1024 && !pupilc && !psfa && psfi && h->nvec == 1) {
        static const bool off = getenv("PYOTF_B200_K1BS_OFF") != nullptr;      // A/B knob: the generic two-pass path
        if (!off) {
            bool done = false;
            int rc = hanser_bigsym_launch<T>(h, z, nz, batch, psfi, smem_limit, s, &done);
            if (rc || done) return rc;
        }
    }
    if (h->N == 1024 && !psfa && psfi) {
        static const bool off = getenv("PYOTF_B200_K1BR_OFF") != nullptr;      // A/B knob: the staged two-pass path
        if (!off) {
            bool done = false;
            int rc = hanser_biggen_launch<T>(h, z, nz, pupilc, batch, pupil_stride, psfi, smem_limit, s, &done);
            if (rc || done) return rc;
        }
    }
    return hanser_launch_big<T>(h, z, nz, pupilc, batch, pupil_stride, psfa, psfi, smem_limit, s);
}

// K1 on caller-built per-plane quadrant tables, no handle: plane p of the stack is fftshift(ifft2(P_p)) with
// P_p(fy, fx) = tables[p][|fy| * dtab_pitch(f0) + |fx|] for |fy|, |fx| <= -f0 (a pupil even in ky and in kx -- what
// `pupil * amp * defocus` is for the ideal pupil of the scalar HanserPSF model, and what the z transform of the scalar
// SheppardPSF cap is at every z, sheppard_impl.cuh).  Runs the even-symmetric variant of the fused kernel (half the
// columns, half the rows).  psfa [nz][N][N] and / or psfi = |psfa|^2.  Returns 3 when the support box does not fit the
// fused on-chip kernel (the caller falls back).
template <typename T> __global__ void hanser_tables_prep_kernel(int N, cplx<T>* __restrict__ tw, double* __restrict__ z, int nz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) {                                       // tw[k] = e^{+2 pi i k / N} (twiddle_kernel)
        double sn, cs;
        sincospi(2.0 * (double)i / (double)N, &sn, &cs);
        tw[i] = mk<T>((T)cs, (T)sn);
    }
    if (i < nz) z[i] = 0.0;                            // the tables are final: no plane position enters
}

// tw_z: optional device buffer the caller has ALREADY filled with K1's N twiddles e^{+2 pi i k / N} (cplx<T>[N]) followed,
// 8-byte aligned, by nz zeros (double[nz]) -- hanser_tables_aux_bytes() -- so that no preparation launch is needed here.
inline size_t hanser_tables_aux_bytes(int N, int nz, size_t cplx_bytes) { return cplx_bytes * (size_t)N + sizeof(double) * (size_t)nz; }

template <typename T>
int hanser_launch_tables(int N, int f0, const void* tables, int nz, void* psfa, void* psfi, cudaStream_t s, const void* tw_z) {
    if (!(N == 32 || N == 64 || N == 128 || N == 256) || f0 >= 0 || -f0 >= N / 2 || nz < 1 || nz > 16383) return 3;
    pyotf_hanser hd;                                   // the launch layers read the geometry and the device only
    PYOTF_CUDA_OK(cudaGetDevice(&hd.device));
    hd.N = N; hd.f0 = f0; hd.K = 1 - 2 * f0;
    hd.KP = hd.K + ((8 - hd.K % 16) + 16) % 16;
    hd.KR = hd.K + PYOTF_PAD_ROWS; hd.nvec = 1;
    const int ntab = hanser_ntab<T>(&hd);
    int smem_limit = 0;
    PYOTF_CUDA_OK(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, hd.device));
    size_t need = (size_t)-1;
    switch (N) {                                       // (launch_n sizes its choice on the full pitch)
        case 32:  need = hanser_smem_for<T, 32>(32, hd.KP); break;
        case 64:  need = hanser_smem_for<T, 64>(32, hd.KP); break;
        case 128: need = hanser_smem_for<T, 128>(32, hd.KP); break;
        default:  need = hanser_smem_for<T, 256>(32, hd.KP); break;
    }
    if (!ntab || need > (size_t)smem_limit) return 3;
    keep_scratch_pool();
    char* tmp = nullptr;
    const size_t tw_bytes = sizeof(cplx<T>) * (size_t)N;
    const cplx<T>* tw = (const cplx<T>*)tw_z;
    const double* z = tw_z ? (const double*)((const char*)tw_z + tw_bytes) : nullptr;
    if (!tw_z) {
        PYOTF_CUDA_OK(cudaMallocAsync(&tmp, hanser_tables_aux_bytes(N, nz, sizeof(cplx<T>)), s));
        hanser_tables_prep_kernel<T><<<(std::max(N, nz) + 255) / 256, 256, 0, s>>>(N, (cplx<T>*)tmp, (double*)(tmp + tw_bytes), nz);
        PYOTF_CUDA_OK(cudaGetLastError());
        tw = (const cplx<T>*)tmp; z = (const double*)(tmp + tw_bytes);
    }
    HanserArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.tw = tw; a.dtab = (const cplx<T>*)tables; a.z = z;
    a.amp_in_tab = 1;                                  // nothing multiplies the table: the even-symmetric variant runs
    a.psfi = (T*)psfi; a.psfa = (cplx<T>*)psfa;
    a.K = hd.K; a.KR = hd.KR; a.KP = hd.KP; a.f0 = f0; a.nz = nz; a.nvec = 1;
    int rc;
    switch (N) {
        case 32:  rc = hanser_launch_n<T, 32>(&hd, a, 1, 1, smem_limit, s); break;
        case 64:  rc = hanser_launch_n<T, 64>(&hd, a, 1, 1, smem_limit, s); break;
        case 128: rc = hanser_launch_n<T, 128>(&hd, a, 1, 1, smem_limit, s); break;
        default:  rc = hanser_launch_n<T, 256>(&hd, a, 1, 1, smem_limit, s); break;
    }
    if (tmp) cudaFreeAsync(tmp, s);
    return rc;
}

}  // namespace pyotf
