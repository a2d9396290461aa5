// Per-dtype host launchers for K1 (instantiated in hanser_f32.cu / hanser_f64.cu).
#pragma once
#include "common.cuh"
#include "hanser.cuh"

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

// octant table entries when the support box is symmetric and the table is small, else 0
template <typename T> int hanser_ntab(const pyotf_hanser* h) {
    if (h->K != 2 * (-h->f0) + 1) return 0;
    int H = -h->f0, ntab = (H + 1) * (H + 1);     // quadrant table [|fy|][|fx|]
    return (size_t)ntab * sizeof(cplx<T>) <= 48 * 1024 ? ntab : 0;
}

template <typename T, int N, int M, bool TAB>
int hanser_launch_nmt(const HanserArgs<T>& a, int batch, size_t smem, cudaStream_t s) {
    using SM = HanserSmem<N, M, T>;
    auto kern = hanser_psf_kernel<T, N, M, TAB>;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        PYOTF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((unsigned)((N / M) * a.nz), (unsigned)batch);
    kern<<<grid, SM::NT, smem, s>>>(a);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T, int N, int M>
int hanser_launch_nm(const pyotf_hanser* h, const HanserArgs<T>& a, int batch, cudaStream_t s) {
    int ntab = hanser_ntab<T>(h);
    size_t smem = HanserSmem<N, M, T>::bytes(a.KP, ntab);
    return ntab ? hanser_launch_nmt<T, N, M, true>(a, batch, smem, s)
                : hanser_launch_nmt<T, N, M, false>(a, batch, smem, s);
}

template <typename T, int N> size_t hanser_smem_for(int M, int KP, int ntab) {
    if (M == 16) return HanserSmem<N, 16, T>::bytes(KP, ntab);
    if (M == 32) return HanserSmem<N, (N >= 32 ? 32 : 16), T>::bytes(KP, ntab);
    return HanserSmem<N, (N >= 64 ? 64 : 16), T>::bytes(KP, ntab);
}

template <typename T, int N>
int hanser_launch_n(const pyotf_hanser* h, const HanserArgs<T>& a, int batch, int smem_limit, cudaStream_t s) {
    // largest rows-per-CTA M whose working set allows two CTAs per SM, else the largest that fits
    int pick = 0;
    const int cand[3] = {64, 32, 16};
    const int ntab = hanser_ntab<T>(h);
    for (int pass = 0; pass < 2 && !pick; ++pass)
        for (int i = 0; i < 3 && !pick; ++i) {
            int M = cand[i];
            if (M > N) continue;
            size_t need = hanser_smem_for<T, N>(M, a.KP, ntab);
            size_t lim = pass == 0 ? (size_t)(smem_limit / 2 - 1024) : (size_t)smem_limit;
            if (need <= lim) pick = M;
        }
    PYOTF_REQUIRE(pick != 0, "hanser_psf: pupil support too large for the fused on-chip path at this size/dtype");
    if (pick == 64) { if constexpr (N >= 64) return hanser_launch_nm<T, N, 64>(h, a, batch, s); }
    if (pick == 32) { if constexpr (N >= 32) return hanser_launch_nm<T, N, 32>(h, a, batch, s); }
    return hanser_launch_nm<T, N, 16>(h, a, batch, s);
}

template <typename T>
int hanser_launch(const pyotf_hanser* h, const double* z, int nz, const void* pupilc, int batch,
                  long long pupil_stride, void* psfa, void* psfi, cudaStream_t s) {
    HanserArgs<T> a;
    a.kzc = h->kzc; a.ampc = (const T*)h->ampc; a.pupilc = (const cplx<T>*)pupilc;
    a.tw = (const cplx<T>*)h->tw; a.z = z; a.psfi = (T*)psfi; a.psfa = (cplx<T>*)psfa;
    a.K = h->K; a.KR = h->KR; a.KP = h->KP; a.f0 = h->f0; a.nz = nz; a.nvec = h->nvec; a.pupil_stride = pupil_stride;
    int smem_limit = 0;
    PYOTF_CUDA_OK(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    switch (h->N) {
        case 32:  return hanser_launch_n<T, 32>(h, a, batch, smem_limit, s);
        case 64:  return hanser_launch_n<T, 64>(h, a, batch, smem_limit, s);
        case 128: return hanser_launch_n<T, 128>(h, a, batch, smem_limit, s);
        case 256: return hanser_launch_n<T, 256>(h, a, batch, smem_limit, s);
        default: break;
    }
    set_error("hanser_psf: size must be one of 32, 64, 128, 256 on the fused path");
    return 1;
}

}  // namespace pyotf
