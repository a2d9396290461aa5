// Host side of K2 (instantiated per dtype in pr_f32.cu / pr_f64.cu): state allocation, the
// iteration loop of retrieve_phase (pyotf/phaseretrieval.py:95-135) as back-to-back kernel
// launches with the convergence test on the device, and the optional z-sharded all-reduce.
#pragma once
#include <thread>
#include "common.cuh"
#include "pr.cuh"
#include "pr_big.cuh"

namespace pyotf {

template <typename T> size_t pr_smem_for(int N, int M, int KP) {
    switch (N) {
        case 32:  return hanser_smem_for<T, 32>(M, KP);
        case 64:  return hanser_smem_for<T, 64>(M, KP);
        case 128: return hanser_smem_for<T, 128>(M, KP);
        default:  return hanser_smem_for<T, 256>(M, KP);
    }
}

// rows per CTA: the largest M that lets two CTAs share an SM; fewer rows (more CTAs) when the
// stack is too short to fill the GPU otherwise.
template <typename T> int pr_pick_rows(const pyotf_hanser* h, int nz, int smem_limit, int sm_count, int want) {
    const int cand[3] = {64, 32, 16};
    if (want) {
        for (int i = 0; i < 3; ++i)
            if (cand[i] == want && want <= h->N && (want != 16 || h->N <= 64) &&
                pr_smem_for<T>(h->N, want, h->KP) <= (size_t)smem_limit) return want;
        return 0;
    }
    int pick = 0;
    // as for K1: 16 rows per CTA only when nothing else fits (register spills)
    for (int pass = 0; pass < 3 && !pick; ++pass)
        for (int i = 0; i < 3 && !pick; ++i) {
            int M = cand[i];
            if (M > h->N || (M == 16 && (pass < 2 || h->N > 64))) continue;
            size_t lim = pass == 0 ? (size_t)(smem_limit / 2 - 1024) : (size_t)smem_limit;
            if (pr_smem_for<T>(h->N, M, h->KP) <= lim) pick = M;
        }
    // short stack: fewer rows per CTA (more CTAs) only while everything still fits one wave
    while (pick > 32 && nz * (h->N / pick) < sm_count) {
        const int half = pick / 2;
        const int per_sm = pr_smem_for<T>(h->N, half, h->KP) <= (size_t)(smem_limit / 2 - 1024) ? 2 : 1;
        if (nz * (h->N / half) > sm_count * per_sm) break;
        pick = half;
    }
    return pick;
}

template <typename T, int N, int M, bool TAB, int NTH, bool FUSE = false>
int pr_plane_launch_nmt(const PrPlaneArgs<T>& a, cudaStream_t s) {
    const size_t smem = HanserSmem<N, M, T, NTH>::bytes(a.KP);
    auto kern = pr_plane_kernel<T, N, M, TAB, NTH, FUSE>;
    static thread_local size_t configured[PYOTF_MAX_DEVICES] = {};      // the attribute is per device
    const int dev_slot = current_device_slot();
    if (smem > configured[dev_slot]) {
        PYOTF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev_slot] = smem;
    }
    if (a.pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((N / M) * a.nz)); cfg.blockDim = dim3(NTH); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        PYOTF_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, a));
    } else {
        kern<<<(unsigned)((N / M) * a.nz), NTH, smem, s>>>(a);
    }
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T, int N, int M>
int pr_plane_launch_nm(const pyotf_hanser* h, int nth, const PrPlaneArgs<T>& a, cudaStream_t s) {
    if constexpr (N == 256 && M == 64) {
        // fused finalize (one kernel per iteration): the 256^2 table variants only
        if (a.arrive && a.dtab && nth == 512) return pr_plane_launch_nmt<T, N, M, true, 512, true>(a, s);
    }
    if (nth == 512) {
        // the wide variant needs at least 512 / 32 = 16 warps of column work: only the 256^2 plans
        if constexpr (N == 256 && M == 64)
            return a.dtab ? pr_plane_launch_nmt<T, N, M, true, 512>(a, s) : pr_plane_launch_nmt<T, N, M, false, 512>(a, s);
    }
    return a.dtab ? pr_plane_launch_nmt<T, N, M, true, 256>(a, s) : pr_plane_launch_nmt<T, N, M, false, 256>(a, s);
}

template <typename T, int N> int pr_plane_launch_n(const pyotf_hanser* h, int M, int nth, const PrPlaneArgs<T>& a, cudaStream_t s) {
    if (M == 64) { if constexpr (N >= 64) return pr_plane_launch_nm<T, N, 64>(h, nth, a, s); }
    if (M == 32) { if constexpr (N >= 32) return pr_plane_launch_nm<T, N, 32>(h, nth, a, s); }
    if constexpr (N <= 64) return pr_plane_launch_nm<T, N, 16>(h, nth, a, s);
    set_error("pr: internal error (rows per CTA)");
    return 1;
}

template <typename T> int pr_plane_launch(const pyotf_pr* s, int cur, bool pdl, cudaStream_t st, const PrFinalArgs* fin = nullptr,
                                          int niter = 1) {
    const pyotf_hanser* h = s->h;
    PrPlaneArgs<T> a;
    a.pdl = pdl ? 1 : 0;
    a.arrive = nullptr;
    a.niter = niter;
    if (fin) {
        a.arrive = s->arrive; a.fin = *fin; a.mask = h->maskc; a.pupil0 = (cplx<T>*)s->pupil[0]; a.pupil1 = (cplx<T>*)s->pupil[1];
        a.diffpart = s->diffpart; a.mse = s->mse; a.mse_diff = s->mse_diff; a.pupil_diff = s->pupil_diff; a.state_rw = (PrState*)s->state;
    }
    a.kzc = h->kzc; a.pupilc = (const cplx<T>*)s->pupil[cur]; a.tw = (const cplx<T>*)h->tw; a.dtab = (const cplx<T>*)s->dtab; a.z = s->z;
    a.data = (const T*)s->data; a.partial = (cplx<T>*)s->partial; a.mse_part = s->mse_part; a.state = (const PrState*)s->state;
    a.K = h->K; a.KR = h->KR; a.KP = h->KP; a.f0 = h->f0; a.nz = s->nz;
    switch (h->N) {
        case 32:  return pr_plane_launch_n<T, 32>(h, s->M, s->nth, a, st);
        case 64:  return pr_plane_launch_n<T, 64>(h, s->M, s->nth, a, st);
        case 128: return pr_plane_launch_n<T, 128>(h, s->M, s->nth, a, st);
        case 256: return pr_plane_launch_n<T, 256>(h, s->M, s->nth, a, st);
        default: break;
    }
    set_error("pr: size must be one of 32, 64, 128, 256 on the fused path");
    return 1;
}

template <typename T> int pr_setup(pyotf_pr* s, const void* data, int data_dtype, int want_rows, cudaStream_t st) {
    pyotf_hanser* h = s->h;
    keep_scratch_pool();
    s->stream = st;
    int smem_limit = 0, sms = 0;
    PYOTF_CUDA_OK(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    PYOTF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    const size_t KK = (size_t)h->K * h->K, KKP = (size_t)h->KR * h->K, npix = (size_t)s->nz * h->N * h->N;
    s->nth = 256;
    s->generic = !(h->N == 32 || h->N == 64 || h->N == 128 || h->N == 256) || want_rows == 1;
    if (s->generic) {
        // four line-FFT passes per iteration (pr_big.cuh): one partial pupil per plane
        s->M = 0; s->nparts = s->nz;
        s->nmse = 0;                                   // set by every pass-B launch (its grid size)
        PYOTF_CUDA_OK(cudaMallocAsync(&s->W, (size_t)s->nz * h->K * h->N * sizeof(cplx<T>), st));
        PYOTF_CUDA_OK(cudaMallocAsync(&s->P, npix * sizeof(cplx<T>), st));
    } else if (want_rows == 0 && h->N == 256 && sizeof(T) == 4 && 2 * (s->nz * 4) <= 3 * sms &&
        HanserSmem<256, 64, T, 512>::bytes(h->KP) <= (size_t)smem_limit) {
        // short stacks (up to ~1.5 plane quarters per SM): one wide 512-thread CTA per SM is faster than two
        // narrow ones (27.4 vs 36.3 us at 25 planes); long stacks prefer the narrow ones (68.6 vs 72.5 at 100)
        s->M = 64; s->nth = 512;
    } else if (want_rows < 0) {
        want_rows = -want_rows; s->nth = 512;   // tuning knob: force the wide variant (256^2, M = 64 only)
        s->M = pr_pick_rows<T>(h, s->nz, smem_limit, sms, want_rows);
    } else {
        s->M = pr_pick_rows<T>(h, s->nz, smem_limit, sms, want_rows);
    }
    if (!s->generic) {
        PYOTF_REQUIRE(s->M != 0, "pr_create: no rows-per-CTA choice fits shared memory at this size / dtype");
        s->nparts = s->nz * (h->N / s->M);
        s->nmse = s->nparts;
    }
    s->nblk = (int)((KK + PR_FIN_PIX - 1) / PR_FIN_PIX);
    PYOTF_CUDA_OK(cudaMallocAsync(&s->data, npix * sizeof(T), st));
    PYOTF_CUDA_OK(cudaMallocAsync(&s->pupil[0], KKP * sizeof(cplx<T>), st));
    PYOTF_CUDA_OK(cudaMallocAsync(&s->pupil[1], KKP * sizeof(cplx<T>), st));
    PYOTF_CUDA_OK(cudaMallocAsync(&s->partial, (size_t)s->nparts * KK * sizeof(cplx<T>), st));
    // fused: one squared-error partial per CTA of the plane kernel; generic: per CTA of pass B (>= 4 lines each)
    PYOTF_CUDA_OK(cudaMallocAsync(&s->mse_part, (s->generic ? (size_t)s->nz * h->N / 4 + 1 : (size_t)s->nparts) * sizeof(double), st));
    PYOTF_CUDA_OK(cudaMallocAsync(&s->sums, (KK + 1) * sizeof(double2), st));
    PYOTF_CUDA_OK(cudaMallocAsync(&s->diffpart, (size_t)4 * s->nblk * sizeof(double), st));
    PYOTF_CUDA_OK(cudaMallocAsync(&s->state, sizeof(PrState), st));
    PYOTF_CUDA_OK(cudaMemsetAsync(s->state, 0, sizeof(PrState), st));
    PYOTF_CUDA_OK(cudaMallocAsync(&s->arrive, sizeof(unsigned), st));
    PYOTF_CUDA_OK(cudaMemsetAsync(s->arrive, 0, sizeof(unsigned), st));
    // one kernel per iteration when every CTA of the plane kernel is resident at once (a grid-wide arrival barrier
    // inside the kernel replaces the finalize launch): 256^2, 64 rows per CTA, at most one CTA per SM
    const bool no_fuse = getenv("PYOTF_B200_PR_NO_FUSE") != nullptr;     // read at every create: tests compare the two loops
    s->fuse = (!no_fuse && !s->generic && h->N == 256 && s->M == 64 && s->nth == 512 && s->nparts <= sms && 2 * s->nparts >= sms) ? 1 : 0;   // (the 256-thread CTAs of the fp64 mode finalize faster in a kernel of their own: 64.9 vs 66.8 us)
    const unsigned nb = (unsigned)((npix + 255) / 256);
    if (data_dtype == PYOTF_F32) pr_convert_data_kernel<float, T><<<nb, 256, 0, st>>>((const float*)data, (T*)s->data, npix);
    else pr_convert_data_kernel<double, T><<<nb, 256, 0, st>>>((const double*)data, (T*)s->data, npix);
    PYOTF_CUDA_OK(cudaGetLastError());
    pr_init_pupil_kernel<T><<<(unsigned)((KKP + 255) / 256), 256, 0, st>>>(h->maskc, (int)KKP, (cplx<T>*)s->pupil[0],
                                                                          (cplx<T>*)s->pupil[1]);
    PYOTF_CUDA_OK(cudaGetLastError());
    // the defocus tables depend on the plane positions only: built once per stack
    if (const int ntab = hanser_ntab<T>(h)) {
        PYOTF_CUDA_OK(cudaMallocAsync(&s->dtab, (size_t)s->nz * ntab * sizeof(cplx<T>), st));
        int rc = hanser_build_dtabs<T>(h, s->z, s->nz, false, (cplx<T>*)s->dtab, st);
        if (rc) return rc;
    }
    s->cur = 0; s->applied = 0;
    return 0;
}

template <typename T> int pr_set_pupil(pyotf_pr* s, const void* pupil_c128, cudaStream_t st) {
    const pyotf_hanser* h = s->h;
    const size_t KKP = (size_t)h->KR * h->K;
    if (!pupil_c128) {
        pr_init_pupil_kernel<T><<<(unsigned)((KKP + 255) / 256), 256, 0, st>>>(h->maskc, (int)KKP, (cplx<T>*)s->pupil[0],
                                                                              (cplx<T>*)s->pupil[1]);
        PYOTF_CUDA_OK(cudaGetLastError());
        s->cur = 0; s->applied = 0;
        return 0;
    }
    int rc = hanser_pack<T>(h, pupil_c128, 1, s->pupil[s->cur], st);
    s->applied = s->cur;
    return rc;
}

template <typename T> int pr_get_pupil(const pyotf_pr* s, int which, void* out_c128, cudaStream_t st) {
    const pyotf_hanser* h = s->h;
    const int n = h->N * h->N;
    pr_unpack_pupil_kernel<T><<<(n + 255) / 256, 256, 0, st>>>(h->N, h->K, h->f0,
                                                              (const cplx<T>*)s->pupil[which ? s->applied : s->cur],
                                                              (double2*)out_c128);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

// one iteration's transforms on the four-pass path (pr_big.cuh); the finalize kernel follows as usual
template <typename T> int pr_generic_passes(pyotf_pr* s, int cur, cudaStream_t st) {
    const pyotf_hanser* h = s->h;
    int smem_limit = 0;
    PYOTF_CUDA_OK(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    const int N = h->N, K = h->K;
    BigArgs<T> ba;
    ba.kzc = h->kzc; ba.ampc = (const T*)h->ampc; ba.pupilc = (const cplx<T>*)s->pupil[cur]; ba.tw = (const cplx<T>*)h->tw;
    ba.z = s->z; ba.W = (cplx<T>*)s->W; ba.psfa = nullptr; ba.psfi = nullptr;
    ba.N = N; ba.K = K; ba.f0 = h->f0; ba.nzc = s->nz; ba.accumulate = 0; ba.scale = 1.0;
    BigRow<T> fa; fa.a = ba;
    int rc = fft_lines_launch<1, T, true>(N, (long long)s->nz * K, fa, smem_limit, st);                          // A
    if (rc) return rc;
    PrBigArgs<T> pa;
    pa.kzc = h->kzc; pa.z = s->z; pa.data = (const T*)s->data; pa.W = (cplx<T>*)s->W; pa.P = (cplx<T>*)s->P;
    pa.partial = (cplx<T>*)s->partial; pa.mse_part = s->mse_part; pa.tw = (const cplx<T>*)h->tw;
    pa.state = (const PrState*)s->state; pa.N = N; pa.K = K; pa.f0 = h->f0;
    pa.scale = 1.0 / ((double)N * (double)N);
    PrColInv<T> fb; fb.a = pa;
    rc = fft_lines_launch<1, T, false>(N, (long long)s->nz * N, fb, smem_limit, st, &s->nmse);                   // B
    if (rc) return rc;
    PrColFwd<T> fc; fc.a = pa;
    rc = fft_lines_launch<-1, T, false>(N, (long long)s->nz * N, fc, smem_limit, st);                            // C
    if (rc) return rc;
    PrRowFwd<T> fd; fd.a = pa;
    return fft_lines_launch<-1, T, true>(N, (long long)s->nz * K, fd, smem_limit, st);                           // D
}

int comm_allreduce_f64(pyotf_comm* c, double* buf, long long n, cudaStream_t st);

inline bool dbg_skip_any() { static const bool v = getenv("PYOTF_B200_PR_DEBUG_SKIP") != nullptr; return v; }

template <class K, class... A>
int pr_launch(bool pdl, K kern, dim3 grid, dim3 block, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    PYOTF_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, args...));
    return 0;
}

template <typename T>
int pr_run(pyotf_pr* s, int max_iters, double pupil_tol, double mse_tol, int phase_only, pyotf_comm* comm,
           double* mse_host, double* mse_diff_host, double* pupil_diff_host, int* iters_run_host, cudaStream_t st) {
    const pyotf_hanser* h = s->h;
    const int KK = h->K * h->K;
    if (s->cap_iters < max_iters) {
        cudaFreeAsync(s->mse, st); cudaFreeAsync(s->mse_diff, st); cudaFreeAsync(s->pupil_diff, st);
        s->mse = s->mse_diff = s->pupil_diff = nullptr; s->cap_iters = 0;
        PYOTF_CUDA_OK(cudaMallocAsync(&s->mse, sizeof(double) * max_iters, st));
        PYOTF_CUDA_OK(cudaMallocAsync(&s->mse_diff, sizeof(double) * max_iters, st));
        PYOTF_CUDA_OK(cudaMallocAsync(&s->pupil_diff, sizeof(double) * max_iters, st));
        s->cap_iters = max_iters;
    }
    constexpr int CHK = 8;
    if (!s->h_prog) {
        s->prog_slot = ProgressSlots::get().acquire(&s->h_prog, &s->d_prog);
        if (s->prog_slot < 0) {
            PYOTF_CUDA_OK(cudaHostAlloc(&s->h_prog, sizeof(int) * 2, cudaHostAllocMapped | cudaHostAllocPortable));
            PYOTF_CUDA_OK(cudaHostGetDevicePointer(&s->d_prog, s->h_prog, 0));
        }
    }
    // (nothing of an earlier run is in flight: every run ends with a stream synchronisation)
    volatile int* hp = s->h_prog;
    hp[0] = 0; hp[1] = 0;
    s->in_flight = 1;
    // true: the device has stopped.  Returns once `need` iterations have finished; the host stays at most 2 * CHK
    // iterations ahead of the device so that little is enqueued past an early stop.
    auto stopped_before = [&](int need) {
        for (unsigned spins = 0;; ++spins) {
            if (hp[0]) return true;
            if (hp[1] >= need) return false;
            if ((spins & 1023u) == 1023u) {
                if (cudaStreamQuery(st) != cudaErrorNotReady) return hp[0] != 0;    // drained (or failed): nothing to wait for
                std::this_thread::yield();
            }
        }
    };
    PYOTF_CUDA_OK(cudaMemsetAsync(s->state, 0, sizeof(PrState), st));
    PrFinalArgs fa;
    fa.host_prog = (int*)s->d_prog;
    fa.K = h->K; fa.phase_only = phase_only;
    fa.npix_total = (double)s->nz_total * h->N * h->N; fa.nz_total = (double)s->nz_total;
    fa.pupil_tol = pupil_tol; fa.mse_tol = mse_tol;
    const int start = s->cur;
    int launched = 0;
    // programmatic dependent launch for the single-GPU fused loop (kernels overlap their neighbours' tails)
    static const bool no_pdl = getenv("PYOTF_B200_NO_PDL") != nullptr;
    // z-sharded: with a peer window (pyotf_b200_pr_window_*) the cross-rank sum happens inside the finalize kernel
    // over NVLink and the loop keeps its programmatic launches; without one it falls back to an NCCL all-reduce
    const bool window = comm && s->win && s->win_world >= 1;
    const bool use_pdl = (!comm || window) && !s->generic && !no_pdl;
    const bool fused = s->fuse && !comm && s->dtab && !dbg_skip_any();
    if (fused) PYOTF_CUDA_OK(cudaMemsetAsync(s->arrive, 0, sizeof(unsigned), st));
    // iterations per fused launch: CHK (the poll period) unless PYOTF_B200_PR_ITERS_PER_LAUNCH says otherwise (1 = the
    // one-launch-per-iteration loop; the tests compare the two)
    int chunk = CHK;
    if (const char* e = getenv("PYOTF_B200_PR_ITERS_PER_LAUNCH")) { chunk = std::max(1, std::min(CHK, atoi(e))); if (CHK % chunk) chunk = 1; }
    s->kernel_launches = 0;
    PrPeers peers = {};
    if (window) { peers.world = s->win_world; peers.rank = s->win_rank; for (int r = 0; r < s->win_world; ++r) peers.win[r] = s->peer[r]; }
    for (int i = 0; i < max_iters; ++i) {
        if (i % CHK == 0 && i >= 2 * CHK && stopped_before(i - CHK)) break;   // stay at most ~2*CHK iterations ahead of the device
        const int cur = (start + i) & 1;
        static const int dbg_skip = [] { const char* e = getenv("PYOTF_B200_PR_DEBUG_SKIP"); return e ? atoi(e) : 0; }();   // timing experiments only: 1 = no plane kernel, 2 = no finalize
        int rc = 0;
        fa.iter = i; fa.cur = cur; fa.last_iter = i == max_iters - 1; fa.nmse = s->nmse;
        if (fused) {
            // one launch runs the iterations up to the next stop-flag poll (at most CHK) behind in-kernel grid barriers
            const int n = std::min(chunk - i % chunk, max_iters - i);
            fa.nparts = s->nparts; fa.last_iter = i + n == max_iters;
            rc = pr_plane_launch<T>(s, cur, use_pdl, st, &fa, n);
            if (rc) return rc;
            launched += n; ++s->kernel_launches;
            i += n - 1;
            continue;
        }
        if (dbg_skip != 1) rc = s->generic ? pr_generic_passes<T>(s, cur, st) : pr_plane_launch<T>(s, cur, use_pdl, st);
        if (rc) return rc;
        fa.nmse = s->nmse;                              // the generic pass B sets it at every launch
        if (dbg_skip == 2) { ++launched; continue; }
        if (window) {
            peers.slot = i & 1; peers.epoch = s->epoch_base + i + 1;
            fa.nparts = s->nparts;
            rc = pr_launch(use_pdl, pr_finalize_peers_kernel<T>, dim3((unsigned)s->nblk), dim3(256), st, fa, peers,
                           (const cplx<T>*)s->partial, (const double*)s->mse_part, (const unsigned char*)h->maskc,
                           (cplx<T>*)s->pupil[0], (cplx<T>*)s->pupil[1], s->diffpart, s->mse, s->mse_diff, s->pupil_diff,
                           (PrState*)s->state);
            if (rc) return rc;
        } else if (comm) {
            pr_reduce_kernel<T><<<(KK + 1 + 255) / 256, 256, 0, st>>>((const cplx<T>*)s->partial, s->mse_part, s->nparts, s->nmse, KK,
                                                                     (double2*)s->sums, (const PrState*)s->state);
            PYOTF_CUDA_OK(cudaGetLastError());
            rc = comm_allreduce_f64(comm, (double*)s->sums, 2ll * (KK + 1), st);
            if (rc) return rc;
            fa.nparts = 0;
            pr_finalize_kernel<T, double2><<<s->nblk, 256, 0, st>>>(fa, (const double2*)s->sums, s->mse_part, h->maskc,
                                                                   (cplx<T>*)s->pupil[0], (cplx<T>*)s->pupil[1], s->diffpart,
                                                                   s->mse, s->mse_diff, s->pupil_diff, (PrState*)s->state);
        } else {
            fa.nparts = s->nparts;
            if (use_pdl) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)s->nblk); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                PYOTF_CUDA_OK(cudaLaunchKernelEx(&cfg, pr_finalize_kernel<T, cplx<T>>, fa, (const cplx<T>*)s->partial,
                                                 (const double*)s->mse_part, (const unsigned char*)h->maskc,
                                                 (cplx<T>*)s->pupil[0], (cplx<T>*)s->pupil[1], s->diffpart, s->mse,
                                                 s->mse_diff, s->pupil_diff, (PrState*)s->state));
            } else {
                pr_finalize_kernel<T, cplx<T>><<<s->nblk, 256, 0, st>>>(fa, (const cplx<T>*)s->partial, s->mse_part, h->maskc,
                                                                       (cplx<T>*)s->pupil[0], (cplx<T>*)s->pupil[1], s->diffpart,
                                                                       s->mse, s->mse_diff, s->pupil_diff, (PrState*)s->state);
            }
        }
        PYOTF_CUDA_OK(cudaGetLastError());
        ++launched;
    }
    PrState hs;
    PYOTF_CUDA_OK(cudaMemcpyAsync(&hs, s->state, sizeof(PrState), cudaMemcpyDeviceToHost, st));
    PYOTF_CUDA_OK(cudaStreamSynchronize(st));
    s->in_flight = 0;
    const int n = hs.iters_run;
    if (window) s->epoch_base += max_iters;            // the same on every rank, whatever was enqueued after a stop
    PYOTF_REQUIRE(hs.stopped != 2, "pr_run: a peer rank did not publish its partial sums in time (z-sharded run)");
    PYOTF_REQUIRE(hs.stopped != 3, "pr_run: the grid-wide barrier of the fused iteration timed out (plane-kernel CTAs not co-resident)");
    PYOTF_REQUIRE(n >= 1 && n <= launched, "pr_run: inconsistent device state");
    s->applied = (start + n - 1) & 1;                  // the pupil iteration n-1 applied   (result.model)
    s->cur = hs.stopped ? s->applied : (s->applied ^ 1);   // final new_pupil                  :144-148
    s->launched = launched;
    if (mse_host) PYOTF_CUDA_OK(cudaMemcpy(mse_host, s->mse, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (mse_diff_host) PYOTF_CUDA_OK(cudaMemcpy(mse_diff_host, s->mse_diff, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (pupil_diff_host) PYOTF_CUDA_OK(cudaMemcpy(pupil_diff_host, s->pupil_diff, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (iters_run_host) *iters_run_host = n;
    return 0;
}

}  // namespace pyotf
