// extern "C" boundary: argument validation, handle lifetime, dtype dispatch.
#include <cstring>

#include "common.cuh"
#include "hanser.cuh"
#include "zernike.cuh"
#include "sheppard.cuh"
#include "prep.cuh"
#include "pr.cuh"

namespace pyotf {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace pyotf

using namespace pyotf;

template <typename T>
static int aberration_launch(const pyotf_hanser* h, const int* n, const int* m, int nmodes, const double* mc,
                             const double* pc, int batch, void* out, cudaStream_t s) {
    constexpr int PIX = 64;
    AberrationParams p;
    p.N = h->N; p.K = h->K; p.KR = h->KR; p.f0 = h->f0; p.nmodes = nmodes; p.batch = batch;
    p.has_mag = mc != nullptr; p.has_phase = pc != nullptr;
    p.kstep = 1.0 / ((double)h->N * h->p.res); p.wl = h->p.wl; p.na = h->p.na; p.diff_limit = h->p.na / h->p.wl;
    size_t smem = sizeof(double) * ((size_t)nmodes * PIX + PIX);
    auto kern = aberration_pupil_kernel<T, PIX>;
    PYOTF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int KK = h->KR * h->K;
    int gy = batch >= 256 ? 4 : 1;
    dim3 grid((unsigned)((KK + PIX - 1) / PIX), (unsigned)gy);
    kern<<<grid, 256, smem, s>>>(p, n, m, mc, pc, (cplx<T>*)out);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}


static int sheppard_params(const pyotf_sheppard_params* q, SheppardParams* out) {
    PYOTF_REQUIRE(q->wl > 0 && q->na > 0 && q->ni > 0 && q->res > 0 && q->zres > 0, "sheppard: wl, na, ni, res, zres must be positive");
    PYOTF_REQUIRE(q->size > 0 && q->zsize > 0 && q->zsize <= 65535 && q->size <= 65535, "sheppard: bad size");
    PYOTF_REQUIRE(q->vec_corr >= 0 && q->vec_corr <= 4 && q->condition >= 0 && q->condition <= 2, "sheppard: bad enum");
    PYOTF_REQUIRE(q->dtype == PYOTF_F32 || q->dtype == PYOTF_F64, "sheppard: bad dtype");
    SheppardParams p;
    p.N = q->size; p.NZ = q->zsize; p.vec_mode = q->vec_corr; p.condition = q->condition; p.dual = q->dual ? 1 : 0;
    p.nvec = q->vec_corr == PYOTF_VEC_NONE ? 1 : (q->vec_corr == PYOTF_VEC_TOTAL ? 6 : 2);
    p.kstep = 1.0 / ((double)q->size * q->res);          // numpy.fft.fftfreq: val = 1.0 / (n * d)
    p.kzstep = 1.0 / ((double)q->zsize * q->zres);
    // dk = k[1] - k[0] (otf.py:508); a length-1 axis has no k[1]: the reference raises IndexError there
    PYOTF_REQUIRE(q->size > 1 && q->zsize > 1, "sheppard: size and zsize must be > 1");
    p.dk = (q->size == 2 ? -1.0 : 1.0) * p.kstep;          // fftfreq(2) = [0, -1] * val
    p.dkz = (q->zsize == 2 ? -1.0 : 1.0) * p.kzstep;
    p.same_dk = p.dk == p.dkz;
    p.kmag = q->ni / q->wl;
    const double nawl = q->na / q->wl;
    const double arg = p.kmag * p.kmag - nawl * nawl;      // otf.py:512
    PYOTF_REQUIRE(arg >= 0, "sheppard: na > ni");
    p.kz_min = sqrt(arg);
    // |kr - kmag| < dkr and dkr <= max(|dk|, |dkz|) (a weighted norm over the plain norm): bounds on kr^2, widened
    const double dmax = fmax(fabs(p.dk), fabs(p.dkz)) * (1.0 + 1e-9);
    const double lo = fmax(0.0, p.kmag - dmax), hi = p.kmag + dmax;
    p.r2_lo = lo * lo * (1.0 - 1e-12); p.r2_hi = hi * hi * (1.0 + 1e-12);
    if (p.dk < 0 || p.dkz < 0) { p.r2_lo = -1.0; p.r2_hi = 1e300; }   // degenerate 2-sample axes: no prefilter
    *out = p;
    return 0;
}

template <typename T>
static void sheppard_generate(const SheppardParams& p, void* otfa, unsigned char* valid, unsigned char* rowflag,
                              unsigned char* planeflag, cudaStream_t s) {
    dim3 grid((unsigned)((p.N + 255) / 256), (unsigned)p.N, (unsigned)p.NZ);
    sheppard_otf_kernel<T><<<grid, 256, 0, s>>>(p, (cplx<T>*)otfa, valid, rowflag, planeflag);
}


extern "C" {

int pyotf_b200_abi_version(void) { return PYOTF_B200_ABI_VERSION; }
int pyotf_b200_version(void) { return PYOTF_B200_ABI_VERSION; }          // the name SURVEY.md section 8(b) lists
const char* pyotf_b200_last_error(void) { return g_last_error.c_str(); }

int pyotf_b200_device_query(int device, int* sm_count, int* smem_optin_bytes, int* cc_major, int* cc_minor) {
    int v;
    if (sm_count) { PYOTF_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device)); *sm_count = v; }
    if (smem_optin_bytes) { PYOTF_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)); *smem_optin_bytes = v; }
    if (cc_major) { PYOTF_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device)); *cc_major = v; }
    if (cc_minor) { PYOTF_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device)); *cc_minor = v; }
    return 0;
}

int pyotf_b200_hanser_create(const pyotf_hanser_params* p, void* stream, pyotf_hanser_t** out) {
    PYOTF_REQUIRE(p && out, "hanser_create: null argument");
    PYOTF_REQUIRE(p->wl > 0 && p->na > 0 && p->ni > 0 && p->res > 0, "hanser_create: wl, na, ni, res must be positive");
    PYOTF_REQUIRE(p->size >= 2 && p->size <= 16384, "hanser_create: size must be in 2..16384");
    PYOTF_REQUIRE(p->vec_corr >= 0 && p->vec_corr <= 4, "hanser_create: bad vec_corr");
    PYOTF_REQUIRE(p->condition >= 0 && p->condition <= 2, "hanser_create: bad condition");
    PYOTF_REQUIRE(p->dtype == PYOTF_F32 || p->dtype == PYOTF_F64, "hanser_create: bad dtype");
    pyotf_hanser* h = new pyotf_hanser();
    h->p = *p;
    h->N = p->size;
    PYOTF_CUDA_OK(cudaGetDevice(&h->device));
    PYOTF_REQUIRE(p->support >= -2, "hanser_create: bad support");
    int H;
    if (p->support < 0) {
        // conservative bound on the ideal pupil radius in bins: kr <= na/wl, k = f / (N res)
        double rad = (p->na / p->wl) * (double)p->size * p->res;
        H = (int)rad + 1;
    } else {
        H = p->support;
    }
    if (H >= h->N / 2) { h->f0 = -h->N / 2; h->K = h->N; }
    else { h->f0 = -H; h->K = 2 * H + 1; }
    // row pitch == 8 (mod 16) elements keeps the two row groups of a half-warp on disjoint banks
    h->KP = h->K + ((8 - h->K % 16) + 16) % 16;
    h->nvec = p->vec_corr == PYOTF_VEC_NONE ? 1 : (p->vec_corr == PYOTF_VEC_TOTAL ? 6 : 2);
    h->KR = h->K + PYOTF_PAD_ROWS;
    size_t KK = (size_t)h->KR * h->K, es = p->dtype == PYOTF_F32 ? 4 : 8;
    cudaStream_t s = (cudaStream_t)stream;
    PYOTF_CUDA_OK(cudaMalloc(&h->kzc, KK * sizeof(double)));
    PYOTF_CUDA_OK(cudaMalloc(&h->ampc, KK * h->nvec * es));
    PYOTF_CUDA_OK(cudaMalloc(&h->maskc, KK));
    PYOTF_CUDA_OK(cudaMalloc(&h->tw, (size_t)h->N * 2 * es));
    int rc = p->dtype == PYOTF_F32 ? hanser_build_tables<float>(h, s) : hanser_build_tables<double>(h, s);
    if (rc) { pyotf_b200_hanser_destroy(h); return rc; }
    // the tables are immutable from here on and visible to every later launch, including launches that read them
    // before waiting for their predecessor in the stream (programmatic dependent launch)
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { set_error(std::string("hanser_create: ") + cudaGetErrorString(e)); pyotf_b200_hanser_destroy(h); return 2; }
    *out = h;
    return 0;
}

int pyotf_b200_hanser_destroy(pyotf_hanser_t* h) {
    if (!h) return 0;
    cudaFree(h->kzc); cudaFree(h->ampc); cudaFree(h->maskc); cudaFree(h->tw); cudaFree(h->dtab); cudaFree(h->octkz); cudaFree(h->octamp); cudaFree(h->octij); cudaFree(h->pscratch); cudaFree(h->zbuf);
    delete h;
    return 0;
}

int pyotf_b200_hanser_set_overlap(pyotf_hanser_t* h, int enable) {
    PYOTF_REQUIRE(h, "hanser_set_overlap: null handle");
    h->overlap = enable ? 1 : 0;
    return 0;
}

int pyotf_b200_hanser_info(const pyotf_hanser_t* h, int* K, int* rows, int* f0, int* nvec) {
    PYOTF_REQUIRE(h, "hanser_info: null handle");
    if (K) *K = h->K;
    if (rows) *rows = h->KR;
    if (f0) *f0 = h->f0;
    if (nvec) *nvec = h->nvec;
    return 0;
}

int pyotf_b200_hanser_tables(const pyotf_hanser_t* h, double* kz_dev, void* amp_dev, unsigned char* mask_dev,
                             void* stream) {
    PYOTF_REQUIRE(h, "hanser_tables: null handle");
    cudaStream_t s = (cudaStream_t)stream;
    size_t KK = (size_t)h->KR * h->K, es = h->p.dtype == PYOTF_F32 ? 4 : 8;
    if (kz_dev) PYOTF_CUDA_OK(cudaMemcpyAsync(kz_dev, h->kzc, KK * 8, cudaMemcpyDeviceToDevice, s));
    if (amp_dev) PYOTF_CUDA_OK(cudaMemcpyAsync(amp_dev, h->ampc, KK * h->nvec * es, cudaMemcpyDeviceToDevice, s));
    if (mask_dev) PYOTF_CUDA_OK(cudaMemcpyAsync(mask_dev, h->maskc, KK, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int pyotf_b200_pupil_support(const void* pupil, int size, int* hmax_host, void* stream) {
    PYOTF_REQUIRE(pupil && hmax_host, "pupil_support: null argument");
    PYOTF_REQUIRE(size >= 2, "pupil_support: bad size");
    cudaStream_t s = (cudaStream_t)stream;
    int* d = nullptr;
    PYOTF_CUDA_OK(cudaMallocAsync(&d, sizeof(int), s));
    PYOTF_CUDA_OK(cudaMemsetAsync(d, 0, sizeof(int), s));
    int n = size * size;
    pupil_support_kernel<<<(n + 255) / 256, 256, 0, s>>>(size, (const double2*)pupil, d);
    PYOTF_CUDA_OK(cudaGetLastError());
    PYOTF_CUDA_OK(cudaMemcpyAsync(hmax_host, d, sizeof(int), cudaMemcpyDeviceToHost, s));
    PYOTF_CUDA_OK(cudaFreeAsync(d, s));
    PYOTF_CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int pyotf_b200_hanser_pack_pupil(const pyotf_hanser_t* h, const void* pupil, int batch, void* out, void* stream) {
    PYOTF_REQUIRE(h && pupil && out && batch > 0, "hanser_pack_pupil: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    return h->p.dtype == PYOTF_F32 ? hanser_pack<float>(h, pupil, batch, out, s)
                                   : hanser_pack<double>(h, pupil, batch, out, s);
}

int pyotf_b200_hanser_psf(const pyotf_hanser_t* h, const double* z_dev, int nz, const void* pupilc_dev, int batch,
                          long long pupil_stride, void* psfa_out_dev, void* psfi_out_dev, void* stream) {
    PYOTF_REQUIRE(h && z_dev, "hanser_psf: null argument");
    PYOTF_REQUIRE(nz > 0 && batch > 0, "hanser_psf: nz and batch must be positive");
    PYOTF_REQUIRE(psfa_out_dev || psfi_out_dev, "hanser_psf: no output requested");
    PYOTF_REQUIRE((h->p.support == -1) == (pupilc_dev == nullptr),
                  "hanser_psf: handle support mode and pupil argument disagree");
    PYOTF_REQUIRE(batch <= 65535, "hanser_psf: batch too large for one launch");
    cudaStream_t s = (cudaStream_t)stream;
    return h->p.dtype == PYOTF_F32
               ? hanser_launch<float>(h, z_dev, nz, pupilc_dev, batch, pupil_stride, psfa_out_dev, psfi_out_dev, s)
               : hanser_launch<double>(h, z_dev, nz, pupilc_dev, batch, pupil_stride, psfa_out_dev, psfi_out_dev, s);
}

int pyotf_b200_hanser_psf_zhost(pyotf_hanser_t* h, const double* z_host, int nz, const void* pupilc_dev, int batch,
                                long long pupil_stride, void* psfa_out_dev, void* psfi_out_dev, void* stream) {
    PYOTF_REQUIRE(h && z_host, "hanser_psf_zhost: null argument");
    PYOTF_REQUIRE(nz > 0 && batch > 0, "hanser_psf_zhost: nz and batch must be positive");
    PYOTF_REQUIRE(psfa_out_dev || psfi_out_dev, "hanser_psf_zhost: no output requested");
    PYOTF_REQUIRE((h->p.support == -1) == (pupilc_dev == nullptr),
                  "hanser_psf_zhost: handle support mode and pupil argument disagree");
    PYOTF_REQUIRE(batch <= 65535, "hanser_psf_zhost: batch too large for one launch");
    cudaStream_t s = (cudaStream_t)stream;
    return h->p.dtype == PYOTF_F32
               ? hanser_launch<float>(h, nullptr, nz, pupilc_dev, batch, pupil_stride, psfa_out_dev, psfi_out_dev, s, z_host)
               : hanser_launch<double>(h, nullptr, nz, pupilc_dev, batch, pupil_stride, psfa_out_dev, psfi_out_dev, s, z_host);
}

int pyotf_b200_hanser_kspace(double wl, double ni, double res, int size, double* kr, double* phi, double* kz,
                             void* stream) {
    PYOTF_REQUIRE(size > 0 && wl > 0 && res > 0, "hanser_kspace: bad argument");
    int n = size * size;
    hanser_kspace_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(size, 1.0 / ((double)size * res), ni / wl,
                                                                           kr, phi, kz);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

int pyotf_b200_hanser_defocus(const double* kz, long long npix, const double* z, int nz, void* out, void* stream) {
    PYOTF_REQUIRE(kz && z && out && npix > 0 && nz > 0 && nz <= 65535, "hanser_defocus: bad argument");
    dim3 grid((unsigned)((npix + 255) / 256), (unsigned)nz);
    hanser_defocus_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((int)npix, kz, z, nz, (double2*)out);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

int pyotf_b200_zernike(const double* r, const double* theta, long long npts, const int* n, const int* m, int nmodes,
                       int norm, double* out, void* stream) {
    PYOTF_REQUIRE(r && theta && n && m && out && npts > 0 && nmodes > 0, "zernike: bad argument");
    zernike_eval_kernel<<<(unsigned)((npts + 127) / 128), 128, 0, (cudaStream_t)stream>>>(r, theta, npts, n, m, nmodes,
                                                                                         norm, out);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

int pyotf_b200_prep_data_u16(const unsigned short* data, int nz, int ny, int nx, int xysize, double multiplier, int dtype,
                             void* out, int* mode_host, void* stream) {
    PYOTF_REQUIRE(data && out && nz > 0 && ny > 0 && nx > 0, "prep_data_u16: bad argument");
    PYOTF_REQUIRE(dtype == PYOTF_F32 || dtype == PYOTF_F64, "prep_data_u16: bad dtype");
    if (xysize <= 0) xysize = ny > nx ? ny : nx;                                  // utils.py:186-187
    cudaStream_t s = (cudaStream_t)stream;
    keep_scratch_pool();
    const long long n = (long long)nz * ny * nx;
    unsigned int* hist = nullptr;        // 65536 counts | argmax key (2 words) | mode
    PYOTF_CUDA_OK(cudaMallocAsync(&hist, (65536 + 4) * sizeof(unsigned int), s));
    PYOTF_CUDA_OK(cudaMemsetAsync(hist, 0, (65536 + 4) * sizeof(unsigned int), s));
    unsigned long long* key = reinterpret_cast<unsigned long long*>(hist + 65536);
    int* mode_dev = reinterpret_cast<int*>(hist + 65536 + 2);
    const int nb = (int)((n + 256 * 8 - 1) / (256 * 8));
    prep_hist_kernel<<<nb < 1 ? 1 : (nb > 2368 ? 2368 : nb), 256, 0, s>>>(data, n, hist, key);
    PYOTF_CUDA_OK(cudaGetLastError());
    prep_mode_kernel<<<1, 1024, 0, s>>>(hist, mode_dev);
    PYOTF_CUDA_OK(cudaGetLastError());
    struct { unsigned long long key; int mode; int pad; } h;
    PYOTF_CUDA_OK(cudaMemcpyAsync(&h, key, 16, cudaMemcpyDeviceToHost, s));
    PYOTF_CUDA_OK(cudaStreamSynchronize(s));
    PYOTF_CUDA_OK(cudaFreeAsync(hist, s));
    if (mode_host) *mode_host = h.mode;
    PrepGeom g = {};
    g.nz = nz; g.ny = ny; g.nx = nx; g.oy = g.ox = xysize;
    if (xysize == ny && xysize == nx) g.mode = 0;                                  // utils.py:190-191
    else if (xysize >= (ny > nx ? ny : nx)) {                                      // fft_pad   utils.py:192-194
        g.mode = 1; g.by = xysize / 2 - ny / 2; g.bx = xysize / 2 - nx / 2;
    } else {                                                                       // centre + crop  utils.py:195-199
        g.mode = 2;
        const long long flat = (long long)(0xFFFFFFFFFFull - (h.key & 0xFFFFFFFFFFull));
        const int pz = (int)(flat / ((long long)ny * nx)), py = (int)((flat / nx) % ny), px = (int)(flat % nx);
        g.sz = nz / 2 - pz; g.sy = ny / 2 - py; g.sx = nx / 2 - px;                 // center_data  utils.py:119-123
        // slice_maker(((ny + 1) // 2, (nx + 1) // 2), xysize): start = centre - width // 2, clipped at 0
        g.y0 = (ny + 1) / 2 - xysize / 2; if (g.y0 < 0) g.y0 = 0;
        g.x0 = (nx + 1) / 2 - xysize / 2; if (g.x0 < 0) g.x0 = 0;
    }
    const double bg = multiplier * (double)h.mode;
    const long long total = (long long)nz * xysize * xysize;
    const unsigned nblk = (unsigned)((total + 255) / 256);
    if (dtype == PYOTF_F32) prep_gather_kernel<float><<<nblk, 256, 0, s>>>(g, data, bg, (float*)out);
    else prep_gather_kernel<double><<<nblk, 256, 0, s>>>(g, data, bg, (double*)out);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

int pyotf_b200_zernike_gram(const double* Z, const double* b, long long npts, int nmodes, int nrhs, double* G, void* stream) {
    PYOTF_REQUIRE(Z && G && npts > 0 && nmodes > 0 && nrhs >= 0 && (nrhs == 0 || b), "zernike_gram: bad argument");
    PYOTF_REQUIRE(nmodes <= 65535 && nmodes + nrhs <= 65535, "zernike_gram: too many modes");
    dim3 grid((unsigned)(nmodes + nrhs), (unsigned)nmodes);
    zernike_gram_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(Z, b, npts, nmodes, nrhs, G);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

int pyotf_b200_hanser_aberration_pupil(const pyotf_hanser_t* h, const int* n, const int* m, int nmodes,
                                       const double* mc, const double* pc, int batch, void* out, void* stream) {
    PYOTF_REQUIRE(h && n && m && out, "aberration_pupil: null argument");
    PYOTF_REQUIRE(h->p.support == -2, "aberration_pupil: handle must be created with support == -2");
    PYOTF_REQUIRE(nmodes > 0 && nmodes <= 400 && batch > 0, "aberration_pupil: bad nmodes/batch");
    cudaStream_t s = (cudaStream_t)stream;
    return h->p.dtype == PYOTF_F32 ? aberration_launch<float>(h, n, m, nmodes, mc, pc, batch, out, s)
                                   : aberration_launch<double>(h, n, m, nmodes, mc, pc, batch, out, s);
}

int pyotf_b200_sheppard_otfa(const pyotf_sheppard_params* q, void* otfa, unsigned char* valid, void* stream) {
    PYOTF_REQUIRE(q && (otfa || valid), "sheppard_otfa: null argument");
    SheppardParams p;
    int rc = sheppard_params(q, &p);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (q->dtype == PYOTF_F32) sheppard_generate<float>(p, otfa, valid, nullptr, nullptr, s);
    else sheppard_generate<double>(p, otfa, valid, nullptr, nullptr, s);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

int pyotf_b200_sheppard_psf(const pyotf_sheppard_params* q, void* otfa_out, void* psfa_out, void* psfi_out, void* stream) {
    PYOTF_REQUIRE(q && (otfa_out || psfa_out || psfi_out), "sheppard_psf: no output requested");
    SheppardParams p;
    int rc = sheppard_params(q, &p);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t es = q->dtype == PYOTF_F32 ? 4 : 8;
    const size_t nvox = (size_t)p.NZ * p.N * p.N, nlines = (size_t)p.nvec * p.NZ * p.N;
    void *otfa = otfa_out, *psfa = psfa_out;
    unsigned char* flags = nullptr;         // rowflag [NZ*N] | planeflag [NZ] | f_x [nlines] | f_y [nlines]
    const bool want_fft = psfa_out || psfi_out;
    keep_scratch_pool();
    if (!otfa_out)       // OTFa not wanted: generate the shell inside the first FFT pass, fuse |.|^2 into the last
        return q->dtype == PYOTF_F32 ? sheppard_fused<float>(p, psfa_out, psfi_out, s)
                                     : sheppard_fused<double>(p, psfa_out, psfi_out, s);
    if (!otfa) PYOTF_CUDA_OK(cudaMallocAsync(&otfa, nvox * p.nvec * 2 * es, s));
    if (want_fft) {
        PYOTF_CUDA_OK(cudaMallocAsync(&flags, (size_t)p.NZ * p.N + p.NZ + 2 * nlines, s));
        PYOTF_CUDA_OK(cudaMemsetAsync(flags, 0, (size_t)p.NZ * p.N + p.NZ, s));
        if (!psfa) PYOTF_CUDA_OK(cudaMallocAsync(&psfa, nvox * p.nvec * 2 * es, s));
    }
    unsigned char* rowflag = flags, *planeflag = flags ? flags + (size_t)p.NZ * p.N : nullptr;
    if (q->dtype == PYOTF_F32) sheppard_generate<float>(p, otfa, nullptr, rowflag, planeflag, s);
    else sheppard_generate<double>(p, otfa, nullptr, rowflag, planeflag, s);
    PYOTF_CUDA_OK(cudaGetLastError());
    if (want_fft) {
        unsigned char* f_x = planeflag + p.NZ, *f_y = f_x + nlines;
        sheppard_flags_kernel<<<(unsigned)((nlines + 255) / 256), 256, 0, s>>>(p.nvec, p.NZ, p.N, rowflag, planeflag, f_x, f_y);
        PYOTF_CUDA_OK(cudaGetLastError());
        // PSFa = easy_ifft(OTFa, axes=(1, 2, 3))  (otf.py:589-592); passes run x, y, z
        const long long shape[4] = {p.nvec, p.NZ, p.N, p.N};
        const int axes[3] = {1, 2, 3};
        // the flags are exact only when every pass's lines coincide with the rows / planes they were built for,
        // i.e. for the contiguous x pass and the y pass; the z pass sees dense lines
        const unsigned char* pass_flags[3] = {f_x, f_y, nullptr};
        rc = q->dtype == PYOTF_F32 ? fftn_run<float>(otfa, psfa, 4, shape, axes, 3, 1, 0, 1, 1, s, pass_flags)
                                   : fftn_run<double>(otfa, psfa, 4, shape, axes, 3, 1, 0, 1, 1, s, pass_flags);
        if (!rc && psfi_out)
            rc = q->dtype == PYOTF_F32 ? abs2_sum0_run<float>(psfa, p.nvec, (long long)nvox, psfi_out, s)
                                       : abs2_sum0_run<double>(psfa, p.nvec, (long long)nvox, psfi_out, s);
        cudaFreeAsync(flags, s);
        if (!psfa_out) cudaFreeAsync(psfa, s);
    }
    if (!otfa_out) cudaFreeAsync(otfa, s);
    return rc;
}

int pyotf_b200_pr_create(const pyotf_hanser_params* params, const double* z_dev, int nz, long long nz_total,
                         const void* data_dev, int data_dtype, int rows_per_cta, void* stream, pyotf_pr_t** out) {
    PYOTF_REQUIRE(params && z_dev && data_dev && out, "pr_create: null argument");
    PYOTF_REQUIRE(nz > 0 && nz_total >= nz, "pr_create: bad nz / nz_total");
    PYOTF_REQUIRE(data_dtype == PYOTF_F32 || data_dtype == PYOTF_F64, "pr_create: bad data dtype");
    PYOTF_REQUIRE(params->size >= 2 && params->size <= 16384, "pr_create: size must be in 2..16384");
    pyotf_hanser_params p = *params;
    p.vec_corr = PYOTF_VEC_NONE; p.condition = PYOTF_COND_NONE; p.support = -2;   // phaseretrieval.py:74-76
    pyotf_pr* s = new pyotf_pr();
    int rc = pyotf_b200_hanser_create(&p, stream, &s->h);
    if (rc) { delete s; return rc; }
    s->nz = nz; s->nz_total = nz_total;
    cudaStream_t st = (cudaStream_t)stream;
    keep_scratch_pool();
    s->stream = st;
    cudaError_t e = cudaMallocAsync(&s->z, sizeof(double) * nz, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->z, z_dev, sizeof(double) * nz, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { set_error(std::string("pr_create: ") + cudaGetErrorString(e)); pyotf_b200_pr_destroy(s); return 2; }
    rc = p.dtype == PYOTF_F32 ? pr_setup<float>(s, data_dev, data_dtype, rows_per_cta, st)
                              : pr_setup<double>(s, data_dev, data_dtype, rows_per_cta, st);
    if (rc) { pyotf_b200_pr_destroy(s); return rc; }
    *out = s;
    return 0;
}

int pyotf_b200_pr_window_alloc(pyotf_pr_t* s, int world, unsigned char handle_out[64]) {
    PYOTF_REQUIRE(s && handle_out, "pr_window_alloc: null argument");
    PYOTF_REQUIRE(world >= 1 && world <= PYOTF_MAX_PEERS, "pr_window_alloc: bad world size");
    PYOTF_REQUIRE(!s->win, "pr_window_alloc: the state already has a window");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    const int KK = s->h->K * s->h->K;
    s->win_bytes = pr_win_bytes(KK, s->nblk, world);
    s->win_alloc_world = world;
    PYOTF_CUDA_OK(cudaMalloc(&s->win, s->win_bytes));            // exportable: not from the stream-ordered pool
    PYOTF_CUDA_OK(cudaMemset(s->win, 0, s->win_bytes));
    PYOTF_CUDA_OK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t hnd;
    PYOTF_CUDA_OK(cudaIpcGetMemHandle(&hnd, s->win));
    std::memcpy(handle_out, &hnd, 64);
    return 0;
}

int pyotf_b200_pr_window_open(pyotf_pr_t* s, const unsigned char* handles, int rank, int world) {
    PYOTF_REQUIRE(s && handles && s->win, "pr_window_open: call pr_window_alloc first");
    PYOTF_REQUIRE(world == s->win_alloc_world && rank >= 0 && rank < world, "pr_window_open: bad rank / world");
    PYOTF_REQUIRE(s->win_world == 0, "pr_window_open: the window is already open");
    for (int r = 0; r < world; ++r) {
        if (r == rank) { s->peer[r] = s->win; continue; }
        cudaIpcMemHandle_t hnd;
        std::memcpy(&hnd, handles + (size_t)r * 64, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&s->peer[r], hnd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; ++q) if (q != rank && s->peer[q]) { cudaIpcCloseMemHandle(s->peer[q]); s->peer[q] = nullptr; }
            set_error(std::string("pr_window_open: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
            return 2;
        }
    }
    s->win_rank = rank; s->win_world = world; s->epoch_base = 0;
    return 0;
}

int pyotf_b200_pr_destroy(pyotf_pr_t* s) {
    if (!s) return 0;
    cudaStream_t st = s->stream;
    if (s->win) {
        cudaStreamSynchronize(st);
        for (int r = 0; r < s->win_world; ++r) if (r != s->win_rank && s->peer[r]) cudaIpcCloseMemHandle(s->peer[r]);
        cudaFree(s->win);
    }
    void* bufs[] = {s->z, s->data, s->pupil[0], s->pupil[1], s->partial, s->mse_part, s->sums, s->diffpart, s->mse,
                    s->mse_diff, s->pupil_diff, s->state, s->dtab, s->W, s->P, s->arrive};
    for (void* b : bufs) if (b) cudaFreeAsync(b, st);         // stream-ordered: returns to the cached pool
    if (s->h_prog) {
        // the words may still be written by kernels of a run that failed half-way: drain before the slot is reused
        if (s->prog_slot >= 0) { if (s->in_flight) cudaStreamSynchronize(st); pyotf::ProgressSlots::get().release(s->prog_slot); }
        else cudaFreeHost(s->h_prog);
    }
    pyotf_b200_hanser_destroy(s->h);
    delete s;
    return 0;
}

int pyotf_b200_pr_info(const pyotf_pr_t* s, int* K, int* rows, int* f0, int* rows_per_cta, int* nparts) {
    PYOTF_REQUIRE(s, "pr_info: null handle");
    if (K) *K = s->h->K;
    if (rows) *rows = s->h->KR;
    if (f0) *f0 = s->h->f0;
    if (rows_per_cta) *rows_per_cta = s->M;
    if (nparts) *nparts = s->nparts;
    return 0;
}

int pyotf_b200_pr_kernels_per_iteration(const pyotf_pr_t* s, int sharded, int* n) {
    PYOTF_REQUIRE(s && n, "pr_kernels_per_iteration: null argument");
    if (s->generic) *n = 5;
    else if (sharded) *n = (s->win && s->win_world >= 1) ? 2 : 3;
    else *n = (s->fuse && s->dtab) ? 1 : 2;
    return 0;
}

int pyotf_b200_pr_debug_buffers(const pyotf_pr_t* s, void** W, void** P, void** partial) {
    PYOTF_REQUIRE(s, "pr_debug_buffers: null handle");
    if (W) *W = s->W;
    if (P) *P = s->P;
    if (partial) *partial = s->partial;
    return 0;
}

int pyotf_b200_pr_set_pupil(pyotf_pr_t* s, const void* pupil_c128_dev, void* stream) {
    PYOTF_REQUIRE(s, "pr_set_pupil: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    return s->h->p.dtype == PYOTF_F32 ? pr_set_pupil<float>(s, pupil_c128_dev, st) : pr_set_pupil<double>(s, pupil_c128_dev, st);
}

int pyotf_b200_pr_get_pupil(const pyotf_pr_t* s, int which, void* out, void* stream) {
    PYOTF_REQUIRE(s && out && (which == 0 || which == 1), "pr_get_pupil: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    return s->h->p.dtype == PYOTF_F32 ? pr_get_pupil<float>(s, which, out, st) : pr_get_pupil<double>(s, which, out, st);
}

int pyotf_b200_pr_run(pyotf_pr_t* s, int max_iters, double pupil_tol, double mse_tol, int phase_only,
                      pyotf_comm_t* comm, double* mse_host, double* mse_diff_host, double* pupil_diff_host,
                      int* iters_run_host, void* stream);
// the name SURVEY.md section 8(b) lists for the loop entry
int pyotf_b200_pr_iterate(pyotf_pr_t* s, int max_iters, double pupil_tol, double mse_tol, int phase_only,
                          pyotf_comm_t* comm, double* mse_host, double* mse_diff_host, double* pupil_diff_host,
                          int* iters_run_host, void* stream) {
    return pyotf_b200_pr_run(s, max_iters, pupil_tol, mse_tol, phase_only, comm, mse_host, mse_diff_host, pupil_diff_host,
                             iters_run_host, stream);
}

int pyotf_b200_pr_run(pyotf_pr_t* s, int max_iters, double pupil_tol, double mse_tol, int phase_only,
                      pyotf_comm_t* comm, double* mse_host, double* mse_diff_host, double* pupil_diff_host,
                      int* iters_run_host, void* stream) {
    PYOTF_REQUIRE(s, "pr_run: null handle");
    PYOTF_REQUIRE(max_iters > 0, "pr_run: Must have at least one iteration");
    cudaStream_t st = (cudaStream_t)stream;
    return s->h->p.dtype == PYOTF_F32
               ? pr_run<float>(s, max_iters, pupil_tol, mse_tol, phase_only, comm, mse_host, mse_diff_host,
                               pupil_diff_host, iters_run_host, st)
               : pr_run<double>(s, max_iters, pupil_tol, mse_tol, phase_only, comm, mse_host, mse_diff_host,
                                pupil_diff_host, iters_run_host, st);
}

int pyotf_b200_fftn(const void* in, void* out, int ndim, const long long* shape, const int* axes, int naxes, int inverse,
                    int real_in, int shift_in, int shift_out, int dtype, void* stream) {
    PYOTF_REQUIRE(in && out && shape && axes, "fftn: null argument");
    PYOTF_REQUIRE(ndim >= 1 && ndim <= 8 && naxes >= 1 && naxes <= ndim, "fftn: bad ndim / naxes");
    PYOTF_REQUIRE(dtype == PYOTF_F32 || dtype == PYOTF_F64, "fftn: bad dtype");
    PYOTF_REQUIRE(!(real_in && in == out), "fftn: real input cannot be transformed in place");
    unsigned seen = 0;
    for (int k = 0; k < naxes; ++k) {
        PYOTF_REQUIRE(axes[k] >= 0 && axes[k] < ndim && !(seen & (1u << axes[k])), "fftn: bad axes");
        seen |= 1u << axes[k];
    }
    for (int d = 0; d < ndim; ++d) PYOTF_REQUIRE(shape[d] > 0, "fftn: empty dimension");
    cudaStream_t s = (cudaStream_t)stream;
    return dtype == PYOTF_F32
               ? fftn_run<float>(in, out, ndim, shape, axes, naxes, inverse, real_in, shift_in, shift_out, s)
               : fftn_run<double>(in, out, ndim, shape, axes, naxes, inverse, real_in, shift_in, shift_out, s);
}

// the 3D convenience form SURVEY.md section 8(b) lists: easy_fft (direction < 0) / easy_ifft (direction > 0) of a
// [nz][ny][nx] array over all three axes (pyotf/utils.py:15-22; OTFi / OTFa / Sheppard PSFa)
int pyotf_b200_fft3d_shifted(const void* in, void* out, int nz, int ny, int nx, int direction, int real_in, int dtype, void* stream) {
    PYOTF_REQUIRE(nz > 0 && ny > 0 && nx > 0 && direction != 0, "fft3d_shifted: bad size / direction");
    const long long shape[3] = {nz, ny, nx};
    const int axes[3] = {0, 1, 2};
    return pyotf_b200_fftn(in, out, 3, shape, axes, 3, direction > 0 ? 1 : 0, real_in, 1, 1, dtype, stream);
}

int pyotf_b200_abs2_sum0(const void* a, int nvec, long long n, void* out, int dtype, void* stream) {
    PYOTF_REQUIRE(a && out && nvec > 0 && n > 0, "abs2_sum0: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    return dtype == PYOTF_F32 ? abs2_sum0_run<float>(a, nvec, n, out, s) : abs2_sum0_run<double>(a, nvec, n, out, s);
}

}  // extern "C"
