// extern "C" boundary: argument validation, handle lifetime, dtype dispatch.
#include "common.cuh"
#include "hanser.cuh"

namespace pyotf {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace pyotf

using namespace pyotf;

extern "C" {

int pyotf_b200_abi_version(void) { return PYOTF_B200_ABI_VERSION; }
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
    PYOTF_REQUIRE(is_pow2(p->size) && p->size >= 32, "hanser_create: size must be a power of two >= 32");
    PYOTF_REQUIRE(p->vec_corr >= 0 && p->vec_corr <= 4, "hanser_create: bad vec_corr");
    PYOTF_REQUIRE(p->condition >= 0 && p->condition <= 2, "hanser_create: bad condition");
    PYOTF_REQUIRE(p->dtype == PYOTF_F32 || p->dtype == PYOTF_F64, "hanser_create: bad dtype");
    pyotf_hanser* h = new pyotf_hanser();
    h->p = *p;
    h->N = p->size;
    PYOTF_CUDA_OK(cudaGetDevice(&h->device));
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
    size_t KK = (size_t)h->K * h->K, es = p->dtype == PYOTF_F32 ? 4 : 8;
    cudaStream_t s = (cudaStream_t)stream;
    PYOTF_CUDA_OK(cudaMalloc(&h->kzc, KK * sizeof(double)));
    PYOTF_CUDA_OK(cudaMalloc(&h->ampc, KK * h->nvec * es));
    PYOTF_CUDA_OK(cudaMalloc(&h->maskc, KK));
    PYOTF_CUDA_OK(cudaMalloc(&h->tw, (size_t)h->N * 2 * es));
    int rc = p->dtype == PYOTF_F32 ? hanser_build_tables<float>(h, s) : hanser_build_tables<double>(h, s);
    if (rc) { pyotf_b200_hanser_destroy(h); return rc; }
    *out = h;
    return 0;
}

int pyotf_b200_hanser_destroy(pyotf_hanser_t* h) {
    if (!h) return 0;
    cudaFree(h->kzc); cudaFree(h->ampc); cudaFree(h->maskc); cudaFree(h->tw);
    delete h;
    return 0;
}

int pyotf_b200_hanser_info(const pyotf_hanser_t* h, int* K, int* f0, int* nvec) {
    PYOTF_REQUIRE(h, "hanser_info: null handle");
    if (K) *K = h->K;
    if (f0) *f0 = h->f0;
    if (nvec) *nvec = h->nvec;
    return 0;
}

int pyotf_b200_hanser_tables(const pyotf_hanser_t* h, double* kz_dev, void* amp_dev, unsigned char* mask_dev,
                             void* stream) {
    PYOTF_REQUIRE(h, "hanser_tables: null handle");
    cudaStream_t s = (cudaStream_t)stream;
    size_t KK = (size_t)h->K * h->K, es = h->p.dtype == PYOTF_F32 ? 4 : 8;
    if (kz_dev) PYOTF_CUDA_OK(cudaMemcpyAsync(kz_dev, h->kzc, KK * 8, cudaMemcpyDeviceToDevice, s));
    if (amp_dev) PYOTF_CUDA_OK(cudaMemcpyAsync(amp_dev, h->ampc, KK * h->nvec * es, cudaMemcpyDeviceToDevice, s));
    if (mask_dev) PYOTF_CUDA_OK(cudaMemcpyAsync(mask_dev, h->maskc, KK, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int pyotf_b200_pupil_support(const void* pupil, int size, int* hmax_host, void* stream) {
    PYOTF_REQUIRE(pupil && hmax_host, "pupil_support: null argument");
    PYOTF_REQUIRE(is_pow2(size), "pupil_support: size must be a power of two");
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
    PYOTF_REQUIRE((h->p.support < 0) == (pupilc_dev == nullptr),
                  "hanser_psf: handle support mode and pupil argument disagree");
    PYOTF_REQUIRE(batch <= 65535, "hanser_psf: batch too large for one launch");
    cudaStream_t s = (cudaStream_t)stream;
    return h->p.dtype == PYOTF_F32
               ? hanser_launch<float>(h, z_dev, nz, pupilc_dev, batch, pupil_stride, psfa_out_dev, psfi_out_dev, s)
               : hanser_launch<double>(h, z_dev, nz, pupilc_dev, batch, pupil_stride, psfa_out_dev, psfi_out_dev, s);
}

}  // extern "C"
