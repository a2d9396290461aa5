// N4: the first consumers of the hot path's output, on the device-resident PSFi.
//   pyotf_b200_bin_lateral    WidefieldMicroscope.PSF (pyotf/microscope.py:106-124): roll by oversample // 2 when the
//                             oversampled size is even, sum oversample x oversample lateral bins (dphtools bin_ndarray,
//                             un-vendored: block sum), normalise to unit sum.
//   pyotf_b200_disk_conv_mul  ConfocalMicroscope.model_psf (pyotf/microscope.py:180-187, 213-227): detection PSF
//                             convolved with the pinhole disk ("same" window of the linear convolution that
//                             scipy.signal.fftconvolve returns), times the excitation PSF.
//   pyotf_b200_wiener_otf     BaseSIMMicroscope.model_psf (pyotf/microscope.py:341-355): the Wiener-filtered OTF
//                             |OTFi|^2 / (|OTFi|^2 + max|OTFi|^2 wiener^2)  (wiener == 0: the support |OTFi|^2 > 1e-16).
//   pyotf_b200_sim_psf        BaseSIMMicroscope.model_psf (pyotf/microscope.py:316-401): the detection PSF times the
//                             excitation intensity of the two-beam (+ DC beam) interference patterns of every
//                             orientation, summed incoherently (2D / 3D SIM, optional DC suppression) or coherently
//                             (lattice SIM).  The patterns are evaluated per voxel in fp64; nothing but the PSF is read.
// All keep the stack in HBM: one read + one write of the stack (+ the excitation stack), fp64 accumulation.
#include "common.cuh"

namespace pyotf {

// out[z][Y][X] = sum_{dy, dx < f} in[z][(Y f + dy - shift) mod N][(X f + dx - shift) mod N]   (np.roll + block sum);
// part[block] = the block's share of the total (fixed order: the normalisation is deterministic)
template <typename T>
__global__ void __launch_bounds__(256) bin_lateral_kernel(const T* __restrict__ in, int nz, int N, int f, int shift,
                                                          T* __restrict__ out, double* __restrict__ part) {
    const int M = N / f;
    const long long total = (long long)nz * M * M;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    if (idx < total) {
        const int X = (int)(idx % M), Y = (int)((idx / M) % M);
        const long long z = idx / ((long long)M * M);
        const T* __restrict__ plane = in + z * (long long)N * N;
        for (int dy = 0; dy < f; ++dy) {
            int y = Y * f + dy - shift; y = y < 0 ? y + N : y;
            for (int dx = 0; dx < f; ++dx) {
                int x = X * f + dx - shift; x = x < 0 ? x + N : x;
                acc += (double)plane[(long long)y * N + x];
            }
        }
        out[idx] = (T)acc;
        acc = (double)(T)acc;                    // the total is the sum of what was stored (psf / psf.sum())
    }
    __shared__ double red[8];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        part[blockIdx.x] = s;
    }
}

// one CTA: total[0] = sum of the partials, fixed order
static __global__ void __launch_bounds__(1024) sum_partials_kernel(const double* __restrict__ part, int n, double* __restrict__ total) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) s += part[i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = red[threadIdx.x];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) total[0] = s;
    }
}

template <typename T>
__global__ void scale_by_total_kernel(T* __restrict__ a, long long n, const double* __restrict__ total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (T)((double)a[i] / total[0]);
}

// pinhole disk of pyotf/microscope.py:180-187: full_size = ceil(2 radius) made odd, kernel = (sqrt(i^2 + j^2) < radius)
// normalised to unit sum.  out = conv_same(em, disk) * exc; tile + halo staged in shared memory.
constexpr int DC_TILE = 32;
template <typename T>
__global__ void __launch_bounds__(DC_TILE * 8) disk_conv_mul_kernel(const T* __restrict__ em, const T* __restrict__ exc, int N, int c,
                                                                    double radius, double inv_count, T* __restrict__ out) {
    extern __shared__ unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    const int W = DC_TILE + 2 * c;
    const long long z = blockIdx.z;
    const int x0 = blockIdx.x * DC_TILE, y0 = blockIdx.y * DC_TILE;
    const T* __restrict__ plane = em + z * (long long)N * N;
    for (int i = threadIdx.x; i < W * W; i += blockDim.x) {
        const int ty = i / W, tx = i - ty * W;
        const int y = y0 + ty - c, x = x0 + tx - c;
        tile[i] = (y >= 0 && y < N && x >= 0 && x < N) ? plane[(long long)y * N + x] : (T)0;   // zero padding ("same")
    }
    __syncthreads();
    const int tx = threadIdx.x % DC_TILE;
    for (int ty = threadIdx.x / DC_TILE; ty < DC_TILE; ty += 8) {
        const int y = y0 + ty, x = x0 + tx;
        if (y >= N || x >= N) continue;
        double acc = 0.0;
        for (int i = -c; i <= c; ++i)
            for (int j = -c; j <= c; ++j)
                if (sqrt((double)(i * i + j * j)) < radius) acc += (double)tile[(ty + c - i) * W + (tx + c - j)];
        const long long o = z * (long long)N * N + (long long)y * N + x;
        out[o] = (T)(acc * inv_count * (double)exc[o]);
    }
}

// ---- SIM (pyotf/microscope.py:265-401) --------------------------------------------------------------------------
constexpr int SIM_MAX_ORIENT = 16;
struct SimParams {
    int nz, N, n_orient, coherent, dc, dc_suppress, psf_complex;
    double res, zres, freq, alpha;               // freq = 2 pi ni / wl_exc, alpha = arcsin(na_exc / ni)
    double orient[SIM_MAX_ORIENT];
};

// block maxima of |otf|^2 (fp64), then one CTA reduces them: part[nb] = max
template <typename T>
__global__ void __launch_bounds__(256) abs2_max_kernel(const T* __restrict__ otf /* interleaved complex */, long long n,
                                                       double* __restrict__ part) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double m = 0.0;
    if (i < n) { const double re = (double)otf[2 * i], im = (double)otf[2 * i + 1]; m = re * re + im * im; }
    __shared__ double red[8];
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) { for (int w = 1; w < 8; ++w) m = fmax(m, red[w]); part[blockIdx.x] = fmax(m, red[0]); }
}
static __global__ void __launch_bounds__(1024) max_partials_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
    __shared__ double red[32];
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) m = fmax(m, part[i]);
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = red[threadIdx.x];
        for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) out[0] = m;
    }
}
// otf = abs(OTFi) ** 2 ; wiener_otf = otf / (otf + otf.max() * wiener ** 2)  or  otf > 1e-16        microscope.py:343-351
template <typename T>
__global__ void __launch_bounds__(256) wiener_otf_kernel(const T* __restrict__ otf, long long n, double wiener,
                                                         const double* __restrict__ vmax, T* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double re = (double)otf[2 * i], im = (double)otf[2 * i + 1];
    const double a = re * re + im * im;                        // abs(OTFi) ** 2
    double v;
    if (wiener == 0.0) v = a > 1e-16 ? 1.0 : 0.0;
    else { const double w = vmax[0] * wiener * wiener; v = a / (a + w); }
    out[i] = (T)v;
}

template <typename T>
__global__ void __launch_bounds__(256) sim_psf_kernel(SimParams p, const T* __restrict__ psf, T* __restrict__ out) {
    const long long total = (long long)p.nz * p.N * p.N;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ix = (int)(idx % p.N), iy = (int)((idx / p.N) % p.N), iz = (int)(idx / ((long long)p.N * p.N));
    // centred coordinates: (arange(n) - (n + 1) // 2) * step                                          microscope.py:323-330
    const double x = (double)(ix - (p.N + 1) / 2) * p.res, y = (double)(iy - (p.N + 1) / 2) * p.res;
    const double z = (double)(iz - (p.nz + 1) / 2) * p.zres;
    const double v = (double)(p.psf_complex ? psf[2 * idx] : psf[idx]);      // easy_ifft(wiener_otf).real
    double bre = 0.0, bim = 0.0;
    if (p.dc) sincos(z * p.freq, &bim, &bre);                                // exp(1j * zz * freq)                     :362
    double acc = 0.0, ere = bre, eim = bim;
    for (int o = 0; o < p.n_orient; ++o) {
        if (!p.coherent) { ere = bre; eim = bim; }                           // base_pattern.copy()                     :373-376
        const double rr = x * cos(p.orient[o]) + y * sin(p.orient[o]);       // :379
        const double th[2] = {-p.alpha, p.alpha};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            double s, c;
            sincos((rr * sin(th[k]) + z * cos(th[k])) * p.freq, &s, &c);     // :386-387
            ere += c; eim += s;
        }
        if (!p.coherent) acc += v * (ere * ere + eim * eim);                 // :390-391
    }
    if (p.coherent) acc = v * (ere * ere + eim * eim);                       // :394-395
    else if (p.dc_suppress) acc -= (double)(2 * p.n_orient - 1) * v;         // :398-399
    out[idx] = (T)acc;
}

template <typename T> static int wiener_otf_run(const void* otf, long long n, double wiener, void* out, cudaStream_t s) {
    const unsigned nb = (unsigned)((n + 255) / 256);
    double* part = nullptr;
    keep_scratch_pool();
    PYOTF_CUDA_OK(cudaMallocAsync(&part, sizeof(double) * ((size_t)nb + 1), s));
    if (wiener != 0.0) {
        abs2_max_kernel<T><<<nb, 256, 0, s>>>((const T*)otf, n, part);
        max_partials_kernel<<<1, 1024, 0, s>>>(part, (int)nb, part + nb);
    }
    wiener_otf_kernel<T><<<nb, 256, 0, s>>>((const T*)otf, n, wiener, part + nb, (T*)out);
    PYOTF_CUDA_OK(cudaGetLastError());
    PYOTF_CUDA_OK(cudaFreeAsync(part, s));
    return 0;
}

template <typename T>
static int bin_lateral_run(const void* in, int nz, int N, int f, int shift, void* out, int normalise, cudaStream_t s) {
    const int M = N / f;
    const long long total = (long long)nz * M * M;
    const unsigned nb = (unsigned)((total + 255) / 256);
    double* part = nullptr;
    keep_scratch_pool();
    PYOTF_CUDA_OK(cudaMallocAsync(&part, sizeof(double) * ((size_t)nb + 1), s));
    bin_lateral_kernel<T><<<nb, 256, 0, s>>>((const T*)in, nz, N, f, shift, (T*)out, part);
    PYOTF_CUDA_OK(cudaGetLastError());
    if (normalise) {
        sum_partials_kernel<<<1, 1024, 0, s>>>(part, (int)nb, part + nb);
        scale_by_total_kernel<T><<<nb, 256, 0, s>>>((T*)out, total, part + nb);
        PYOTF_CUDA_OK(cudaGetLastError());
    }
    PYOTF_CUDA_OK(cudaFreeAsync(part, s));
    return 0;
}

template <typename T>
static int disk_conv_mul_run(const void* em, const void* exc, int nz, int N, double radius, void* out, cudaStream_t s) {
    int full = (int)ceil(radius * 2.0);
    if (full % 2 == 0) full += 1;
    const int c = (full - 1) / 2;
    long long count = 0;
    for (int i = -c; i <= c; ++i)
        for (int j = -c; j <= c; ++j)
            if (sqrt((double)(i * i + j * j)) < radius) ++count;
    PYOTF_REQUIRE(count > 0, "disk_conv_mul: empty pinhole kernel");
    const int W = DC_TILE + 2 * c;
    const size_t smem = sizeof(T) * (size_t)W * W;
    int smem_limit = 0, dev = 0;
    PYOTF_CUDA_OK(cudaGetDevice(&dev));
    PYOTF_CUDA_OK(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    PYOTF_REQUIRE(smem <= (size_t)smem_limit, "disk_conv_mul: pinhole too large for the tiled convolution");
    auto kern = disk_conv_mul_kernel<T>;
    PYOTF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((N + DC_TILE - 1) / DC_TILE), (unsigned)((N + DC_TILE - 1) / DC_TILE), (unsigned)nz);
    kern<<<grid, DC_TILE * 8, smem, s>>>((const T*)em, (const T*)exc, N, c, radius, 1.0 / (double)count, (T*)out);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace pyotf

using namespace pyotf;

extern "C" {

int pyotf_b200_bin_lateral(const void* in_dev, int nz, int size, int factor, int shift, void* out_dev, int normalise,
                           int dtype, void* stream) {
    PYOTF_REQUIRE(in_dev && out_dev, "bin_lateral: null argument");
    PYOTF_REQUIRE(nz > 0 && size > 0 && factor > 0 && size % factor == 0, "bin_lateral: size must be a multiple of the bin factor");
    PYOTF_REQUIRE(shift >= 0 && shift < size, "bin_lateral: bad shift");
    PYOTF_REQUIRE(dtype == PYOTF_F32 || dtype == PYOTF_F64, "bin_lateral: bad dtype");
    cudaStream_t s = (cudaStream_t)stream;
    return dtype == PYOTF_F32 ? bin_lateral_run<float>(in_dev, nz, size, factor, shift, out_dev, normalise, s)
                              : bin_lateral_run<double>(in_dev, nz, size, factor, shift, out_dev, normalise, s);
}

int pyotf_b200_disk_conv_mul(const void* em_dev, const void* exc_dev, int nz, int size, double radius, void* out_dev,
                             int dtype, void* stream) {
    PYOTF_REQUIRE(em_dev && exc_dev && out_dev, "disk_conv_mul: null argument");
    PYOTF_REQUIRE(nz > 0 && nz <= 65535 && size > 0 && radius > 0, "disk_conv_mul: bad size / radius");
    PYOTF_REQUIRE(dtype == PYOTF_F32 || dtype == PYOTF_F64, "disk_conv_mul: bad dtype");
    cudaStream_t s = (cudaStream_t)stream;
    return dtype == PYOTF_F32 ? disk_conv_mul_run<float>(em_dev, exc_dev, nz, size, radius, out_dev, s)
                              : disk_conv_mul_run<double>(em_dev, exc_dev, nz, size, radius, out_dev, s);
}

int pyotf_b200_wiener_otf(const void* otf_c_dev, long long n, double wiener, void* out_dev, int dtype, void* stream) {
    PYOTF_REQUIRE(otf_c_dev && out_dev && n > 0, "wiener_otf: bad argument");
    PYOTF_REQUIRE(wiener >= 0.0, "wiener_otf: the wiener parameter must not be negative");
    PYOTF_REQUIRE(dtype == PYOTF_F32 || dtype == PYOTF_F64, "wiener_otf: bad dtype");
    cudaStream_t s = (cudaStream_t)stream;
    return dtype == PYOTF_F32 ? wiener_otf_run<float>(otf_c_dev, n, wiener, out_dev, s) : wiener_otf_run<double>(otf_c_dev, n, wiener, out_dev, s);
}

int pyotf_b200_sim_psf(const void* psf_dev, int psf_is_complex, int nz, int size, double res, double zres, double ni,
                       double wl_exc, double na_exc, const double* orientations_host, int n_orient, int coherent, int dc,
                       int dc_suppress, void* out_dev, int dtype, void* stream) {
    PYOTF_REQUIRE(psf_dev && out_dev && orientations_host, "sim_psf: null argument");
    PYOTF_REQUIRE(nz > 0 && size > 0 && n_orient > 0 && n_orient <= SIM_MAX_ORIENT, "sim_psf: bad size / number of orientations (<= 16)");
    PYOTF_REQUIRE(na_exc > 0 && na_exc <= ni && wl_exc > 0, "sim_psf: bad excitation optics");
    PYOTF_REQUIRE(dtype == PYOTF_F32 || dtype == PYOTF_F64, "sim_psf: bad dtype");
    SimParams p;
    p.nz = nz; p.N = size; p.n_orient = n_orient; p.coherent = coherent; p.dc = dc; p.dc_suppress = dc_suppress;
    p.psf_complex = psf_is_complex; p.res = res; p.zres = zres;
    p.freq = 2 * 3.14159265358979323846 * ni / wl_exc;                     // microscope.py:335
    p.alpha = asin(na_exc / ni);                                            // :338
    for (int i = 0; i < n_orient; ++i) p.orient[i] = orientations_host[i];
    const long long total = (long long)nz * size * size;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned nb = (unsigned)((total + 255) / 256);
    if (dtype == PYOTF_F32) sim_psf_kernel<float><<<nb, 256, 0, s>>>(p, (const float*)psf_dev, (float*)out_dev);
    else sim_psf_kernel<double><<<nb, 256, 0, s>>>(p, (const double*)psf_dev, (double*)out_dev);
    PYOTF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
