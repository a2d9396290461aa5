// N4: the first consumers of the hot path's output, on the device-resident PSFi.
//   pyotf_b200_bin_lateral    WidefieldMicroscope.PSF (pyotf/microscope.py:106-124): roll by oversample // 2 when the
//                             oversampled size is even, sum oversample x oversample lateral bins (dphtools bin_ndarray,
//                             un-vendored: block sum), normalise to unit sum.
//   pyotf_b200_disk_conv_mul  ConfocalMicroscope.model_psf (pyotf/microscope.py:180-187, 213-227): detection PSF
//                             convolved with the pinhole disk ("same" window of the linear convolution that
//                             scipy.signal.fftconvolve returns), times the excitation PSF.
// Both keep the stack in HBM: one read + one write of the stack (+ the excitation stack), fp64 accumulation.
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

}  // extern "C"
