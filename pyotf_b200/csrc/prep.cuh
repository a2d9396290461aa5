// N3: data preparation for phase retrieval on the device (pyotf/utils.py:76-201): mode-based background
// removal (remove_bg :127-147), centring on the maximum (center_data :76-124), pad / crop to xysize and
// clipping at zero (prep_data_for_PR :161-201) for unsigned 16-bit camera stacks.
//   prep_hist_kernel   : 65536-bin histogram (shared-memory privatised per CTA) + packed arg-max key
//   prep_mode_kernel   : mode = first index of the histogram maximum (numpy.bincount(...).argmax())
//   prep_gather_kernel : out[z][y][x] = max(0, data[src] - multiplier * mode), src = rolled / shifted index,
//                        zero outside the source (padding)
#pragma once
#include <cuda_runtime.h>

namespace pyotf {

// histogram of uint16 values; key = max over elements of (value << 40 | (2^40 - 1 - flat index)): the largest
// value, ties broken towards the smallest flat index like numpy.argmax
static __global__ void __launch_bounds__(256) prep_hist_kernel(const unsigned short* __restrict__ data, long long n,
                                                               unsigned int* __restrict__ hist,
                                                               unsigned long long* __restrict__ argmax_key) {
    unsigned long long best = 0ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned int v = data[i];
        // camera stacks pile up on the background value: lanes holding the same value add once per warp
        const unsigned act = __activemask();
        const unsigned peers = __match_any_sync(act, v);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[v], (unsigned)__popc(peers));
        const unsigned long long key = ((unsigned long long)v << 40) | (0xFFFFFFFFFFull - (unsigned long long)i);
        best = key > best ? key : best;
    }
    for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(argmax_key, best);
}

// out[0] = mode (first index of the maximum count), single CTA of 1024 threads
static __global__ void __launch_bounds__(1024) prep_mode_kernel(const unsigned int* __restrict__ hist, int* __restrict__ out) {
    __shared__ unsigned long long red[32];
    unsigned long long best = 0ull;
    for (int b = threadIdx.x; b < 65536; b += blockDim.x) {
        const unsigned long long key = ((unsigned long long)hist[b] << 32) | (unsigned long long)(0xFFFFu - b);
        best = key > best ? key : best;
    }
    for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w) best = red[w] > best ? red[w] : best;
        best = red[0] > best ? red[0] : best;
        out[0] = 0xFFFF - (int)(best & 0xFFFFull);
    }
}

struct PrepGeom {
    int nz, ny, nx;          // source shape
    int oy, ox;              // output lateral shape (nz is kept)
    int mode;                // 0 copy, 1 pad (offsets by, bx), 2 centre + crop (roll sz, sy, sx; window y0, x0)
    int by, bx;
    int sz, sy, sx, y0, x0;
};

template <typename T>
__global__ void prep_gather_kernel(PrepGeom g, const unsigned short* __restrict__ data, double bg, T* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)g.nz * g.oy * g.ox;
    if (idx >= total) return;
    const int x = (int)(idx % g.ox), y = (int)((idx / g.ox) % g.oy), z = (int)(idx / ((long long)g.ox * g.oy));
    int szz = z, syy = y, sxx = x;
    bool inside = true;
    if (g.mode == 1) {                                   // fft_pad: source element i sits at i + before
        syy = y - g.by; sxx = x - g.bx;
        inside = syy >= 0 && syy < g.ny && sxx >= 0 && sxx < g.nx;
    } else if (g.mode == 2) {                            // np.roll by (sz, sy, sx), then the window [y0, y0+oy) x [x0, x0+ox)
        szz = ((z - g.sz) % g.nz + g.nz) % g.nz;
        syy = ((y + g.y0 - g.sy) % g.ny + g.ny) % g.ny;
        sxx = ((x + g.x0 - g.sx) % g.nx + g.nx) % g.nx;
    }
    double v = 0.0;
    if (inside) v = (double)data[((long long)szz * g.ny + syy) * g.nx + sxx] - bg;    // remove_bg  utils.py:147
    out[idx] = (T)fmax(0.0, v);                          // np.fmax(0, ...)  utils.py:201
}

}  // namespace pyotf
