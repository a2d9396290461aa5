#include "fftnd_impl.cuh"
namespace pyotf {
template int fftn_run<float>(const void*, void*, int, const long long*, const int*, int, int, int, int, int, cudaStream_t,
                             const unsigned char* const*);
template int abs2_sum0_run<float>(const void*, int, long long, void*, cudaStream_t);
int fft_twiddles_f32(int L, const void** out, cudaStream_t s) { return fft_twiddles<float>(L, out, s); }
}  // namespace pyotf
