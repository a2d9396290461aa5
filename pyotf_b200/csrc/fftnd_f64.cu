#include "fftnd_impl.cuh"
namespace pyotf {
template int fftn_run<double>(const void*, void*, int, const long long*, const int*, int, int, int, int, int, cudaStream_t,
                             const unsigned char* const*);
template int abs2_sum0_run<double>(const void*, int, long long, void*, cudaStream_t);
int fft_twiddles_f64(int L, const void** out, cudaStream_t s) { return fft_twiddles<double>(L, out, s); }
}  // namespace pyotf
