#include "hanser_impl.cuh"
namespace pyotf {
template int hanser_build_tables<double>(pyotf_hanser*, cudaStream_t);
template int hanser_pack<double>(const pyotf_hanser*, const void*, int, void*, cudaStream_t);
template int hanser_launch<double>(const pyotf_hanser*, const double*, int, const void*, int, long long, void*, void*,
                                  cudaStream_t, const double*);
template int hanser_launch_tables<double>(int, int, const void*, int, void*, void*, cudaStream_t, const void*);
}  // namespace pyotf
