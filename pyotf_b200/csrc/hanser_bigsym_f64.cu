#include "hanser_bigsym.cuh"
namespace pyotf {
template int hanser_bigsym_launch<double>(const pyotf_hanser*, const double*, int, int, void*, int, cudaStream_t, bool*);
template int hanser_biggen_launch<double>(const pyotf_hanser*, const double*, int, const void*, int, long long, void*, int,
                                       cudaStream_t, bool*);
}  // namespace pyotf
