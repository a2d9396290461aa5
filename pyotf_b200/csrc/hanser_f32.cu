#include "hanser_impl.cuh"
namespace pyotf {
template int hanser_build_tables<float>(pyotf_hanser*, cudaStream_t);
template int hanser_pack<float>(const pyotf_hanser*, const void*, int, void*, cudaStream_t);
template int hanser_launch<float>(const pyotf_hanser*, const double*, int, const void*, int, long long, void*, void*,
                                  cudaStream_t, const double*);
template int hanser_launch_tables<float>(int, int, const void*, int, void*, void*, cudaStream_t, const void*);
}  // namespace pyotf

#ifdef PYOTF_K1_TIMING
extern "C" int pyotf_b200_debug_k1_marks(unsigned long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, pyotf::g_k1_marks, sizeof(unsigned long long) * n);
}
#endif
