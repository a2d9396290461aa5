#include "hanser_sym.cuh"
namespace pyotf {
template int hanser_sym_launch<float>(const HanserArgs<float>&, int, int, cudaStream_t, bool*);
}  // namespace pyotf

#ifdef PYOTF_K1_TIMING
extern "C" int pyotf_b200_debug_sym_marks(unsigned long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, pyotf::g_sym_marks, sizeof(unsigned long long) * n);
}
#endif
