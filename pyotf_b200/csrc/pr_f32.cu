#include "hanser_impl.cuh"
#include "pr_impl.cuh"
namespace pyotf {
template int pr_setup<float>(pyotf_pr*, const void*, int, int, cudaStream_t);
template int pr_set_pupil<float>(pyotf_pr*, const void*, cudaStream_t);
template int pr_get_pupil<float>(const pyotf_pr*, int, void*, cudaStream_t);
template int pr_run<float>(pyotf_pr*, int, double, double, int, pyotf_comm*, double*, double*, double*, int*, cudaStream_t);
}  // namespace pyotf

#ifdef PYOTF_PR_TIMING
extern "C" int pyotf_b200_debug_pr_marks(unsigned long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, pyotf::g_pr_marks, sizeof(unsigned long long) * n);
}
#endif
