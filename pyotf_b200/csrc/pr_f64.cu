#include "hanser_impl.cuh"
#include "pr_impl.cuh"
namespace pyotf {
template int pr_setup<double>(pyotf_pr*, const void*, int, int, cudaStream_t);
template int pr_set_pupil<double>(pyotf_pr*, const void*, cudaStream_t);
template int pr_get_pupil<double>(const pyotf_pr*, int, void*, cudaStream_t);
template int pr_run<double>(pyotf_pr*, int, double, double, int, pyotf_comm*, double*, double*, double*, int*, cudaStream_t);
}  // namespace pyotf
