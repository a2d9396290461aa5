#include "sheppard_impl.cuh"
namespace pyotf {
template int sheppard_fused<double>(const SheppardParams&, void*, void*, cudaStream_t);
}  // namespace pyotf
