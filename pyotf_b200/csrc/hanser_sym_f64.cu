#include "hanser_sym.cuh"
namespace pyotf {
template int hanser_sym_launch<double>(const HanserArgs<double>&, int, int, cudaStream_t, bool*);
}  // namespace pyotf
