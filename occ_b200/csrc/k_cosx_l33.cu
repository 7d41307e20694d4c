// k_cosx_l33.cu — instantiation of the COSX ESP kernel for f-f shell pairs (see k_cosx_esp.cuh)
#include "k_cosx_esp.cuh"
namespace xcb {
template void launch_class<3, 3>(const CosxParams &, int, int, int, cudaStream_t);
} // namespace xcb
