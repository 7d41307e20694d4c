// k_cosx_la0.cu — instantiations of the COSX ESP kernel for shell pairs with la = 0 (see k_cosx_esp.cuh)
#include "k_cosx_esp.cuh"
namespace xcb {
template void launch_class<0, 0>(const CosxParams &, int, int, int, cudaStream_t);
} // namespace xcb
