// k_cosx_l44.cu — instantiations of the COSX ESP kernel for shell pairs with a g shell (see k_cosx_esp.cuh)
#include "k_cosx_esp.cuh"
namespace xcb {
template void launch_class<4, 4>(const CosxParams &, int, int, int, cudaStream_t);
} // namespace xcb
