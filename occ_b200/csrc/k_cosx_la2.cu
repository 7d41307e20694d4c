// k_cosx_la2.cu — instantiations of the COSX ESP kernel for shell pairs with la = 2 (see k_cosx_esp.cuh)
#include "k_cosx_esp.cuh"
namespace xcb {
template void launch_class<2, 0>(const CosxParams &, int, int, int, cudaStream_t);
template void launch_class<2, 1>(const CosxParams &, int, int, int, cudaStream_t);
template void launch_class<2, 2>(const CosxParams &, int, int, int, cudaStream_t);
} // namespace xcb
