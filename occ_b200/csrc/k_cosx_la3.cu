// k_cosx_la3.cu — instantiations of the COSX ESP kernel for shell pairs with la = 3, lb < 3 (see k_cosx_esp.cuh)
#include "k_cosx_esp.cuh"
namespace xcb {
template void launch_class<3, 0>(const CosxParams &, int, int, int, cudaStream_t);
template void launch_class<3, 1>(const CosxParams &, int, int, int, cudaStream_t);
template void launch_class<3, 2>(const CosxParams &, int, int, int, cudaStream_t);
} // namespace xcb
