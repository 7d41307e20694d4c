// k_functional_f0s1.cu — instantiation of the functional kernels for family 0 (LDA), restricted (see k_functional_impl.cuh)
#include "k_functional_impl.cuh"
namespace xcb {
template void launch_functional_t<0, 1>(const BuildParams &, int, int, cudaStream_t);
template void launch_eval_points_t<0, false>(unsigned, int, int, size_t, const double *, const double *, const double *, double *,
                                              double *, double *, double *);
} // namespace xcb
