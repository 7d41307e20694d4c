// k_functional_f1s1.cu — instantiation of the functional kernels for family 1 (GGA), restricted (see k_functional_impl.cuh)
#include "k_functional_impl.cuh"
namespace xcb {
template void launch_functional_t<1, 1>(const BuildParams &, int, int, cudaStream_t);
template void launch_eval_points_t<1, false>(unsigned, int, int, size_t, const double *, const double *, const double *, double *,
                                              double *, double *, double *);
} // namespace xcb
