// k_mo_det.cu — the order-independent (options.deterministic) build of the kernels in k_mo.cu: same source, every
// cross-CTA accumulation on two exactly-summable bins (red_add_binned, gemm_common.cuh).  A separate translation unit
// so that the default kernels keep their tuned ptxas schedule.
#define XC_DETERMINISTIC 1
#include "k_mo.cu"
