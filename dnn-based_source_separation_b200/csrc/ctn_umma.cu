// tcgen05 (sm_100a) pointwise contraction path -- placeholder until the kernels land.
#include "ctn_internal.h"
extern "C" int ctn_has_tcgen05(void) { return 0; }
int ctn_pw_umma(const PwArgs&, int, int, int, cudaStream_t) { return CTN_ENOTBUILT; }
