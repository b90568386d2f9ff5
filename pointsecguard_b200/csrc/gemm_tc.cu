// gemm_tc.cu -- tcgen05 / TMEM TF32 GEMM for the shared MLPs (placeholder until the kernel lands:
// reports "unsupported" so that callers fail loudly instead of silently changing precision).
#include "psg_common.cuh"
#include "psg_internal.h"
int psg_gemm_tc(const PsgGemmArgs &, cudaStream_t) { return PSG_EUNSUPPORTED; }
