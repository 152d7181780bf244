// gemm_tc.cu — tcgen05 implicit-GEMM back end of conv_gemm() (placeholder until the kernel lands).
#include "ops.h"
bool gemm_tc_supported(const ConvGemm&) { return false; }
void gemm_tc_launch(idx_engine*, const ConvGemm&) { throw IdxError(IDX_ERR_STATE, "tcgen05 GEMM not built"); }
