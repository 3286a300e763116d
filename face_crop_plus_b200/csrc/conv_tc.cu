// conv_tc.cu — tcgen05 / TMA implicit-GEMM convolution (3xTF32 split).  Placeholder until the kernel lands.
#include "common.h"
namespace fcp {
int launch_conv_tc(fcp_ctx* ctx, const ConvOp&) { return fail(ctx, FCP_ERR_INVALID, "tcgen05 conv kernel not built"); }
}  // namespace fcp
