// tcgen05 implicit-GEMM convolution (placeholder until the tensor-core path lands).
#include "layers.cuh"
namespace yb {
struct TcPlan {};
bool tc_supported(const ConvArgs&) { return false; }
int tc_plan_create(const ConvArgs&, int, TcPlan**) { set_error("tc path not built"); return YB_ERR_UNSUPPORTED; }
void tc_plan_destroy(TcPlan*) {}
int launch_conv_tc(const TcPlan*, const ConvArgs&, cudaStream_t) { set_error("tc path not built"); return YB_ERR_UNSUPPORTED; }
}  // namespace yb
