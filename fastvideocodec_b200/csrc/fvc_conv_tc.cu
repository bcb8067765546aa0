// placeholder until the tcgen05 engine lands (next commit)
#include "fvc_kernels.cuh"
namespace fvc {
struct TcPlan { int dummy; };
bool tc_supported(const ConvLayer&, int) { return false; }
int tc_plan_create(const ConvLayer&, const float*, ActT, int, int, const Epilogue&, TcPlan**, cudaStream_t) {
    set_error("tcgen05 engine not built");
    return FVC_ERR_STATE;
}
int tc_plan_launch(TcPlan*, cudaStream_t) { return FVC_ERR_STATE; }
void tc_plan_destroy(TcPlan* p) { delete p; }
}  // namespace fvc
