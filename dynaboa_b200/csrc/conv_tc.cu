// tcgen05 (5th-gen tensor core) path for GEMM-shaped convolutions -- see DESIGN.md "Tensor-core conv".
// Round-1 status: the dispatch hook and the enable switch exist; the kernel body lands after the fp32
// path is parity-green on the GPU (it is validated against conv.cu on-device).
#include "common.cuh"
#include "kernels.h"

namespace dboa {

static bool g_tc_enabled = false;
bool conv_tc_enabled() { return g_tc_enabled; }
void conv_tc_set_enabled(bool on) { g_tc_enabled = on; }

int conv1x1_tc_fwd(const float*, const float*, float*, int, int, int, cudaStream_t) { return DBOA_ERR_UNSUPPORTED; }

}  // namespace dboa
