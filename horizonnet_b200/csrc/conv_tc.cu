// tcgen05 tensor-core convolution (split-bf16, 3 products) -- placeholder until the kernel lands.
#include "hn_common.cuh"
namespace hn {
bool conv_tc_supported(const ConvDesc&, const Act&, const Act&) { return false; }
int conv_tc(const ConvDesc&, const Act&, const Act&, const float*, cudaStream_t) {
    return fail("conv_tc: not built");
}
}  // namespace hn
