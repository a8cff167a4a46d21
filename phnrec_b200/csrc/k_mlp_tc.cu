// k_mlp_tc.cu — placeholder until the tcgen05 kernels land (next commit).
#include "internal.h"
namespace phn {
int mlp_tc_prepare(phn_ctx *c) { return fail(c, PHN_ERR_UNSUPPORTED, "tensor-core MLP mode not built\n"); }
int launch_mlp_tc(phn_ctx *c, int64_t, int64_t) { return fail(c, PHN_ERR_UNSUPPORTED, "tensor-core MLP mode not built\n"); }
}  // namespace phn
