// Bucketed tensor-core (tcgen05) TT path -- placeholder until the kernels land.
#include "ttb_common.cuh"

namespace ttb {
bool fast_supported(const ChainDims&) { return false; }
size_t fast_workspace_bytes(const ChainDims&, int64_t) { return 0; }
int launch_fwd_fast(const ChainDims&, int64_t, const int64_t*, const int64_t*, const int64_t*,
                    const CorePtrs&, float*, void*, size_t, cudaStream_t) {
  set_error("fast path not built");
  return 1;
}
int launch_bwd_fast(const ChainDims&, int64_t, const int64_t*, const int64_t*, const int64_t*,
                    const float*, const CorePtrs&, const CorePtrsRW&, void*, size_t, cudaStream_t) {
  set_error("fast path not built");
  return 1;
}
}  // namespace ttb
