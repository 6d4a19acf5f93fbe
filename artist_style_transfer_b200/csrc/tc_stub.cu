// Placeholder until conv_tc.cu / gram_tc.cu land: requesting the tensor-core path is an error, never a fallback.
#include "common.cuh"
namespace ast {
int conv_gather_tc(const ast_image*, const void*, const float*, const float*, const ast_image*, const ast_image*,
                   const ast_image*, const ast_gather_geom*, cudaStream_t) {
  set_error("tcgen05 conv kernel not built into this library");
  return -2;
}
int gram_tc(const ast_image*, float*, float, cudaStream_t) {
  set_error("tcgen05 gram kernel not built into this library");
  return -2;
}
int tc_capabilities() { return 0; }
}  // namespace ast
