// tcgen05 implicit-GEMM convolution (placeholder until the kernel lands: reports "not available").
#include "conv_tc.cuh"

namespace ddn {
bool tc_available() { return false; }
bool tc_conv_supported(int, int, int, int, int, int, int, int) { return false; }
size_t tc_workspace_bytes(int, int, int) { return 0; }
int tc_conv_forward(const float*, const float*, float*, int, int, int, int, int, int, int, int, int, void*, size_t, cudaStream_t) {
  set_error("tcgen05 conv path not built");
  return DDN_EUNSUPPORTED;
}
int tc_conv_backward(const float*, const float*, const float*, float*, const float*, float*, int, int, int, int, int, int, int, int,
                     int, void*, size_t, float*, cudaStream_t) {
  set_error("tcgen05 conv path not built");
  return DDN_EUNSUPPORTED;
}
}  // namespace ddn
