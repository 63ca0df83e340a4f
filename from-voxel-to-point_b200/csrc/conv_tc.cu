// Tensor-core (tcgen05) sparse convolution -- placeholder until the UMMA kernel lands.
#include "common.cuh"

namespace fv2p {
int launch_conv_tc(const void *, const void *, const int *, int64_t, int, int64_t, const int *, int, int,
                   const float *, const float *, const float *, const void *, int, int, void *, cudaStream_t) {
  set_error("conv_fwd: tensor-core modes are not built yet");
  return FV2P_ERR_UNSUPPORTED;
}
}  // namespace fv2p

extern "C" size_t fv2p_pack_weight_bytes(int, int, int, int) { return 0; }
extern "C" int fv2p_pack_weight(const float *, int, int, int, int, void *, fv2p_stream_t) {
  fv2p::set_error("pack_weight: tensor-core modes are not built yet");
  return FV2P_ERR_UNSUPPORTED;
}
