// placeholder until the tcgen05 kernel lands (next commit)
#include "common.cuh"
extern "C" size_t nerfb_packed_weights_bytes(int) { return 16; }
extern "C" int nerfb_pack_weights(const float* const*, int, int, void*, void*) {
  nerfb::set_error("bf16 tensor-core path not built yet");
  return NERFB_ERR_ARG;
}
extern "C" int nerfb_mlp_fwd_bf16(int, int, const float*, int, const float*, const float*, const float*, int,
                                  const void*, float*, void*) {
  nerfb::set_error("bf16 tensor-core path not built yet");
  return NERFB_ERR_ARG;
}
