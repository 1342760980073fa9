// K2 (tensor-core engine) -- placeholder until the tcgen05 megakernel lands.
#include "field.cuh"

namespace neddf {
bool tc_supported(const neddf_field*) { return false; }
void tc_destroy(neddf_field*) {}
int32_t tc_pack_weights(neddf_field*, const float* const*, const float* const*, cudaStream_t) { return NEDDF_OK; }
int32_t launch_field_tc(const neddf_field*, FieldParams&, int, cudaStream_t) {
  return fail(NEDDF_E_UNSUPPORTED, "tensor-core engine not built");
}
}  // namespace neddf

extern "C" int32_t neddf_tc_selftest(const float*, const float*, int32_t, int32_t, int32_t, float*, void*) {
  return neddf::fail(NEDDF_E_UNSUPPORTED, "tensor-core engine not built");
}
