// placeholder until the tcgen05 back end lands
#include "vgg_common.cuh"
namespace ha {
struct Arena;
int vgg_forward_tc(const char*, const PackedLayout&, const float*, int, int, int, int, int, float* const*, Arena&, cudaStream_t) {
  return HA_EUNSUPPORTED;
}
}  // namespace ha
