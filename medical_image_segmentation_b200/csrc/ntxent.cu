// placeholder -- replaced by the tcgen05 kernels
#include "common.cuh"
using namespace mis;
extern "C" int64_t mis_ntxent_scratch_bytes(int, int, int) { return 0; }
extern "C" int mis_ntxent_prep(const void*, int, int, int, float, float*, float*, float*, void*) {
  return set_error(MIS_ERR_UNSUPPORTED, "not built yet");
}
extern "C" int mis_ntxent_fwd(const float*, int, int, int, int, float, const float*, float*, float*, void*, int64_t, void*) {
  return set_error(MIS_ERR_UNSUPPORTED, "not built yet");
}
extern "C" int mis_ntxent_bwd(const float*, const float*, const float*, int, int, int, int, float, float, const float*, void*,
                              int, void*, int64_t, void*) {
  return set_error(MIS_ERR_UNSUPPORTED, "not built yet");
}
extern "C" int mis_byol_loss_fwd_bwd(const float*, const float*, int, int, float*, float*, void*) {
  return set_error(MIS_ERR_UNSUPPORTED, "not built yet");
}
