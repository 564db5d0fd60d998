// Entry points declared in awx.h whose kernels land in later commits of this round.
// They fail loudly (AWX_E_UNSUPPORTED); nothing here computes anything.
#include "awx_internal.cuh"
using namespace awx;

extern "C" size_t awx_corrupt_workspace_bytes(int64_t, int32_t, int32_t) { return 0; }
extern "C" int awx_corrupt(const uint8_t*, uint8_t*, int64_t, int32_t, int32_t, const AwxCorruptParams*, const void*,
                           int32_t, const int32_t*, int64_t, void*, void*) {
  set_error("awx_corrupt: not implemented yet");
  return AWX_E_UNSUPPORTED;
}
extern "C" int awx_synth_depth(const double*, void*, int32_t, double*, int64_t, int32_t, int32_t, double, void*) {
  set_error("awx_synth_depth: not implemented yet");
  return AWX_E_UNSUPPORTED;
}
extern "C" int awx_fogloss(const float*, const void*, int32_t, const float*, const float*, const float*, float, int32_t,
                           int64_t, int32_t, int64_t, double*, float*, float*, int64_t*, void*) {
  set_error("awx_fogloss: not implemented yet");
  return AWX_E_UNSUPPORTED;
}
extern "C" int awx_scale_inplace(float*, int64_t, const float*, void*) {
  set_error("awx_scale_inplace: not implemented yet");
  return AWX_E_UNSUPPORTED;
}
