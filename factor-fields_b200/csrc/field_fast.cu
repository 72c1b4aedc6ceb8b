// field_fast.cu — specialised field-query kernels for the grid x grid shapes (placeholder: not eligible yet).
#include "ffb_common.cuh"

extern "C" {
int ffb_field_fast_eligible(ffb_field_t f) { (void)f; return 0; }
int ffb_field_fast_fwd(ffb_field_t, const float*, int64_t, const int32_t*, float*, float*, void*) {
  ffb::set_error("fast path not available");
  return FFB_EINVAL;
}
int ffb_field_fast_bwd(ffb_field_t, const float*, int64_t, const int32_t*, const float*, const float*, float* const*, void*) {
  ffb::set_error("fast path not available");
  return FFB_EINVAL;
}
}
