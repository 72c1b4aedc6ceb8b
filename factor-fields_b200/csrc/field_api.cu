// field_api.cu — public field-query entry points: dispatch between the specialised grid x grid kernels
// (field_fast.cu) and the descriptor-driven generic kernels (field_generic.cu).
#include "ffb_common.cuh"

extern "C" {
int ffb_field_generic_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff,
                          float* basis_out, void* stream);
int ffb_field_generic_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats,
                          const float* g_coeff, float* const* h_grads, void* stream);
int ffb_field_fast_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff, void* stream);
int ffb_field_fast_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats, const float* g_coeff,
                       float* const* h_grads, void* stream);

int ffb_field_fast_fwd_train(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff, float* basis,
                             void* stream);
int ffb_field_fast_bwd_saved(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats, const float* g_coeff,
                             const float* coeff, const float* basis, float* const* h_grads, void* stream);

// vector x lines (CP) fields: shared-memory resident factors (field_lines.cu)
int ffb_field_lines_eligible(ffb_field_t f);
int ffb_field_lines_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff, float* basis, void* stream);
int ffb_field_lines_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats, const float* g_coeff,
                        float* const* h_grads, void* stream);

// vector-matrix (vm) fields: coefficient lines x per-level plane triples (field_planes.cu)
int ffb_field_planes_eligible(ffb_field_t f);
int ffb_field_planes_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff, float* basis, void* stream);
int ffb_field_planes_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats, const float* g_coeff,
                         float* const* h_grads, void* stream);
int ffb_field_planes_bwd_saved(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats, const float* g_coeff,
                               const float* coeff, const float* basis, float* const* h_grads, void* stream);

// Training forward: also writes the concatenated basis row (needed by ffb_field_query_bwd_saved).
int ffb_field_query_fwd_train(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff, float* basis,
                              void* stream) {
  if (ffb_field_fast_eligible(f) == 1) return ffb_field_fast_fwd_train(f, x, n, n_dev, feats, coeff, basis, stream);
  if (ffb_field_lines_eligible(f) == 1) return ffb_field_lines_fwd(f, x, n, n_dev, feats, coeff, basis, stream);
  if (ffb_field_planes_eligible(f) == 1) return ffb_field_planes_fwd(f, x, n, n_dev, feats, coeff, basis, stream);
  return ffb_field_generic_fwd(f, x, n, n_dev, feats, coeff, basis, stream);
}

// Backward from the saved coefficient / basis rows (no re-gather) where the specialised kernels apply; otherwise
// identical to ffb_field_query_bwd.
int ffb_field_query_bwd_saved(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats, const float* g_coeff,
                              const float* coeff, const float* basis, float* const* h_grads, void* stream) {
  if (ffb::g_deterministic) return ffb_field_generic_bwd(f, x, n, n_dev, g_feats, g_coeff, h_grads, stream);
  if (ffb_field_fast_eligible(f) == 1) return ffb_field_fast_bwd_saved(f, x, n, n_dev, g_feats, g_coeff, coeff, basis, h_grads, stream);
  if (ffb_field_lines_eligible(f) == 1) return ffb_field_lines_bwd(f, x, n, n_dev, g_feats, g_coeff, h_grads, stream);
  if (ffb_field_planes_eligible(f) == 1) return ffb_field_planes_bwd_saved(f, x, n, n_dev, g_feats, g_coeff, coeff, basis, h_grads, stream);
  return ffb_field_generic_bwd(f, x, n, n_dev, g_feats, g_coeff, h_grads, stream);
}

int ffb_field_query_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff, void* stream) {
  if (ffb_field_fast_eligible(f) == 1) return ffb_field_fast_fwd(f, x, n, n_dev, feats, coeff, stream);
  if (ffb_field_lines_eligible(f) == 1) return ffb_field_lines_fwd(f, x, n, n_dev, feats, coeff, nullptr, stream);
  if (ffb_field_planes_eligible(f) == 1) return ffb_field_planes_fwd(f, x, n, n_dev, feats, coeff, nullptr, stream);
  return ffb_field_generic_fwd(f, x, n, n_dev, feats, coeff, nullptr, stream);
}

int ffb_field_query_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats, const float* g_coeff,
                        float* const* h_grads, void* stream) {
  if (ffb::g_deterministic) return ffb_field_generic_bwd(f, x, n, n_dev, g_feats, g_coeff, h_grads, stream);
  if (ffb_field_fast_eligible(f) == 1) return ffb_field_fast_bwd(f, x, n, n_dev, g_feats, g_coeff, h_grads, stream);
  if (ffb_field_lines_eligible(f) == 1) return ffb_field_lines_bwd(f, x, n, n_dev, g_feats, g_coeff, h_grads, stream);
  if (ffb_field_planes_eligible(f) == 1) return ffb_field_planes_bwd(f, x, n, n_dev, g_feats, g_coeff, h_grads, stream);
  return ffb_field_generic_bwd(f, x, n, n_dev, g_feats, g_coeff, h_grads, stream);
}
}
