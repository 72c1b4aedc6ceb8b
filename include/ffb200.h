/* ffb200.h — C ABI of libffb200.so: the B200 (sm_100a) query-and-render hot path of Factor Fields.
 *
 * The reference (autonomousvision/factor-fields) is pure Python/PyTorch and has no FFI of its own;
 * the interface each entry point replaces is therefore a *Python call site* of the reference, cited
 * as file:line (paths relative to the reference checkout).  The Python host in
 * factor-fields_b200/ binds these symbols with ctypes.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch types.  Unless a parameter is named `h_*`, every
 *    pointer is a DEVICE pointer on the current CUDA device; `stream` is a cudaStream_t passed as
 *    void* (NULL = legacy default stream).  Nothing synchronises the stream except the functions
 *    whose name ends in `_host`.
 *  - all functions return 0 on success, a negative FFB_E* code otherwise; ffb_last_error() returns a
 *    thread-local message.  There is NO CPU fallback: without a CUDA device every compute entry
 *    point fails with FFB_ECUDA.
 *  - factor tensors are CHANNELS-LAST: a reference parameter of logical shape [1,C,(D,)H,W]
 *    (FactorFields.py:315,408) is stored as [(D,)H,W,C] — exactly torch's channels_last(_3d) memory
 *    format of the same logical tensor, so state_dict shapes are unchanged.
 *  - "n, n_dev": `n` is the row count (or an upper bound used to size the launch); if `n_dev` is
 *    non-NULL the kernel reads the actual count from device memory (min(*n_dev, n)) so that
 *    data-dependent sizes never need a host round trip.
 */
#ifndef FFB200_H
#define FFB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FFB_ABI_VERSION 1
#define FFB_OK 0
#define FFB_EINVAL (-1)
#define FFB_ECUDA (-2)
#define FFB_ENOMEM (-3)

#define FFB_MAX_OPS 64
#define FFB_MAX_TERMS 40
#define FFB_MAX_FREQ 16
#define FFB_MAX_LAYERS 10

/* basis_mapping, FactorFields.py:11-33 */
enum { FFB_MAP_SAWTOOTH = 0, FFB_MAP_TRIANGLE = 1, FFB_MAP_SINC = 2, FFB_MAP_TRIG = 3, FFB_MAP_X = 4 };

const char* ffb_last_error(void);
int ffb_abi_version(void);
/* Number of kernels launched by this library in this process (bench.py's gpu_launches). */
uint64_t ffb_launch_count(void);
int ffb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Field query: get_coeff / get_basis / get_coding (FactorFields.py:425-533) and their autograd.
 *
 * One ffb_gather_op == one F.grid_sample call site (:433,439,446-450,458,489,495,502-508).
 * A term is the channel-wise product of 1..3 gathers (cp factors, :446-450,:501-508) written to
 * `col .. col+C-1` of the coefficient row or of the concatenated basis row (:513).  The final row is
 * feats[perm[q]] = basis_cat[q] * coeff[perm[q]]  (:514-515 re-ordering for vm, :527 product).
 * ------------------------------------------------------------------------------------------- */
typedef struct ffb_gather_op {
  const float* data;    /* channels-last texels [S2][S1][S0][C] */
  float* grad;          /* same layout; NULL = no gradient for this tensor (fix-grid, frozen) */
  int32_t C;
  int32_t nd;           /* spatial dims sampled: 1, 2 or 3 */
  int32_t size[3];      /* S0 (fastest; grid_sample's x / W), S1 (y / H), S2 (z / D) */
  int32_t src[3];       /* coordinate column feeding each dim; -1 = constant cst[k] */
  float cst[3];
  int32_t space;        /* 0: normalize_coord(x) (:635-637)   1: grid_mapping(x)[..., level] (:481) */
  int32_t level;
  int32_t align_corners;
  int32_t border;       /* 1: padding_mode='border', 0: zeros */
  int32_t nearest;      /* 1: mode='nearest', 0: (bi/tri)linear */
} ffb_gather_op;

typedef struct ffb_term {
  int32_t n_ops;
  int32_t op[3];
  int32_t col;
} ffb_term;

typedef struct ffb_field_desc {
  int32_t xdim;                 /* columns of x (in_dim, +1 in 'images' mode :469-470) */
  int32_t in_dim;               /* dims fed to grid_mapping */
  float aabb_min[3], aabb_max[3];
  int32_t mapping;              /* FFB_MAP_* */
  int32_t n_freq;
  float freq[FFB_MAX_FREQ];     /* self.freq_bands */
  int32_t n_ops;
  ffb_gather_op ops[FFB_MAX_OPS];
  int32_t n_cterms, n_bterms;
  ffb_term cterms[FFB_MAX_TERMS], bterms[FFB_MAX_TERMS];
  int32_t coeff_width;          /* 0 = coeff_type 'none' */
  int32_t basis_width;          /* 0 = basis_type 'none' */
  int32_t basis_is_x;           /* basis_type 'x' (:510-511): basis row = mapped coordinates */
  const int32_t* basis_perm;    /* device [basis_width] or NULL (identity) */
} ffb_field_desc;

typedef struct ffb_field* ffb_field_t;

/* Uploads a copy of the descriptor (h_desc is HOST memory).  Texel pointers must stay valid. */
int ffb_field_create(const ffb_field_desc* h_desc, ffb_field_t* out);
int ffb_field_destroy(ffb_field_t f);
/* get_coding (:523-533): x [n, xdim] -> feats [n, W], coeff [n, W] (either may be NULL). */
int ffb_field_query_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev,
                        float* feats, float* coeff, void* stream);
/* autograd of get_coding: scatter-adds d(sum feats*g_feats + coeff*g_coeff) into the gradient tensors
 * (not zeroed here).  h_grads: HOST array of n_ops device pointers (NULL entry = no gradient for that
 * tensor), or NULL to use ops[i].grad of the descriptor.  g_coeff may be NULL. */
int ffb_field_query_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev,
                        const float* g_feats, const float* g_coeff, float* const* h_grads, void* stream);
/* Training pair: the forward also stores the concatenated basis row in `basis`, an OPAQUE buffer of
 * ceil(n / 32) * 32 * W floats private to this pair (the specialised kernels block it by 32 queries so that a warp
 * writes / reads 128 contiguous bytes per column); the backward then needs no
 * re-gather — it reads the saved `coeff` / `basis` rows, recomputes only the tap indices and weights, and issues the
 * scatter as 16-byte vector reductions.  Fields outside the specialised grid x grid kernels fall back to
 * ffb_field_query_bwd's behaviour (coeff / basis are then ignored). */
int ffb_field_query_fwd_train(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev,
                              float* feats, float* coeff, float* basis, void* stream);
int ffb_field_query_bwd_saved(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev,
                              const float* g_feats, const float* g_coeff, const float* coeff,
                              const float* basis, float* const* h_grads, void* stream);
/* How the training pair lays the opaque `basis` buffer out for a batch of n queries (for tests that inspect it):
 * 0 = blocked by 32 queries (narrow rows), 1 = row-major [n, W] (rows of >= 64 channels: the column-parallel kernels of
 * image.yaml / image_set.yaml), -1 = the field does not use saved rows. */
int ffb_field_saved_basis_layout(ffb_field_t f, int64_t n);
/* grid_mapping (:11-33) on its own: x [n, in_dim] -> out [n, in_dim, F] (trig: [n, in_dim, 2F]). */
int ffb_grid_mapping(const float* x, int64_t n, int32_t in_dim, const float* h_aabb_min,
                     const float* h_aabb_max, const float* h_freq, int32_t n_freq, int32_t mapping,
                     float* out, void* stream);

/* Specialised fast path for grid x grid fields (nerf.yaml / sdf.yaml / image.yaml shapes).
 * Same semantics as ffb_field_query_fwd/bwd; returns FFB_EINVAL if the descriptor is not eligible. */
int ffb_field_fast_eligible(ffb_field_t f);
int ffb_field_fast_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats,
                       float* coeff, void* stream);
int ffb_field_fast_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev,
                       const float* g_feats, const float* g_coeff, float* const* h_grads, void* stream);
int ffb_field_fast_fwd_train(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats,
                             float* coeff, float* basis, void* stream);
int ffb_field_fast_bwd_saved(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev,
                             const float* g_feats, const float* g_coeff, const float* coeff,
                             const float* basis, float* const* h_grads, void* stream);
/* Knobs: launch configurations / kernel generations of the field kernels ("field_fwd_cfg", "field_bwd_cfg", "field_level_parallel",
 * "field_wide", "field_planes_v2", "field_planes_unroll", "field_lines_walk", ...; also FFB_TUNING=key=value,... in the environment) and
 * "field_deterministic" (1: the scatter-add of every field runs serially in query order -> bit-reproducible gradients; for
 * debugging and gradient tests at small sizes). */
int ffb_set_tuning(const char* key, int value);
/* ---------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange: in-place sum all-reduce of the flat fp32 gradient arena as ONE kernel over
 * NVLink / NVSwitch peer memory (allreduce.cu).  The arena is symmetric memory: d_peer_ptrs[p] = rank p's arena as mapped
 * into this process (device array of `world` 64-bit addresses), multicast_ptr = one address backed by all of them (0 when
 * the fabric has no multicast: the kernel then sums peer loads instead of using multimem.ld_reduce / multimem.st).
 * d_signal_pads[p]: rank p's signal pad (>= world*4 bytes, zeroed once); d_epoch: 4 zeroed uint32 on this device; blocks <= SM count.
 * Stream-ordered and capturable in a CUDA graph; every rank launches it with the same arguments in the same order.
 * ------------------------------------------------------------------------------------------- */
int ffb_allreduce_symm(float* local, const uint64_t* d_peer_ptrs, uint64_t multicast_ptr,
                       const uint64_t* d_signal_pads, uint32_t* d_epoch, int32_t rank, int32_t world,
                       int64_t n_floats, int32_t blocks, void* stream);

/* Measurement probe (bench.py): issues blocks*256*iters `red.global.add.v4.f32` into buf[0..n_floats) — the instruction
 * the scatter kernels are made of — so that the backward pass can be reported against a MEASURED L2-reduction
 * throughput instead of the HBM roofline.  pattern 0: random 16-byte slots; 1: 512 contiguous bytes per warp. */
int ffb_probe_red(float* buf, int64_t n_floats, int32_t blocks, int32_t iters, int32_t pattern,
                  int64_t* n_ops_out, void* stream);
/* Vector-coefficient x line-product (CP) fields — coeff_type 'vec', basis_type 'cp' (FactorFields.py:437-441,497-509) — with
 * the factors resident in shared memory one level at a time (TMA bulk copies; shared-memory-privatised gradient
 * accumulation, one vector reduction per touched 16 bytes per CTA).  The ffb_field_query_* entry points pick them when
 * ffb_field_lines_eligible(f) == 1; ffb_set_field_lines(0) forces the generic kernels (tests compare both). */
int ffb_set_field_lines(int enabled);
int ffb_field_lines_eligible(ffb_field_t f);
int ffb_field_lines_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff,
                        float* basis, void* stream);
int ffb_field_lines_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats,
                        const float* g_coeff, float* const* h_grads, void* stream);
/* Vector-matrix (vm) fields — coeff_type 'vm' (three coefficient lines, FactorFields.py:443-450) x basis_type 'vm' (three planes
 * per level + the column re-ordering of :514-515) — as specialised gather / vector-reduction kernels (field_planes.cu); picked by
 * the ffb_field_query_* entry points when ffb_field_planes_eligible(f) == 1 (2- or 4-channel planes, single scene). */
int ffb_set_field_planes(int enabled);
int ffb_field_planes_eligible(ffb_field_t f);
int ffb_field_planes_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff,
                         float* basis, void* stream);
int ffb_field_planes_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats,
                         const float* g_coeff, float* const* h_grads, void* stream);
/* coeff: the [n, W] coefficient rows ffb_field_planes_fwd returned (row-major); with basis == NULL — what the training path
 * passes — the kernel streams them in instead of re-gathering the three lines and re-gathers only the (L2-resident) planes.
 * coeff == NULL: everything is re-gathered.  basis != NULL (rows of a forward that was asked for them): first-generation kernel. */
int ffb_field_planes_bwd_saved(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats,
                               const float* g_coeff, const float* coeff, const float* basis, float* const* h_grads,
                               void* stream);
/* The descriptor-driven generic kernels, callable directly (parity tests compare both paths).
 * basis_out: optional [n, W] copy of the (re-ordered) basis row. */
int ffb_field_generic_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats,
                          float* coeff, float* basis_out, void* stream);
int ffb_field_generic_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev,
                          const float* g_feats, const float* g_coeff, float* const* h_grads, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense layers: nn.Linear call sites of MLPMixer (:153-156) and MLPRender_Fea (:197-200).
 * Row-major: x [n,K], W [M,K] (torch layout), y [n,M].
 * ------------------------------------------------------------------------------------------- */
/* y = act(x W^T + b); act: 0 none, 1 relu, 2 sigmoid.  b may be NULL. */
int ffb_linear_fwd(const float* x, const float* W, const float* b, float* y, int64_t n,
                   const int32_t* n_dev, int32_t K, int32_t M, int32_t act, void* stream);
/* Same, choosing the tensor-core operand split: split_terms 3 = fp32-class accuracy (default), 2 = ~5e-6 relative and
 * half the MMAs (used for the appearance MLP, whose output does not feed exp()).  Ignored by the exact SIMT path. */
int ffb_linear_fwd_ex(const float* x, const float* W, const float* b, float* y, int64_t n,
                      const int32_t* n_dev, int32_t K, int32_t M, int32_t act, int32_t split_terms,
                      void* stream);
/* gx = (gy .* act'(y)) W, where y is the saved forward OUTPUT of the layer and act the activation that
 * produced it (mask applied on the fly; gy is not modified).  act == 0: y may be NULL. */
int ffb_linear_bwd_input(float* gy, const float* y, const float* W, float* gx, int64_t n,
                         const int32_t* n_dev, int32_t K, int32_t M, int32_t act, void* stream);
/* gW += (gy .* act'(y))^T x ; gb += sum_rows (gy .* act'(y)).  gb may be NULL.  Not zeroed here. */
int ffb_linear_bwd_weight_act(const float* gy, const float* y, int32_t act, const float* x, float* gW,
                              float* gb, int64_t n, const int32_t* n_dev, int32_t K, int32_t M,
                              void* stream);
/* Skinny layers (M <= 8 outputs): input gradient, weight gradient and bias gradient in one exact-fp32 pass.
 * gx [n,K] = (gy .* act'(y)) W (may be NULL); gW / gb accumulate (+=). */
int ffb_linear_bwd_skinny(const float* gy, const float* y, int32_t act, const float* x, const float* W,
                          float* gx, float* gW, float* gb, int64_t n, const int32_t* n_dev, int32_t K,
                          int32_t M, void* stream);
int ffb_linear_bwd_weight(const float* gy, const float* x, float* gW, float* gb, int64_t n,
                          const int32_t* n_dev, int32_t K, int32_t M, void* stream);

/* Tensor-core (tcgen05.mma, bf16 hi+lo split x3, fp32 accumulation in TMEM) variants of the three layer kernels.
 * ffb_linear_fwd / _bwd_input / _bwd_weight_act dispatch to them when the shape is eligible and n >= 1024;
 * ffb_set_tensor_cores(0) forces the exact-fp32 SIMT kernels. */
int ffb_set_tensor_cores(int enabled);
int ffb_tensor_cores_enabled(void);
int ffb_linear_tc_eligible(int32_t K, int32_t N);
int ffb_linear_tc_wgrad_eligible(int32_t K, int32_t M);
int ffb_linear_tc_fwd(const float* x, const float* W, const float* b, float* y, int64_t n,
                      const int32_t* n_dev, int32_t K, int32_t M, int32_t act, void* stream);
int ffb_linear_tc_fwd_ex(const float* x, const float* W, const float* b, float* y, int64_t n,
                         const int32_t* n_dev, int32_t K, int32_t M, int32_t act, int32_t split_terms,
                         void* stream);
int ffb_linear_tc_bwd_input(const float* gy, const float* y, const float* W, float* gx, int64_t n,
                            const int32_t* n_dev, int32_t K, int32_t M, int32_t act, void* stream);
int ffb_linear_tc_bwd_weight(const float* gy, const float* y, int32_t act, const float* x, float* gW,
                             float* gb, int64_t n, const int32_t* n_dev, int32_t K, int32_t M,
                             void* stream);

/* Whole 2-layer MLPMixer (:144-159 with pe = 0, no dropout) in one launch: y = relu(x W1^T + b1) W2^T.
 * x [n,K0], W1 [H,K0], b1 [H] (may be NULL), W2 [N,H], y [n,N].  The hidden activation stays on the SM (TMEM / shared
 * memory); the backward pass recomputes its values and takes the ReLU decisions from relu_mask [n, H/16] uint16 (bit i
 * of word c = hidden unit 16c+i was positive), written by the forward pass (NULL: not written / re-decided).
 * Eligible shapes: H == 64 and the operand tiles fit in shared memory (ffb_mlp2_eligible); everything else goes
 * through the per-layer entry points above.
 * ffb_mlp2_bwd: gx [n,K0] (may be NULL), gW1 [H,K0], gb1 [H], gW2 [N,H] are ACCUMULATED (+=), not zeroed. */
int ffb_set_fused_mlp(int enabled);
/* Shapes with K0 <= 31, H == 64, N <= 32 (linear_mat of nerf.yaml) run as pipelined warp-specialised kernels (mlp_pipe.cu:
 * producer warps -> operand-tile ring -> MMA issuer -> epilogue warps); ffb_set_mlp_pipelined(0) keeps the one-tile-at-a-time kernels. */
int ffb_set_mlp_pipelined(int enabled);
int ffb_mlp2_pipelined_eligible(int32_t K0, int32_t H, int32_t N);
int ffb_mlp2_eligible(int32_t K0, int32_t H, int32_t N);
int ffb_mlp2_fwd(const float* x, const float* W1, const float* b1, const float* W2, float* y,
                 uint16_t* relu_mask, int64_t n, const int32_t* n_dev, int32_t K0, int32_t H, int32_t N,
                 void* stream);
int ffb_mlp2_bwd(const float* x, const float* gy, const float* W1, const float* b1, const float* W2,
                 const uint16_t* relu_mask, float* gx, float* gW1, float* gb1, float* gW2, int64_t n,
                 const int32_t* n_dev, int32_t K0, int32_t H, int32_t N, void* stream);
/* ffb_mlp2_bwd for the SPARSE upstream gradient of the render path (shapes with ffb_mlp2_pipelined_eligible == 1): every
 * sample has a density gradient, only the shaded ones (weight > threshold, :879-881) have feature gradients.
 * gy[i, 0] = g0[i];  gy[i, 1:] = g_rows[row_slot[i], 1:] where row_slot[i] >= 0, zero elsewhere (g_rows [*, N]; its column
 * 0 is ignored).  The dense [n, N] gradient (126 MB written and read per nerf.yaml step) is never materialised. */
int ffb_mlp2p_bwd_sparse(const float* x, const float* g0, const int32_t* row_slot, const float* g_rows,
                         const float* W1, const float* b1, const float* W2, const uint16_t* relu_mask, float* gx,
                         float* gW1, float* gb1, float* gW2, int64_t n, const int32_t* n_dev, int32_t K0,
                         int32_t H, int32_t N, void* stream);

/* ---------------------------------------------------------------------------------------------
 * get_coding + linear_mat as ONE kernel per direction (field_mlp.cu): FactorFields.py:425-533 followed by
 * MLPMixer.forward :144-159 (2 layers, hidden 64, no PE / dropout), for grid x grid fields with linear taps.
 * The feature row never reaches HBM: gather warps write it into shared-memory tensor-core operand tiles.
 *   y [n,N]; relu_bits [n,4] uint16 (as ffb_mlp2_fwd); coeff_blk / basis_blk: the coefficient / basis rows BLOCKED by 32
 *   queries (element (i,c) at (i/32)*32*W + c*32 + i%32; buffers of ceil(n/32)*32 rows) for ffb_field_mlp_bwd; optional
 *   row-major copies feats [n,W] / coeff [n,W] (NULL: not written).
 * ffb_field_mlp_bwd: g_y [n,N] -> factor gradients (h_grads as ffb_field_query_bwd) and gW1 / gb1 / gW2, all ACCUMULATED. */
int ffb_set_field_mlp(int enabled);
int ffb_field_mlp_eligible(ffb_field_t f, int32_t K0, int32_t H, int32_t N);
int ffb_field_mlp_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* W1,
                      const float* b1, const float* W2, float* y, uint16_t* relu_bits, float* coeff_blk,
                      float* basis_blk, float* feats, float* coeff, int32_t K0, int32_t H, int32_t N,
                      void* stream);

/* positional_encoding (:74-79) appended to the input: out [n, D + 2*D*pe] = [x, sin, cos]. */
int ffb_pe_concat_fwd(const float* x, float* out, int64_t n, const int32_t* n_dev, int32_t D,
                      int32_t pe, void* stream);
/* gx [n,D] = g[:, :D] + dPE */
int ffb_pe_concat_bwd(const float* x, const float* g, float* gx, int64_t n, const int32_t* n_dev,
                      int32_t D, int32_t pe, void* stream);
/* MLPRender_Fea input assembly (:190-196) with gather: for row j, i = app_idx ? app_idx[j] : j,
 * features = feat[i, 1:1+C] (row stride ld_feat), viewdir = rays[ray_id ? ray_id[i] : i, 3:6]
 * (row stride 6) -> out [n, 3 + C + 6*viewpe + 2*feape*C]. */
int ffb_render_input_fwd(const float* feat, int32_t ld_feat, const float* rays, const int32_t* ray_id,
                         const int32_t* app_idx, float* out, int64_t n, const int32_t* n_dev,
                         int32_t C, int32_t viewpe, int32_t feape, void* stream);
/* scatter of the gradient back: g_feat[i, 1:1+C] += d(features) (rows are unique -> plain add).
 * ld_gin: row stride of g_in in floats (0 = the input width). */
int ffb_render_input_bwd(const float* feat, int32_t ld_feat, const int32_t* app_idx, const float* g_in,
                         int32_t ld_gin, float* g_feat, int64_t n, const int32_t* n_dev, int32_t C,
                         int32_t viewpe, int32_t feape, void* stream);
/* Same, WRITING row j of the compact g_app [n, ld_app] (columns 1..C; column 0 untouched) instead of adding into row
 * app_idx[j] of a dense gradient: the rows ffb_mlp2p_bwd_sparse consumes. */
int ffb_render_input_bwd_compact(const float* feat, int32_t ld_feat, const int32_t* app_idx, const float* g_in,
                                 int32_t ld_gin, float* g_app, int32_t ld_app, int64_t n, const int32_t* n_dev,
                                 int32_t C, int32_t viewpe, int32_t feape, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused appearance MLP (MLPRender_Fea.forward, FactorFields.py:188-203, on the shaded samples of
 * forward :879-885): input assembly + 3 layers + sigmoid in one tcgen05 kernel (mlp_rgb.cu).
 * Shape: hidden width 128, 3 layers (bias, bias, none), 3 outputs.
 * ------------------------------------------------------------------------------------------- */
/* Bytes of the packed-weight workspace, or 0 when the shape is not eligible (callers then use the per-layer ops). */
int64_t ffb_rgbmlp_workspace_bytes(int32_t Cf, int32_t hidden, int32_t view_pe, int32_t fea_pe);
/* Split W1 [128, K0] (+ b1 as an extra column), W2 [128,128], W3 [3,128] into bf16 operand slices (once per step). */
int ffb_rgbmlp_pack(const float* W1, const float* b1, const float* W2, const float* W3, void* workspace,
                    int32_t Cf, int32_t view_pe, int32_t fea_pe, void* stream);
/* rgb [n,3] = MLPRender_Fea(viewdirs = rays[ray_id[i], 3:6], features = feat[i, 1:1+Cf]) for i = app_idx[j] (NULL:
 * identity).  relu_bits [n,16] uint16 (optional): ReLU decisions of both hidden layers for the backward pass.
 * x_out [n,K0] / h1_out / h2_out [n,128] (optional): fp32 copies of the MLP input and hidden activations.
 * stream_x / stream_h1 / stream_h2 (all or none): bf16 operand streams of the same activations for ffb_rgbmlp_bwd,
 * ffb_rgbmlp_stream_bytes(.., n, which) bytes each (which: 0 = x, 1 = h1 / h2). */
int ffb_rgbmlp_fwd(const float* feat, int32_t ld_feat, const float* rays, const int32_t* ray_id,
                   const int32_t* app_idx, const void* workspace, const float* b2, float* rgb,
                   uint16_t* relu_bits, float* x_out, float* h1_out, float* h2_out, void* stream_x,
                   void* stream_h1, void* stream_h2, int64_t n, const int32_t* n_dev, int32_t Cf,
                   int32_t view_pe, int32_t fea_pe, void* stream);
int64_t ffb_rgbmlp_stream_bytes(int32_t Cf, int32_t view_pe, int32_t fea_pe, int64_t n, int32_t which);
/* Backward of the same MLP (the autograd of FactorFields.py:188-203): g_rgb [n,3] = dL/d rgb, rgb = the forward
 * output.  Writes g_x [n, ld_gx] (ld_gx >= K0, a multiple of 4: 16-byte row alignment for vector stores; gradient
 * w.r.t. the assembled input; ffb_render_input_bwd folds it onto the features)
 * and ACCUMULATES gW1 [128,K0], gb1, gW2 [128,128], gb2, gW3 [3,128] (any may be NULL).  workspace: the one
 * ffb_rgbmlp_pack filled for this step's weights; W3: the fp32 colour-head weight. */
int ffb_rgbmlp_bwd(const float* g_rgb, const float* rgb, const uint16_t* relu_bits, const void* stream_x,
                   const void* stream_h1, const void* stream_h2, const void* workspace, const float* W3,
                   float* g_x, int32_t ld_gx, float* gW1, float* gb1, float* gW2, float* gb2, float* gW3,
                   int64_t n, const int32_t* n_dev, int32_t Cf, int32_t view_pe, int32_t fea_pe, void* stream);
int ffb_set_fused_rgbmlp(int enabled);

/* ---------------------------------------------------------------------------------------------
 * Ray sampling + alpha-mask stream compaction: sample_point (:586-602), AlphaGridMask.sample_alpha
 * (:103-110) and the boolean-mask indexing of forward (:864-867,874).
 * ------------------------------------------------------------------------------------------- */
enum {
  FFB_SAMPLE_BOUNDED = 0, /* sample_point (:586-602): march from the box entry with step_size, keep in-box samples */
  FFB_SAMPLE_NDC = 1,     /* sample_point_ndc (:575-584): interpx from z_table, keep in-box samples, dists * |d| (:854-856) */
  FFB_SAMPLE_UNBOUND = 2  /* sample_point_unbound (:604-633): interpx from z_table, inf-norm contraction of the points
                             outside the unit cube, every sample kept (:863), last dist repeats the previous (:850) */
};

typedef struct ffb_sampler_desc {
  float aabb_min[3], aabb_max[3];
  float step_size;              /* self.stepSize (:697), an fp32 value */
  int32_t n_samples;
  const uint8_t* alpha_volume;  /* [D][H][W] 0/1, or NULL (self.alphaMask is None) */
  int32_t alpha_size[3];        /* W, H, D */
  float alpha_aabb_min[3];
  float alpha_inv_size[3];      /* AlphaGridMask.invgridSize (:98) */
  float alpha_thres;            /* > thres keeps the sample: 0.5 in forward (:866), 0 in filtering (:833) */
  int32_t mode;                 /* FFB_SAMPLE_* */
  float bg_len;                 /* self.bg_len (:250), FFB_SAMPLE_UNBOUND only */
  const float* z_table;         /* device [n_samples]: the interpx row shared by all rays (:578-580, :611-623); NULL in mode 0 */
  int32_t alpha_outside;        /* 1: look the alpha volume up for out-of-box samples too (filtering_rays :832-833) */
} ffb_sampler_desc;

/* Pass 1: counts[r] = number of valid samples of ray r; tmin[r] = entry distance.
 * jitter [R] (is_train, :593-595) or NULL. rays [R,6]. */
int ffb_sample_count(const ffb_sampler_desc* h_desc, const float* rays, const float* jitter,
                     int64_t R, int32_t* counts, float* tmin, void* stream);
/* Exclusive prefix sum: offsets [R+1]; offsets[R] = total. */
int ffb_exclusive_scan_i32(const int32_t* counts, int32_t* offsets, int64_t R, void* stream);
/* Pass 2: compacted, row-major (ray, sample) order — identical to torch boolean-mask indexing.
 * xyz [Nv,3], ray_id [Nv], sample_id [Nv], z [Nv] (interpx), dist [Nv] (:861; last sample 0). */
int ffb_sample_fill(const ffb_sampler_desc* h_desc, const float* rays, const float* jitter,
                    const float* tmin, const int32_t* offsets, int64_t R, int64_t cap, float* xyz,
                    int32_t* ray_id, int32_t* sample_id, float* z, float* dist, void* stream);
/* Dense variant behind the sample_point* methods and the parity tests: mask [R,S] (uint8; in-box / inner mask, or
 * ray_valid when an alpha volume is attached), z [R,S] or NULL, pts [R,S,3] or NULL (contracted in mode 2). */
int ffb_sample_dense(const ffb_sampler_desc* h_desc, const float* rays, const float* jitter, int64_t R,
                     uint8_t* mask, float* z, float* pts, void* stream);
/* sample_alpha (:103-110) for arbitrary points: out [n] float. */
int ffb_alpha_sample(const ffb_sampler_desc* h_desc, const float* xyz, int64_t n, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Volume-rendering composite: basis2density (:639-643), raw2alpha (:82-88), weight threshold
 * (:879-881), accumulation (:887-896) — over the COMPACTED sample list.
 * ------------------------------------------------------------------------------------------- */
typedef struct ffb_composite_desc {
  float density_shift;          /* cfg.renderer.density_shift */
  int32_t softplus;             /* 1: softplus, 0: relu (fea2denseAct) */
  float distance_scale;
  float weight_thres;           /* rayMarch_weight_thres */
  int32_t white_bg;             /* 1: rgb_map += 1 - acc (FactorFields.py:890-891) */
  const int32_t* white_bg_dev;  /* optional DEVICE flag that overrides white_bg when non-null: the per-step coin flip of
                                   non-white-background scenes (:890 `is_train and torch.rand((1,)) < 0.5`) stays a
                                   run-time value inside a captured CUDA graph */
} ffb_composite_desc;

/* Phase A, per ray: sigma, alpha, T, weight for each valid sample; app_counts[r] = #(weight>thres).
 * feat0 = density feature with row stride ld_feat (feat[:,0] of linear_mat's output). */
int ffb_composite_weights(const ffb_composite_desc* h_desc, const float* feat0, int32_t ld_feat,
                          const float* dist, const int32_t* offsets, int64_t R, float* sigma,
                          float* trans, float* weight, int32_t* app_counts, void* stream);
/* app_idx [Na]: indices (into the valid list) of shaded samples, in order. */
int ffb_composite_app_fill(const float* weight, float weight_thres, const int32_t* offsets,
                           const int32_t* app_offsets, int64_t R, int32_t* app_idx, void* stream);
/* Same, also writing the inverse map app_slot [Nv] (may be NULL): slot of sample i in app_idx, or -1 when it is not shaded. */
int ffb_composite_app_fill_ex(const float* weight, float weight_thres, const int32_t* offsets,
                              const int32_t* app_offsets, int64_t R, int32_t* app_idx, int32_t* app_slot,
                              void* stream);
/* Phase B: rgb_map [R,3] (clamped), acc [R], depth [R]; rgb [Na,3] in app order. pre_clamp [R,3]. */
int ffb_composite_accum(const ffb_composite_desc* h_desc, const float* weight, const float* z,
                        const float* rgb, const int32_t* offsets, const int32_t* app_offsets, int64_t R,
                        float* rgb_map, float* pre_clamp, float* acc, float* depth, void* stream);
/* Backward: g_rgb_map [R,3] -> g_rgb [Na,3], g_feat0 (row stride ld_g; d loss / d density feature).
 * zero_rest != 0: also zero columns 1..ld_g-1 of every valid sample's row (the caller then needs no memset of the
 * [Nv, ld_g] gradient buffer before the appearance MLP adds its feature gradients). */
int ffb_composite_bwd(const ffb_composite_desc* h_desc, const float* g_rgb_map, const float* pre_clamp,
                      const float* feat0, int32_t ld_feat, const float* dist, const float* sigma,
                      const float* trans, const float* weight, const float* rgb, const int32_t* offsets,
                      const int32_t* app_offsets, int64_t R, float* g_rgb, float* g_feat0, int32_t ld_g,
                      int32_t zero_rest, void* stream);
/* compute_alpha (:710-727) tail: alpha[i] = 1 - exp(-basis2density(feat0[i]) * length). */
int ffb_density_alpha(const ffb_composite_desc* h_desc, const float* feat0, int32_t ld_feat, float length,
                      int64_t n, const int32_t* n_dev, float* alpha, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Train-step glue: loss (train_per_scene.py:158) and Adam (:124-132,160-162,170-171).
 * ------------------------------------------------------------------------------------------- */
/* loss[0] += mean((pred-target)^2) over n elements (loss must be zeroed by the caller);
 * g_pred = 2 (pred-target)/n * g_scale * (g_scale_dev ? *g_scale_dev : 1): g_scale_dev is an optional DEVICE
 * scalar, e.g. the decaying loss scale of scripts/2D_regression.ipynb cell 4 / sdf_regression.ipynb cell 2. */
int ffb_mse_fwd_bwd(const float* pred, const float* target, int64_t n, float g_scale,
                    const float* g_scale_dev, float* loss, float* g_pred, void* stream);
/* *d_value *= factor in double (the notebooks' `loss_scale *= lr_factor`); *d_out_f32 = (float)*d_value if given. */
int ffb_scalar_decay(double* d_value, double factor, float* d_out_f32, void* stream);
/* torch.optim.Adam semantics (no amsgrad, no weight decay), fp32; step is 1-based. grad_scale
 * multiplies g first (1/world after a sum all-reduce). */
int ffb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, int32_t step, float grad_scale, void* stream);

/* Multi-tensor Adam, one launch for all parameter tensors, CUDA-graph friendly: every step-dependent scalar lives in
 * DEVICE memory.  d_table [T][6] int64 = {p, g, m, v (device addresses), n, group}; block b updates `chunk` elements of
 * tensor d_chunk_tensor[b] from d_chunk_start[b]; d_hyper [n_groups][2] = {lr/(1-beta1^t), 1/sqrt(1-beta2^t)}. */
/* Device-side scalar bookkeeping for ffb_adam_multi: *d_step += 1; d_hyper[g] = {lr_g/(1-beta1^t), 1/sqrt(1-beta2^t)};
 * d_lr[g] *= lr_decay (train_per_scene.py:170-171), all in double as the Python reference. */
int ffb_adam_hyper_advance(double* d_lr, int64_t* d_step, float* d_hyper, int32_t n_groups, double beta1,
                           double beta2, double lr_decay, void* stream);
int ffb_adam_multi(const int64_t* d_table, const int32_t* d_chunk_tensor, const int64_t* d_chunk_start,
                   int32_t n_chunks, int32_t chunk, const float* d_hyper, float beta1, float beta2, float eps,
                   float grad_scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer entry point (the end-to-end boundary: host rays in, host pixels out; includes the
 * H2D/D2H copies and a stream synchronise).  Mirrors renderer.py:8-27 + FactorFields.py:586-602 for
 * the sampling stage; used by tests to exercise the C ABI without torch.
 * ------------------------------------------------------------------------------------------- */
int ffb_sample_dense_host(const ffb_sampler_desc* h_desc, const float* h_rays, const float* h_jitter,
                          int64_t R, uint8_t* h_mask, float* h_z);

#ifdef __cplusplus
}
#endif
#endif /* FFB200_H */
