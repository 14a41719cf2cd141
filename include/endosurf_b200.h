/* endosurf_b200 -- C ABI of the B200-native EndoSurf volume-rendering hot path.
 *
 * The reference (Ruyi-Zha/endosurf) has no FFI: its "operator interface" for this path is the Python class
 * EndoSurfRenderer (src/renderer/endosurf.py:14-521) and the network queries of EndoSurfNet (:524-689).  Each entry
 * point below states which reference method it replaces.  The Python mirror of the class
 * (endosurf_b200/renderer.py) binds these symbols with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 unless stated otherwise; the caller owns all buffers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls only enqueue work;
 *   - return value: 0 ok, >0 a cudaError_t, <0 an argument error (ES_E_*); es_last_error() gives text;
 *   - no global state: everything lives in the es_ctx (packed weights, layer programs, scratch workspace).
 */
#ifndef ENDOSURF_B200_H
#define ENDOSURF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ES_E_BADARG (-1)
#define ES_E_UNSUPPORTED (-2) /* network shape outside what the sm_100a kernels are built for */
#define ES_E_NOWEIGHTS (-3)
#define ES_E_DEVICE (-4) /* device-side watchdog / barrier error word set; see es_sync_check */

typedef struct es_ctx es_ctx;

/* Mirrors cfg["net"] of the reference (configs/endosurf/baseline/base_pull.yml:40-82). */
typedef struct es_net_config {
  int32_t use_deform;       /* net.use_deform */
  int32_t n_layers;         /* 9  (all three networks) */
  int32_t skip_layer;       /* 4  (skips=[4]); -1 = none */
  int32_t hidden_dim;       /* 256 (only value supported) */
  int32_t multires_deform_pos, multires_deform_time; /* 6, 6 */
  int32_t multires_sdf_pos;                          /* 6 */
  int32_t multires_color_pos, multires_color_dir;    /* 10, 4 */
  int32_t precision_terms;  /* 3 = fp16x3 hi/lo split (fp32 parity, default); 1 = single fp16 pass */
} es_net_config;

enum { ES_NET_DEFORM = 0, ES_NET_SDF = 1, ES_NET_COLOR = 2 };

int es_create(es_ctx** out, const es_net_config* cfg);
void es_destroy(es_ctx* ctx);
const char* es_last_error(const es_ctx* ctx);
/* Synchronises `stream` and returns ES_E_DEVICE if any kernel recorded a device-side error. */
int es_sync_check(es_ctx* ctx, void* stream);
/* Non-blocking variant for the production path: enqueues a copy of the error word to pinned host memory and reports
 * what the PREVIOUS poll delivered, so a tripped watchdog raises at most one call late instead of training on. */
int es_poll_error(es_ctx* ctx, void* stream);
int es_num_sms(const es_ctx* ctx);
/* Synchronises the device and frees the grow-only scratch workspace (it is re-grown on the next call that needs it). */
int es_release_workspace(es_ctx* ctx);

/* Upload one network's EFFECTIVE weights W_l = g_l * v_l / ||v_l|| (reference utils.py:57-58, folded by the caller)
 * and biases: w[l] -> [out_l, in_l] row-major, b[l] -> [out_l], l = 0..n_layers-1, in the reference's own layout
 * (DeformNetwork/SDFNetwork/ColorNetwork.net[l], endosurf.py:713,762,817).  Packs them into fp16 hi/lo UMMA units. */
int es_load_network(es_ctx* ctx, int net, const float* const* w, const float* const* b, void* stream);

/* EndoSurfNet.get_sdf_from_observed_space (endosurf.py:570-579).
 * x [n,3]; t: element (i / t_div) * t_stride is the time of point i; sdf_out [n]. */
int es_sdf_query(es_ctx* ctx, const float* x, const float* t, int64_t t_div, int64_t t_stride, int64_t n,
                 float* sdf_out, void* stream);

/* extract_fields (utils.py:139-157): the SDF on a resolution^3 grid over [bound_min, bound_max] at one time t (device
 * scalar), sdf_out [res][res][res] in meshgrid(indexing="ij") order.  bound_* are HOST float[3].  The grid points are
 * generated on the device slab by slab; nothing goes through the host. */
int es_sdf_grid(es_ctx* ctx, const float* bound_min3, const float* bound_max3, int32_t resolution, const float* t,
                float* sdf_out, void* stream);

/* EndoSurfNet.forward (endosurf.py:660-689) plus the three gradient queries it and render_core use
 * (get_sdf_grad_from_canonical_space :603-619, get_deform_grad_from_observed_space :621-658,
 *  get_sdf_grad_from_observed_space :581-601 == J^T g_c).
 * dirs: view direction of point i is row (i / dir_div) with row stride dir_stride floats.
 * Outputs (any may be NULL except those needed downstream): x_c [n,3], jac [n,9] (row-major d x_c_i / d x_j),
 * sdf [n], g_c [n,3], feat [n,256], rgb [n,3]. */
int es_point_forward(es_ctx* ctx, const float* x, const float* t, int64_t t_div, int64_t t_stride, const float* dirs,
                     int64_t dir_div, int64_t dir_stride, int64_t n, float* x_c, float* jac, float* sdf, float* g_c,
                     float* feat, float* rgb, void* stream);

/* Same as es_load_network but from the reference's parameters themselves (old-API weight norm, utils.py:57-58):
 * v[l] = weight_v [out_l, in_l], g[l] = weight_g [out_l, 1], b[l] = bias.  The fold W = g v / |v|_row runs on the
 * device into buffers owned by the context. */
int es_load_network_wn(es_ctx* ctx, int net, const float* const* v, const float* const* g, const float* const* b,
                       void* stream);

/* ---- differentiable (training) path -------------------------------------------------------------------------------
 * The reference differentiates render_rays with torch.autograd: loss.backward() (trainer_endosurf.py:94-104) runs
 * through render_core (endosurf.py:168-203), EndoSurfNet.forward and the create_graph=True gradient queries
 * (endosurf.py:594-658).  Here the forward keeps the input rows of every MMA layer (primal + 3 tangent rows per
 * point) as fp16 plane records in the caller's `stash` buffer, and the backward is: fused compositing backward ->
 * three reverse tcgen05 chains on transposed weights (activation gates from the stash, adjoints of every
 * pre-activation written as plane records) -> input-adjoint launches (adjoints of the positional encodings, g_c,
 * J d and the geometry feature) -> split-K tcgen05 weight-gradient kernel over the plane records -> weight-norm
 * backward.  No library GEMMs, no PyTorch ops.
 *
 * A plane record is [tile][chunks per tile][16 KiB]; a chunk is 64 columns of a 128-row tile in the kernels' ring
 * slot layout [k-group of 8 columns][128 rows][8 fp16].  Geometry tiles hold 32 points x 4 streams (row 32Q + 8s + p),
 * colour tiles 128 points.  es_train_stash_bytes gives the size of `stash` for n points in the current plane mode.
 * Plane mode 0 (default): fp16 hi planes only (1-term weight gradients, hi-only activation gates) plus the lo halves
 * the input-adjoint launches need; 1: every lo half as well (3-term weight gradients, exact gates). */
int es_set_plane_mode(es_ctx* ctx, int32_t full_planes);
int es_train_stash_bytes(const es_ctx* ctx, int64_t n, int64_t* out);

/* Parameters and gradient outputs of a backward call; index [net][layer], net = ES_NET_*, tables of n_layers
 * pointers (the deform tables may be NULL without a deformation network).  v / g: weight_v / weight_g as given to
 * es_load_network_wn; grad_*: written (not accumulated), same shapes as the parameters. */
typedef struct es_train_params {
  const float* const* v[3];
  const float* const* g[3];
  float* const* grad_v[3];
  float* const* grad_g[3];
  float* const* grad_b[3];
} es_train_params;

/* EndoSurfNet.forward + gradient queries on explicit points with the stash kept (errorondepth /
 * surface_neighbour_error during training, endosurf.py:289-342).  Outputs as es_point_forward (all required). */
int es_point_train_forward(es_ctx* ctx, const float* x, const float* t, int64_t t_div, int64_t t_stride,
                           const float* dirs, int64_t dir_div, int64_t dir_stride, int64_t n, float* x_c, float* jac,
                           float* sdf, float* g_c, float* rgb, uint8_t* stash, void* stream);
/* ... and its reverse: adjoints of sdf [n], g_c [n,3], jac [n,9], rgb [n,3] (any may be NULL = zero) -> parameter
 * gradients.  x_c, jac, g_c, rgb: the forward's outputs. */
int es_point_train_backward(es_ctx* ctx, int64_t n, const float* dirs, int64_t dir_div, int64_t dir_stride,
                            const float* x_c, const float* jac, const float* g_c, const float* rgb,
                            const uint8_t* stash, const float* sdf_bar, const float* gc_bar, const float* jac_bar,
                            const float* rgb_bar, const es_train_params* prm, void* stream);

struct es_render_out;
/* Adjoints of the render outputs (any may be NULL); gradient_o_error: device scalar. */
typedef struct es_render_grads {
  const float* color_map;        /* [R,3] */
  const float* depth_map;        /* [R] */
  const float* gradients_o;      /* [R,M,3] */
  const float* gradient_o_error; /* [1] */
  const float* weights;          /* [R,M] */
  const float* cdf;              /* [R,M] */
  const float* sdf;              /* [R,M] */
  const float* sampled_color;    /* [R,M,3] */
} es_render_grads;

/* render_core (endosurf.py:134-213) on given sample positions z_vals [R,M] with everything the backward needs kept:
 * x_c [P,3], jac [P,9] (NULL without deform), sdf [P], g_c [P,3], rgb [P,3] (P = R M) are outputs the caller keeps,
 * eik_den [1] receives sum(relax) + 1e-6.  `out` as in es_render_rays (8 required keys). */
int es_render_train_forward(es_ctx* ctx, const float* rays, int64_t n_rays, const float* z_vals, int32_t m,
                            int32_t n_samples, float cos_anneal_ratio, const float* variance, float* x_c, float* jac,
                            float* sdf, float* g_c, float* rgb, uint8_t* stash, const struct es_render_out* out,
                            float* eik_den, void* stream);
/* Its reverse.  variance_grad [1] receives d loss / d variance through inv_s (the caller adds s_val's own path). */
int es_render_train_backward(es_ctx* ctx, const float* rays, int64_t n_rays, const float* z_vals, int32_t m,
                             int32_t n_samples, float cos_anneal_ratio, const float* variance, const float* x_c,
                             const float* jac, const float* sdf, const float* g_c, const float* rgb,
                             const uint8_t* stash, const float* eik_den, const es_render_grads* bar,
                             const es_train_params* prm, float* variance_grad, void* stream);

/* EndoSurfRenderer.up_sample (endosurf.py:221-266) incl. sample_pdf(det=True) (utils.py:160-191).
 * rays [R,9]; z, sdf [R,n]; u_vals [n_imp] = linspace(.5/n_imp, 1-.5/n_imp); new_z [R,n_imp]. */
int es_up_sample(es_ctx* ctx, const float* rays, int64_t n_rays, const float* z, const float* sdf, int32_t n,
                 int32_t n_imp, const float* u_vals, float inv_s, float* new_z, void* stream);

typedef struct es_render_params {
  int32_t n_samples;        /* render.n_samples */
  int32_t n_importance;     /* render.n_importance (0 or not up-sampling: no hierarchical sampling) */
  int32_t up_sample_steps;  /* render.up_sample_steps */
  int32_t do_upsample;      /* iter_step >= important_begin_iter && n_importance > 0 (endosurf.py:85) */
  float cos_anneal_ratio;   /* get_cos_anneal_ratio(iter_step) (endosurf.py:215-219) */
  const float* variance;    /* device scalar: SingleVarianceNetwork.variance; inv_s = clip(exp(10 v),1e-6,1e6) */
  const float* t_vals;      /* device [n_samples]  = torch.linspace(0,1,n_samples) (endosurf.py:78) */
  const float* u_vals;      /* device [n_importance/up_sample_steps] (utils.py:170) */
  const float* t_rand;      /* device [R] jitter in [-.5,.5) when perturb (endosurf.py:81) or NULL */
  const float* z_override;  /* device [R,M]: skip sampling, run render_core on these z (endosurf.py:134) or NULL */
} es_render_params;

typedef struct es_render_out { /* the reference's 8-key dict (endosurf.py:123-132) + optional extras */
  float* color_map;        /* [R,3] */
  float* depth_map;        /* [R,1] */
  float* gradients_o;      /* [R,M,3] */
  float* gradient_o_error; /* [1] */
  float* weights;          /* [R,M] */
  float* weight_max;       /* [R,1] */
  float* cdf;              /* [R,M] */
  float* s_val;            /* [R,1] */
  float* z_vals;           /* [R,M]   optional (NULL to skip) */
  float* sdf;              /* [R,M]   optional */
  float* sampled_color;    /* [R,M,3] optional */
} es_render_out;

/* EndoSurfRenderer.render_rays (endosurf.py:60-132) / render_core (:134-213): the whole per-ray hot path as one
 * call.  rays [R,9] = o(3) d(3) near far time.  M = n_samples + (do_upsample ? n_importance : 0). */
int es_render_rays(es_ctx* ctx, const float* rays, int64_t n_rays, const es_render_params* p,
                   const es_render_out* out, void* stream);

/* Per-kernel device timing for bench.py's roofline: when enabled every fused MLP-chain launch is bracketed with
 * CUDA events on the launching stream.  es_profile_read synchronises the stream, returns the accumulated durations
 * since the last read and resets them.  kind: 0 = geometry chain (deform+sdf with tangent rows and feature layer),
 * 1 = colour chain, 2 = sdf query chain, 3/4/5 = reverse (training) chain of the deform / sdf / colour network,
 * 6 = input-adjoint launches, 7 = weight-gradient kernel, 8 = its reduce + weight-norm backward, 9 = compositing
 * forward / backward, 10 = 3-wide output-layer gradients. */
typedef struct es_profile {
  double ms[12];
  int64_t launches[12];
  int64_t points[12];
} es_profile;
int es_profile_enable(es_ctx* ctx, int32_t on);
int es_profile_read(es_ctx* ctx, es_profile* out, void* stream);

/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
int64_t es_launch_count(const es_ctx* ctx);

/* Weight-column order of the kernel's encoder chunks: fills out[64] with the reference input column each of the
 * 64 chunk columns reads for network `net` (or -1 for padding).  src: 1 deform enc, 2 sdf enc, 3 colour A, 4 colour B. */
int es_chunk_colmap(const es_ctx* ctx, int net, int src, int32_t* out64);

/* Self-test of the weight-gradient kernel on caller-built plane records (tests/test_gpu_wgrad.py):
 * out[256][64 n_b] = zbar^T in over n_tiles 128-row tiles, zbar_rec = [tile][4 chunks][16 KiB],
 * in_rec = [tile][n_b chunks][16 KiB]; bias_out[256] = column sums of zbar over all rows (bias_mode 1) or over the
 * primal rows 32Q + p of every tile (bias_mode 2).  Synchronises the stream. */
int es_wgrad_probe(es_ctx* ctx, const uint8_t* zbar_rec, const uint8_t* in_rec, int64_t n_tiles, int32_t n_b,
                   int32_t bias_mode, float* out, float* bias_out, void* stream);

/* Debug knobs: key 0 = ablation flags (ES_ABLATE builds), 1 / 2 = LBO / SBO bytes of the weight-gradient kernel's
 * MN-major operand descriptors (0 = built-in 128 / 2048; the parity test checks the convention on hardware). */
int es_debug_set(es_ctx* ctx, int32_t key, int32_t value);

/* Debug: pipeline trace of CTA 0 of the next fused-chain launches.  host_out == NULL arms it; a second call with a
 * buffer of 1 + 2*capacity_pairs int64 copies out [count, (clock64, code) ...] and disarms (tools/trace_chain.py). */
int es_debug_trace(es_ctx* ctx, int64_t* host_out, int64_t capacity_pairs);

/* Debug: tcgen05 issue-rate microbenchmark (tools/mma_bench.py); cfg15 = the 15 int32 fields of es::MmaBenchCfg. */
int es_mma_bench(es_ctx* ctx, const int32_t* cfg15, int32_t grid, int64_t* cycles_host);

/* tcgen05 self-test: d[128,256] = a[128,64] (fp16 bits) * b[256,64]^T (fp16 bits) through the same shared-memory
 * descriptors / TMA / TMEM path as the fused kernels.  lbo/sbo <= 0 selects the built-in strides. */
int es_umma_probe(es_ctx* ctx, const uint16_t* a, const uint16_t* b, float* d, int32_t a_lbo, int32_t a_sbo,
                  int32_t b_lbo, int32_t b_sbo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ENDOSURF_B200_H */
