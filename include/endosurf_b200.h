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
int es_num_sms(const es_ctx* ctx);

/* Upload one network's EFFECTIVE weights W_l = g_l * v_l / ||v_l|| (reference utils.py:57-58, folded by the caller)
 * and biases: w[l] -> [out_l, in_l] row-major, b[l] -> [out_l], l = 0..n_layers-1, in the reference's own layout
 * (DeformNetwork/SDFNetwork/ColorNetwork.net[l], endosurf.py:713,762,817).  Packs them into fp16 hi/lo UMMA units. */
int es_load_network(es_ctx* ctx, int net, const float* const* w, const float* const* b, void* stream);

/* EndoSurfNet.get_sdf_from_observed_space (endosurf.py:570-579).
 * x [n,3]; t: element (i / t_div) * t_stride is the time of point i; sdf_out [n]. */
int es_sdf_query(es_ctx* ctx, const float* x, const float* t, int64_t t_div, int64_t t_stride, int64_t n,
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

/* ---- differentiable (training) path -------------------------------------------------------------------------------
 * The reference differentiates EndoSurfNet.forward with torch.autograd (create_graph=True at endosurf.py:594-658 so
 * that the eikonal / colour losses can back-propagate through the normals).  Here the forward keeps, per MMA layer,
 * the layer input of every row (primal + 3 tangent rows per point) as fp16 hi/lo planes ("stash"), and the reverse
 * pass runs the same fused tcgen05 chains on transposed weights, writing the adjoint of every forward
 * pre-activation ("zbar") so that weight gradients are plain [256 x rows] x [rows x 256] GEMMs.
 *   es_train_layout: out6 = {geometry stash rows, geometry stash slots, colour stash rows, colour stash slots,
 *                            zbar slots per reverse chain, geometry slot offset of the sdf layers}.
 *   Planes are uint16 (fp16 bits) [slots][rows][256].  Geometry rows follow the kernels' 128-row tiles of 32 points:
 *   point = 32*tile + 8*Q + p (Q = 0..3, p = 0..7), stream s (0 = primal, 1..3 = d/dx, d/dy, d/dz) of that point is
 *   row 128*tile + 32*Q + 8*s + p (one epilogue thread then owns the four streams of a point through the 16x256b
 *   TMEM fragment loads; endosurf_b200/training.py rows_from_points / points_from_rows convert).  Colour rows =
 *   point.  The adjoint input `adj` of es_point_backward stays indexed by (4*point + stream).                        */
int es_train_layout(const es_ctx* ctx, int64_t n, int64_t* out6);
/* Which fp16 lo planes the training launches write/read.  0 (default): only those the 1-term weight-gradient path
 * needs - the sdf stash slots (softplus gating) and the zbar slots of the layers that read the network input;
 * 1: every plane (needed when the caller forms 3-term weight gradients from hi and lo). */
int es_set_plane_mode(es_ctx* ctx, int32_t full_planes);
int es_point_forward_train(es_ctx* ctx, const float* x, const float* t, int64_t t_div, int64_t t_stride,
                           const float* dirs, int64_t dir_div, int64_t dir_stride, int64_t n, float* x_c, float* jac,
                           float* sdf, float* g_c, float* feat, float* rgb, uint16_t* geom_stash_hi,
                           uint16_t* geom_stash_lo, uint16_t* color_stash_hi, uint16_t* color_stash_lo, void* stream);
/* Reverse chain of one network (net = ES_NET_*).  adj: [logical rows][4] = adjoint of the row's 3-wide output
 * (deform: d/dDelta on primal rows, d/d(dDelta/dx_j) on tangent rows; colour: d/d(pre-sigmoid rgb)) in xyz and of the
 * sdf-row output in w (sdf chain: d/dsdf on primal rows, d/dg_c[j] on tangent rows).  adj_feat: d/dfeat [n,256]
 * (sdf chain).  zbar planes: [zbar slots][rows][256], slot m = adjoint of forward layer m's pre-activation. */
int es_point_backward(es_ctx* ctx, int net, int64_t n, const uint16_t* stash_hi, const uint16_t* stash_lo,
                      const float* adj, const float* adj_feat, uint16_t* zbar_hi, uint16_t* zbar_lo, void* stream);

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
 * 1 = colour chain, 2 = sdf query chain, 3/4/5 = reverse (training) chain of the deform / sdf / colour network. */
typedef struct es_profile {
  double ms[6];
  int64_t launches[6];
  int64_t points[6];
} es_profile;
int es_profile_enable(es_ctx* ctx, int32_t on);
int es_profile_read(es_ctx* ctx, es_profile* out, void* stream);

/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
int64_t es_launch_count(const es_ctx* ctx);

/* Weight-column order of the kernel's encoder chunks: fills out[64] with the reference input column each of the
 * 64 chunk columns reads for network `net` (or -1 for padding).  src: 1 deform enc, 2 sdf enc, 3 colour A, 4 colour B. */
int es_chunk_colmap(const es_ctx* ctx, int net, int src, int32_t* out64);

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
