// Internal launcher interface between es_api.cu (C ABI + orchestration) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "es_program.h"

namespace es {

enum { CHAIN_SDF = 0, CHAIN_COLOR = 1 };

// Global-memory operands of one fused MLP chain launch.  Unused pointers may be null.
struct ChainIO {
  long long n_points;
  int* err;  // device error word (0 = ok; watchdog / barrier site code otherwise)
  // geometry chains (deform + sdf)
  const float* x;   // [P,3] observed-space points
  const float* t;   // [P / t_div] times
  long long t_div;  // points per time value (1: per point, M: per ray)
  long long t_stride;  // floats between consecutive time values (9 when t aliases rays[:,8])
  float* out_xc;    // [P,3]
  float* out_jac;   // [P,9]   J[i][j] = d x_c_i / d x_j   (tangent mode only)
  float* out_sdf;   // [P]
  float* out_gc;    // [P,3]   d sdf / d x_c               (tangent mode only)
  float* out_feat;  // [P,256] geometry feature            (feat layer only)
  // geometry feature as fp16 hi/lo plane records of the COLOUR chain's tiles (128 points): [tile][chunk 0..3][hi, lo]
  // [16 KiB], each [k-group of 8 columns][128 rows][8 fp16] - exactly the colour chain's A-operand ring slot, so its
  // SRC_FEAT chunks become two bulk copies instead of row-per-thread 16-byte loads of the fp32 rows
  uint8_t* out_feat_rec;
  // colour chain
  const float* x_c;      // [P,3]
  const float* g_c;      // [P,3]
  const float* jac;      // [P,9] or null (identity)
  const float* dirs;     // view directions, row (p / dir_div) * dir_stride
  long long dir_div;
  long long dir_stride;
  const float* feat;     // [P,256]
  const uint8_t* feat_rec;  // or the plane records written through out_feat_rec (feat may then be null)
  float* out_rgb;        // [P,3]
  // training planes (see LayerProg::dump): records of 16 KiB chunks, [tile][chunks per tile][16 KiB]
  uint8_t* dump_hi;          // written by this launch (forward: activation stash; reverse: zbar, the adjoints of the
  uint8_t* dump_lo;          //   forward pre-activations); lo record only for the chunks that keep their lo half
  const uint8_t* gate_hi;    // reverse: the forward stash whose activations gate the adjoints
  const uint8_t* gate_lo;    //   (lo record indexed like the hi record; only read when ChainProg::gate_use_lo)
  const uint8_t* plane_hi;   // SRC_PLANE source records (input-adjoint launches: the zbar planes of the reverse chain)
  const uint8_t* plane_lo;
  const float* adj;          // reverse: [logical row][4] = (o.x, o.y, o.z, r)  (logical row = 4*pt+s or pt), unscaled
  const float* adj_feat;     // reverse sdf chain: d loss / d feat [P,256], unscaled
  const float* scale;        // reverse: device scalar, power of two; every adjoint entering the chain is multiplied by
                             //   it so that the fp16 planes keep their precision (null: 1)
  // input-adjoint launches (POST_INADJ_* / POST_FEAT_BAR); results are divided by *scale again
  float* feat_bar;           // POST_FEAT_BAR: [P,256]
  float* adj_sdf;            // POST_INADJ_COLOR: sdf-chain adjoint rows [4P][4]; .w of rows 4pt+1+j += d loss / d g_c[j]
  float* adj_deform;         // POST_INADJ_*: deform-chain adjoint rows [4P][4]; .xyz of row 4pt += d loss / d x_c,
                             //   of row 4pt+1+j += d loss / d J[:,j]   (null without a deformation network)
  unsigned int* amax_bits;   // POST_FEAT_BAR: atomicMax of the fp32 bit pattern of |feat_bar| (next chain's scale)
  // debug: when non-null, CTA 0 records (clock64, code) pairs: trace[0] = count, then pairs (tools/trace_chain.py)
  long long* trace;
  int store_hint;   // L2 policies, bit mask: 1 plane-record stores evict_first (every training launch), 2 the same for
                    // forward launches only, 4 weight units evict_last
  int debug_flags;  // perf experiments only (ES_DEBUG_FLAGS): 1 = no weight copies, 2 = no A stores, 4 = no MMAs
};

cudaError_t launch_mlp_chain(int chain, bool tangent, bool pair, const ChainProg& prog, const ChainIO& io,
                             int n_sms, cudaStream_t stream, bool bwd = false);
int mlp_chain_smem_bytes();

// ---- tcgen05 layout self-test (es_probe.cu): one 128x256x64 fp16 GEMM through the same descriptors
cudaError_t launch_umma_probe(const uint16_t* a_f16 /*[128][64]*/, const uint16_t* b_f16 /*[256][64]*/,
                              float* d /*[128][256]*/, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int* err,
                              cudaStream_t stream);

// ---- tcgen05 issue-rate microbenchmark (es_mmabench.cu)
struct MmaBenchCfg {
  int n;        // UMMA N (M = 128, K = 16 per instruction)
  int iters;    // outer iterations; each issues `ksteps` MMAs
  int ksteps;
  int a_layout, a_lbo, a_sbo, a_kadv, a_tiles, a_tile_bytes;  // descriptor layout code / strides / K advance in bytes
  int b_layout, b_lbo, b_sbo, b_kadv, b_tiles, b_tile_bytes;
};
cudaError_t launch_mma_bench(const MmaBenchCfg& cfg, int grid, long long* cycles_out, cudaStream_t stream);

// ---- weight packing (es_pack.cu): one launch per network over a job table
enum { PACK_FORWARD = 0, PACK_TRANSPOSED = 1, PACK_COPY = 2 };
struct PackJob {
  int kind;
  int n_out, n_in;     // source matrix w [n_out][n_in] (row stride n_in)
  int k_total;         // PACK_FORWARD: padded K (multiple of 32)
  int n_mma;           // PACK_TRANSPOSED: N of the operand (256 for the reverse chains)
  int n_valid;         // PACK_TRANSPOSED without a column table: rows n < n_valid read column n
  int n_copy;          // PACK_COPY: floats
  int pair;            // units in the CTA-pair layout (two contiguous half-row blocks)
  int work, block0;    // filled by launch_pack_jobs
  float scale;
  const float* w;
  const int* cols;     // PACK_FORWARD: [k_total] kernel K order -> source column; PACK_TRANSPOSED: [n_mma] or null
  uint8_t* units;
  float* dst;          // PACK_COPY
};
struct PackJobs {
  int n;
  PackJob j[48];
};
cudaError_t launch_pack_jobs(PackJobs& jobs, cudaStream_t stream);

// ---- weight gradients (es_wgrad.cu)
struct WgradBases {          // plane records of one backward call (the work list holds offsets, so it can be cached)
  const uint8_t* p[12];
};
struct WgradItem {           // one CTA of the split-K weight-gradient kernel
  long long a_off;           // zbar record: byte offset of (tile 0, first of the two adjacent chunks of this M half)
  long long b_off;           // layer-input record: byte offset of (tile 0, first of n_b adjacent chunks)
  long long a_stride;        // bytes between consecutive tiles of the record
  long long b_stride;
  int a_buf, b_buf;          // indices into WgradBases
  int tile0, tile1;          // tile range [tile0, tile1), never empty
  int n_b;                   // 1..4 input chunks (N = 64 n_b)
  int bias_mode;             // 0: no bias sums; 1: every row carries a bias (colour tiles); 2: primal rows 32Q + p only
  int out;                   // partial tile index: partial[out][128][256], bias_partial[out][128]
};
struct WgradJob {            // one (layer, M half, N group) of the reduce + scatter kernel
  int slot0, n_slices;       // partial tiles slot0 .. slot0 + n_slices - 1 are summed
  int n_bias_slices;         // the first n_bias_slices of them carry bias partials
  int n_cols;                // 64 n_b
  int row0;                  // first weight row of this M half (0 / 128)
  int row_off;               // row offset in gw / gb (1 for the feature rows of the sdf output layer)
  int n_out;                 // valid rows (row0 + r < n_out)
  int n_in;                  // row stride of gw
  const int* colmap;         // [n_cols] kernel K order -> reference input column (-1: padding)
  float* gw;                 // effective-weight gradient [rows][n_in]
  float* gb;                 // bias gradient or null
  float mul;                 // scale folded into the packed weights of this layer (1/sqrt 2 on skip layers)
  const float* scale;        // device scalar: loss scale of the reverse chain that wrote zbar (divided out)
};
struct WnLayer {             // one weight-normalised layer (reference utils.py:57-58)
  const float* v;            // weight_v [n_out][n_in]
  const float* g;            // weight_g [n_out]
  const float* gw;           // d loss / d W  (backward)
  float* gv;                 // d loss / d weight_v
  float* gg;                 // d loss / d weight_g
  float* w_eff;              // W = g v / |v|   (fold)
  int n_out, n_in;
};
struct WgradJobs {
  int n;
  WgradJob j[112];
};
struct WnLayers {
  int n;
  WnLayer l[32];
};
cudaError_t launch_wgrad(const WgradItem* items_dev, int n_items, const WgradBases& bases, float* partial,
                         float* bias_partial, int lbo, int sbo, int* err, cudaStream_t stream, bool paired);
cudaError_t launch_wgrad_reduce(const WgradJobs& jobs, const float* partial, const float* bias_partial,
                                cudaStream_t stream);
int smallm_blocks(long long n_tiles, int n_sms);
cudaError_t launch_smallm_wgrad(const uint8_t* plane, long long tile_stride, long long n_tiles, const float* adj,
                                int tangent, long long n_points, float* part, int n_blocks, cudaStream_t stream);
cudaError_t launch_smallm_reduce(const float* part, int n_blocks, int j0, int nj, float* gw, float* gb, int row_off,
                                 int n_in, cudaStream_t stream);
cudaError_t launch_wn_backward(const WnLayers& layers, int max_rows, cudaStream_t stream);
cudaError_t launch_wn_fold(const WnLayers& layers, int max_rows, cudaStream_t stream);
cudaError_t launch_amax(const float* p, long long n, unsigned int* amax_bits, cudaStream_t stream);
cudaError_t launch_scale_from_amax(const unsigned int* amax_bits, float* scale, float target, cudaStream_t stream);

// ---- per-ray kernels (es_rays.cu)
struct RayGeom {
  const float* rays;  // [R, 9]  o(3) d(3) near far time   (reference endosurf.py:64-65)
  long long n_rays;
};
cudaError_t launch_coarse_z(const RayGeom& rg, int n_samples, const float* t_vals /*[n]*/,
                            const float* t_rand /*[R] or null*/, float sample_dist, float* z /*[R,n]*/,
                            cudaStream_t stream);
cudaError_t launch_points_from_z(const RayGeom& rg, const float* z, int n, int mid, float sample_dist,
                                 float* pts /*[R*n,3]*/, cudaStream_t stream);
cudaError_t launch_upsample(const RayGeom& rg, const float* z, const float* sdf, int n, int n_imp,
                            const float* u_vals /*[n_imp]*/, float inv_s, float* new_z /*[R,n_imp]*/,
                            cudaStream_t stream);
cudaError_t launch_merge_z(long long n_rays, const float* z, const float* sdf, int n, const float* new_z,
                           const float* new_sdf /*null on the last step*/, int m, float* z_out, float* sdf_out,
                           cudaStream_t stream);
struct CompositeOut {
  float* color_map;    // [R,3]
  float* depth_map;    // [R]
  float* gradients_o;  // [R,M,3]
  float* weights;      // [R,M]
  float* cdf;          // [R,M]
  float* weight_max;   // [R]
  float* eik_partial;  // [R,2]  (sum relax*(|g|-1)^2, sum relax) per ray
  float* s_val;        // [R]    1 / inv_s
};
cudaError_t launch_composite(const RayGeom& rg, const float* z, int m, float sample_dist, const float* sdf,
                             const float* g_c, const float* jac /*or null*/, const float* rgb, const float* variance,
                             float cos_anneal, const CompositeOut& out, cudaStream_t stream);
cudaError_t launch_eikonal_reduce(const float* eik_partial, long long n_rays, float* out_scalar, float* out_den,
                                  cudaStream_t stream);
// adjoints of the compositing outputs (any may be null) and what the backward hands to the reverse chains
struct CompositeBwd {
  const float* color_bar;    // [R,3]
  const float* depth_bar;    // [R]
  const float* go_bar;       // [R,M,3]  adjoint of gradients_o
  const float* weights_bar;  // [R,M]
  const float* cdf_bar;      // [R,M]
  const float* sdf_bar;      // [R,M]    adjoint of the per-sample sdf returned as an extra
  const float* rgb_bar;      // [R,M,3]  adjoint of the per-sample colour returned as an extra
  const float* eik_bar;      // device scalar: adjoint of gradient_o_error
  const float* eik_den;      // device scalar: sum(relax) + 1e-6 of the forward
  float* adj_color;          // [P][4]   (rgb pre-sigmoid adjoint, 0)
  float* adj_sdf;            // [4P][4]  .w of row 4p = sdf adjoint, of row 4p+1+j = adjoint of g_c[j]
  float* adj_deform;         // [4P][4]  .xyz of row 4p+1+j = adjoint of J[:,j] (row 4p zeroed); null without deform
  float* invs_partial;       // [R]      d loss / d variance contributions
};
cudaError_t launch_composite_bwd(const RayGeom& rg, const float* z, int m, float sample_dist, const float* sdf,
                                 const float* g_c, const float* jac, const float* rgb, const float* variance,
                                 float cos_anneal, const CompositeBwd& b, cudaStream_t stream);
cudaError_t launch_sum_reduce(const float* part, long long n, float* out, int accumulate, cudaStream_t stream);
cudaError_t launch_point_adjoints(long long n, const float* rgb, const float* sdf_bar, const float* gc_bar,
                                  const float* jac_bar, const float* rgb_bar, float* adj_color, float* adj_sdf,
                                  float* adj_deform, cudaStream_t stream);
cudaError_t launch_fill_identity_jac(float* jac, long long n, cudaStream_t stream);
cudaError_t launch_grid_points(const float* lo3, const float* hi3, int res, int x0, int nx, float* pts,
                               cudaStream_t stream);

}  // namespace es
