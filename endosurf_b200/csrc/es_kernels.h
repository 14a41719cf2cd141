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
  // colour chain
  const float* x_c;      // [P,3]
  const float* g_c;      // [P,3]
  const float* jac;      // [P,9] or null (identity)
  const float* dirs;     // view directions, row (p / dir_div) * dir_stride
  long long dir_div;
  long long dir_stride;
  const float* feat;     // [P,256]
  float* out_rgb;        // [P,3]
  // training: activation stash (fp16 hi/lo planes, [slot][stash_rows][256], row = tile*128 + tile row)
  uint16_t* stash_hi;      // forward-train: written; reverse: read
  uint16_t* stash_lo;
  long long stash_rows;
  uint16_t* zbar_hi;       // reverse: adjoints of the forward pre-activations, same layout, [zbar slot]
  uint16_t* zbar_lo;
  const float* adj;        // reverse: [logical row][4] = (o.x, o.y, o.z, r)  (logical row = 4*pt+s or pt)
  const float* adj_feat;   // reverse sdf chain: d loss / d feat [P,256]
  // debug: when non-null, CTA 0 records (clock64, code) pairs: trace[0] = count, then pairs (tools/trace_chain.py)
  long long* trace;
  int debug_flags;  // perf experiments only (ES_DEBUG_FLAGS): 1 = no weight copies, 2 = no A stores, 4 = no MMAs
};

cudaError_t launch_mlp_chain(int chain, bool tangent, bool use_deform, const ChainProg& prog, const ChainIO& io,
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

// ---- weight packing (es_pack.cu)
// Gathers columns of an fp32 [n_out, n_in] matrix into the kernel's K order (colmap[k] = source column or -1),
// scales, zero-pads rows to 256 and K to 32*n_sub, splits into fp16 hi/lo and writes 16 KiB units
// [hi(sub0), lo(sub0), hi(sub1), lo(sub1), ...] in the canonical no-swizzle K-major UMMA layout.
cudaError_t launch_pack_layer(const float* w, int n_out, int n_in, const int* colmap_dev, int k_total, float scale,
                              uint8_t* units_out, cudaStream_t stream);
// Transposed pack for the reverse chains: B[n][k] = w[k][n] * scale for k < k_valid (forward out-features),
// n < n_valid (forward in-features kept: the first 256 / 204 columns); K = 256 (16 units).
cudaError_t launch_pack_layer_T(const float* w, int k_valid, int n_in_stride, int n_valid, float scale,
                                uint8_t* units_out, cudaStream_t stream);

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
cudaError_t launch_eikonal_reduce(const float* eik_partial, long long n_rays, float* out_scalar,
                                  cudaStream_t stream);

}  // namespace es
