// Weight gradients of the training backward on sm_100a tensor cores.
//
//   dW_l[out, in] = sum over rows  zbar_l[row, out] * a_l[row, in]
//
// zbar_l (adjoint of layer l's pre-activations, written by the reverse chains) and a_l (layer l's input, written by
// the forward training chains) live in HBM as fp16 plane records in the chains' ring-slot layout: a 64-column chunk of
// a 128-row tile is 16 KiB = [k-group of 8 columns][128 rows][8 fp16] (es_program.h LayerProg::dump).  Read as an
// MN-major no-swizzle UMMA operand that is exactly "MN = column, K = row" with 16-byte K rows, LBO = 128 B between
// 8-row K groups and SBO = 2048 B between 8-column MN groups - so one plain 1-D bulk copy per operand brings a tile
// straight from HBM into an MMA-ready stage, and the contraction over the 2 M rows of a batch is split-K over tiles:
//
//   CTA = (layer, 128 out columns, <= 256 in columns, slice of tiles)
//   warp 0      : TMA producer   - cp.async.bulk of the A (32 KiB) and B (<= 64 KiB) tile, 2 stages
//   warp 1      : MMA issuer     - 8 x tcgen05.mma M128 N<=256 K16 (both operands MN-major) per tile, fp32 in TMEM
//   warps 2..5  : bias gradients - column sums of the A tile's primal rows from shared memory while the MMAs run;
//                 afterwards they read the accumulator and write the CTA's partial [128][256] tile
// A second kernel sums the partial tiles of a layer deterministically, undoes the loss scale, and scatters through the
// layer's column map into the reference's [out, in] layout; a third applies the weight-norm backward
// (reference utils.py:57-58, W = g v / |v|) so the caller receives d loss / d (bias, weight_g, weight_v) directly.
//
// Replaces the autograd of the reference's Linear layers (loss.backward() at src/trainer/trainer_endosurf.py:94-104
// through src/renderer/endosurf.py:724-842 and the create_graph=True normals at :594,:612,:636-650).
#include "es_common.cuh"
#include "es_program.h"
#include "es_kernels.h"

namespace es {

constexpr int WG_STAGES = 2;
constexpr int WG_A_BYTES = 2 * SLOT_HALF_BYTES;   // 128 out columns of zbar: 2 chunks
constexpr int WG_B_BYTES = 4 * SLOT_HALF_BYTES;   // up to 256 in columns: 4 chunks
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_BAR_OFF = WG_STAGES * WG_STAGE_BYTES;
constexpr int WG_SMEM = WG_BAR_OFF + 64;
constexpr int WG_THREADS = 192;
constexpr int WG_LBO = 128;        // bytes between 8-row K groups
constexpr int WG_SBO = A_LBO;      // bytes between 8-column MN groups (2048)
static_assert(WG_SMEM <= 232448, "shared memory budget");

__device__ __forceinline__ void umma_f16_desc(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  umma_f16_ss(d_tmem, adesc, bdesc, idesc, accumulate);
}

// PAIRED: the two M halves of a (layer, N group, tile slice) run as the two CTAs of a cluster.  Each streams its own A
// tile and HALF of the shared B tile, multicast into both CTAs' stages, so B crosses L2 / HBM once per pair instead of
// once per CTA (ncu of the unpaired kernel: 58.5 GB of DRAM reads for 40 GB of records).
template <bool PAIRED>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const WgradItem* __restrict__ items, const __grid_constant__ WgradBases bases, float* __restrict__ partial,
             float* __restrict__ bias_partial, int lbo, int sbo, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const WgradItem it = items[blockIdx.x];
  const uint8_t* a_ptr = bases.p[it.a_buf] + it.a_off;
  const uint8_t* b_ptr = bases.p[it.b_buf] + it.b_off;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sm = smem_u32(smem);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + WG_BAR_OFF);
  uint64_t* bar_empty = bar_full + WG_STAGES;
  uint64_t* bar_done = bar_empty + WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);
  const bool bias = it.bias_mode != 0;
  const uint32_t rank = PAIRED ? cluster_ctarank() : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) {
      mbar_init(bar_full + i, 1);
      // stage free: this CTA's MMAs are done with it (+ the peer's, whose B half lands here too; + the bias warps)
      mbar_init(bar_empty + i, (PAIRED ? 2 : 1) + (bias ? 4 : 0));
    }
    mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  if constexpr (PAIRED) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = it.tile1 - it.tile0;
  const int n_cols = 64 * it.n_b;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t b_bytes = static_cast<uint32_t>(it.n_b) * SLOT_HALF_BYTES;
      for (int i = 0; i < n_tiles; ++i) {
        const int st = i % WG_STAGES;
        mbar_wait(bar_empty + st, ((i / WG_STAGES) & 1) ^ 1, err, 600);
        mbar_arrive_expect_tx(bar_full + st, WG_A_BYTES + b_bytes);
        const long long t = it.tile0 + i;
        uint8_t* stage = smem + st * WG_STAGE_BYTES;
        tma_bulk_g2s(stage, a_ptr + t * it.a_stride, WG_A_BYTES, bar_full + st);
        if constexpr (PAIRED) {
          const uint32_t half = b_bytes / 2;
          tma_bulk_g2s_mcast(stage + WG_A_BYTES + rank * half, b_ptr + t * it.b_stride + rank * half, half,
                             bar_full + st, 3);
        } else {
          tma_bulk_g2s(stage + WG_A_BYTES, b_ptr + t * it.b_stride, b_bytes, bar_full + st);
        }
      }
    }
  } else if (warp == 1) {
    // both operands MN-major (instruction descriptor bits 15 / 16), fp16 x fp16 -> fp32
    const uint32_t idesc = make_idesc_f16(TILE_ROWS, n_cols) | (1u << 15) | (1u << 16);
    const bool leader = elect_one_sync();
    uint32_t accum = 0;
    for (int i = 0; i < n_tiles; ++i) {
      const int st = i % WG_STAGES;
      mbar_wait(bar_full + st, (i / WG_STAGES) & 1, err, 610);
      tc_fence_after();
      if (leader) {
        const uint32_t sa = sm + st * WG_STAGE_BYTES;
        const uint32_t sb = sa + WG_A_BYTES;
#pragma unroll
        for (int ks = 0; ks < TILE_ROWS / 16; ++ks) {  // K = 16 rows per MMA = two 8-row K groups
          umma_f16_desc(tmem_base, make_smem_desc(sa + ks * 2 * WG_LBO, lbo, sbo),
                        make_smem_desc(sb + ks * 2 * WG_LBO, lbo, sbo), idesc, accum);
          accum = 1;
        }
        if constexpr (PAIRED) umma_commit_mcast(bar_empty + st, 3);
        else umma_commit(bar_empty + st);
      }
      __syncwarp();
    }
    if (leader) umma_commit(bar_done);
    __syncwarp();
  } else {
    const int w = warp - 2;  // 0..3
    if (bias) {
      // column sums of zbar over the rows that carry a bias (every row of a colour tile; the primal stream rows
      // 32Q + p of a geometry tile).  Warp w owns k-groups 4w..4w+3 (columns 32w..32w+31 of this M half); lanes run
      // along rows, 16-byte loads.
      float acc[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
      const int nrep = it.bias_mode == 1 ? 4 : 1;
      const int row_base = it.bias_mode == 1 ? lane : 32 * (lane >> 3) + (lane & 7);
      for (int i = 0; i < n_tiles; ++i) {
        const int st = i % WG_STAGES;
        mbar_wait(bar_full + st, (i / WG_STAGES) & 1, err, 620);
        const uint32_t sa = sm + st * WG_STAGE_BYTES;
        for (int rep = 0; rep < nrep; ++rep) {
          const int row = row_base + 32 * rep;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint32_t r0, r1, r2, r3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                         : "r"(sa + (4 * w + k) * WG_SBO + row * 16));
            const uint32_t rr[4] = {r0, r1, r2, r3};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rr[j]));
              acc[k][2 * j] += f.x;
              acc[k][2 * j + 1] += f.y;
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + st);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float v = acc[k][e];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) bias_partial[static_cast<size_t>(it.out) * TILE_ROWS + (4 * w + k) * 8 + e] = v;
        }
    }
    // accumulator -> partial tile.  TMEM lane quadrant = warp index % 4.
    mbar_wait(bar_done, 0, err, 630);
    tc_fence_after();
    const int quad = warp & 3;
    const int row = 32 * quad + lane;
    float* dst = partial + (static_cast<size_t>(it.out) * TILE_ROWS + row) * HID;
    for (int c0 = 0; c0 < n_cols; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(32 * quad) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        reinterpret_cast<float4*>(dst + c0)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  }
  tc_fence_before();
  if constexpr (PAIRED) cluster_sync_all();  // the peer may still multicast into this CTA's stages / barriers
  else __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// ---------------------------------------------------------------------------------------------- reduce + scatter
// One block per (job, 8 rows): sum the partial tiles of the job's slices in a fixed order, undo the loss scale and the
// layer's input scale, scatter through the column map into the effective-weight gradient [n_out][n_in].
__global__ void wgrad_reduce_kernel(const __grid_constant__ WgradJobs jobs, const float* __restrict__ partial,
                                    const float* __restrict__ bias_partial) {
  const WgradJob& jb = jobs.j[blockIdx.x];
  const float mul = jb.mul / (jb.scale ? __ldg(jb.scale) : 1.f);
  const int r0 = blockIdx.y * 8;
  for (int idx = threadIdx.x; idx < 8 * jb.n_cols; idx += blockDim.x) {
    const int r = r0 + idx / jb.n_cols, cidx = idx % jb.n_cols;
    const int wr = jb.row0 + r;
    const int col = jb.colmap ? __ldg(jb.colmap + cidx) : cidx;
    if (wr >= jb.n_out || col < 0) continue;
    float s = 0.f;
    for (int k = 0; k < jb.n_slices; ++k)
      s += partial[(static_cast<size_t>(jb.slot0 + k) * TILE_ROWS + r) * HID + cidx];
    jb.gw[static_cast<size_t>(wr + jb.row_off) * jb.n_in + col] = s * mul;
  }
  if (jb.gb && threadIdx.x < 8) {
    const int r = r0 + threadIdx.x;
    const int wr = jb.row0 + r;
    if (wr < jb.n_out) {
      float s = 0.f;
      for (int k = 0; k < jb.n_bias_slices; ++k) s += bias_partial[static_cast<size_t>(jb.slot0 + k) * TILE_ROWS + r];
      jb.gb[wr + jb.row_off] = s / (jb.scale ? __ldg(jb.scale) : 1.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------- 3-wide output layers
// out[j][col] = sum over rows adj[row][j] * h[row][col]  (j < 4) for the fp32 FFMA output layers (deform / colour
// tails: adj.xyz; sdf row: adj.w), h = the layer's input kept by the forward chain as 4 dump-only chunks.  CUDA cores:
// 4 x 256 outputs, HBM-bound on the one pass over h.  Block = 32 warps, warp = k-group (8 columns), lanes along rows.
__global__ void __launch_bounds__(1024)
smallm_wgrad_kernel(const uint8_t* __restrict__ plane, long long tile_stride, long long n_tiles,
                    const float* __restrict__ adj, int tangent, long long n_points, float* __restrict__ part) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[8][4];
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[e][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint8_t* base = plane + tile * tile_stride + warp * A_LBO;
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
      const int r = lane + 32 * rep;
      long long pt, logical;
      bool primal;
      if (tangent) {
        pt = tile * 32 + 8 * (r >> 5) + (r & 7);
        const int s = (r >> 3) & 3;
        logical = pt * 4 + s;
        primal = s == 0;
      } else {
        pt = tile * TILE_ROWS + r;
        logical = pt;
        primal = true;
      }
      if (pt >= n_points) continue;
      const float4 a = __ldg(reinterpret_cast<const float4*>(adj) + logical);
      const uint4 hv = __ldg(reinterpret_cast<const uint4*>(base + r * 16));
      const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[q]));
        acc[2 * q][0] = fmaf(a.x, f.x, acc[2 * q][0]);
        acc[2 * q][1] = fmaf(a.y, f.x, acc[2 * q][1]);
        acc[2 * q][2] = fmaf(a.z, f.x, acc[2 * q][2]);
        acc[2 * q][3] = fmaf(a.w, f.x, acc[2 * q][3]);
        acc[2 * q + 1][0] = fmaf(a.x, f.y, acc[2 * q + 1][0]);
        acc[2 * q + 1][1] = fmaf(a.y, f.y, acc[2 * q + 1][1]);
        acc[2 * q + 1][2] = fmaf(a.z, f.y, acc[2 * q + 1][2]);
        acc[2 * q + 1][3] = fmaf(a.w, f.y, acc[2 * q + 1][3]);
      }
      if (warp == 0 && primal) {
        bsum[0] += a.x; bsum[1] += a.y; bsum[2] += a.z; bsum[3] += a.w;
      }
    }
  }
  float* dst = part + static_cast<size_t>(blockIdx.x) * (4 * HID + 4);
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[e][j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) dst[j * HID + warp * 8 + e] = v;
    }
  if (warp == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = bsum[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) dst[4 * HID + j] = v;
    }
  }
}
// rows j0 .. j0+nj-1 of the per-block partials -> gw[row_off + j][256], gb[row_off + j]
__global__ void smallm_reduce_kernel(const float* __restrict__ part, int n_blocks, int j0, int nj, float* gw, float* gb,
                                     int row_off, int n_in) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < nj * HID) {
    const int j = idx / HID, c = idx % HID;
    float s = 0.f;
    for (int b = 0; b < n_blocks; ++b) s += part[static_cast<size_t>(b) * (4 * HID + 4) + (j0 + j) * HID + c];
    gw[static_cast<size_t>(row_off + j) * n_in + c] = s;
  } else if (idx < nj * HID + nj) {
    const int j = idx - nj * HID;
    float s = 0.f;
    for (int b = 0; b < n_blocks; ++b) s += part[static_cast<size_t>(b) * (4 * HID + 4) + 4 * HID + j0 + j];
    gb[row_off + j] = s;
  }
}

// ---------------------------------------------------------------------------------------------- weight norm
// W = g v / |v|_row (old-API weight_norm, reference utils.py:57-58).  One warp per row:
//   d g = <dW, v> / |v| ;   d v = (g / |v|) (dW - v <dW, v> / |v|^2)
__global__ void wn_backward_kernel(const __grid_constant__ WnLayers layers) {
  const WnLayer& L = layers.l[blockIdx.y];
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= L.n_out) return;
  const float* v = L.v + static_cast<size_t>(row) * L.n_in;
  const float* gw = L.gw + static_cast<size_t>(row) * L.n_in;
  float nn = 0.f, dot = 0.f;
  for (int i = lane; i < L.n_in; i += 32) {
    const float vi = v[i];
    nn = fmaf(vi, vi, nn);
    dot = fmaf(gw[i], vi, dot);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nn += __shfl_xor_sync(0xffffffffu, nn, o);
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
  }
  const float n = sqrtf(nn);
  const float g = L.g[row];
  const float k = g / n, d2 = dot / nn;
  float* gv = L.gv + static_cast<size_t>(row) * L.n_in;
  for (int i = lane; i < L.n_in; i += 32) gv[i] = k * (gw[i] - v[i] * d2);
  if (lane == 0) L.gg[row] = dot / n;
}

// effective weights W = g v / |v| (the forward fold; one warp per row)
__global__ void wn_fold_kernel(const __grid_constant__ WnLayers layers) {
  const WnLayer& L = layers.l[blockIdx.y];
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= L.n_out) return;
  const float* v = L.v + static_cast<size_t>(row) * L.n_in;
  float nn = 0.f;
  for (int i = lane; i < L.n_in; i += 32) nn = fmaf(v[i], v[i], nn);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
  const float k = L.g[row] / sqrtf(nn);
  float* w = L.w_eff + static_cast<size_t>(row) * L.n_in;
  for (int i = lane; i < L.n_in; i += 32) w[i] = v[i] * k;
}

// ---------------------------------------------------------------------------------------------- loss scale
__global__ void amax_kernel(const float* __restrict__ p, long long n, unsigned int* __restrict__ amax_bits) {
  float m = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    m = fmaxf(m, fabsf(p[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
}
// scale = 2^floor(log2(target / amax)): the largest adjoint entering a reverse chain becomes ~target, so that the fp16
// planes keep their relative precision whatever the loss normalisation (mean losses give adjoints of 1e-5..1e-8)
__global__ void scale_from_amax_kernel(const unsigned int* __restrict__ amax_bits, float* __restrict__ scale,
                                       float target) {
  const float a = __uint_as_float(*amax_bits);
  float s = 1.f;
  if (a > 0.f && a < 3.0e38f) s = exp2f(floorf(log2f(target / a)));
  if (!(s > 0.f) || s > 1.0e30f) s = 1.0e30f;
  *scale = s;
}

// ---------------------------------------------------------------------------------------------- launchers
cudaError_t launch_wgrad(const WgradItem* items_dev, int n_items, const WgradBases& bases, float* partial,
                         float* bias_partial, int lbo, int sbo, int* err, cudaStream_t stream, bool paired) {
  if (n_items <= 0) return cudaSuccess;
  if (paired) {
    // items 2k, 2k+1 = the two M halves of one (layer, N group, slice): one cluster
    if (n_items % 2) return cudaErrorInvalidValue;
    auto kern = wgrad_kernel<true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(n_items);
    cfg.blockDim = dim3(WG_THREADS);
    cfg.dynamicSmemBytes = WG_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, items_dev, bases, partial, bias_partial, lbo > 0 ? lbo : WG_LBO,
                              sbo > 0 ? sbo : WG_SBO, err);
  }
  auto kern = wgrad_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
  if (e != cudaSuccess) return e;
  kern<<<n_items, WG_THREADS, WG_SMEM, stream>>>(items_dev, bases, partial, bias_partial, lbo > 0 ? lbo : WG_LBO,
                                                 sbo > 0 ? sbo : WG_SBO, err);
  return cudaGetLastError();
}
cudaError_t launch_wgrad_reduce(const WgradJobs& jobs, const float* partial, const float* bias_partial,
                                cudaStream_t stream) {
  if (jobs.n <= 0) return cudaSuccess;
  wgrad_reduce_kernel<<<dim3(jobs.n, TILE_ROWS / 8), 256, 0, stream>>>(jobs, partial, bias_partial);
  return cudaGetLastError();
}
int smallm_blocks(long long n_tiles, int n_sms) {
  const long long want = 2LL * n_sms;
  return static_cast<int>(n_tiles < want ? (n_tiles > 0 ? n_tiles : 1) : want);
}
cudaError_t launch_smallm_wgrad(const uint8_t* plane, long long tile_stride, long long n_tiles, const float* adj,
                                int tangent, long long n_points, float* part, int n_blocks, cudaStream_t stream) {
  smallm_wgrad_kernel<<<n_blocks, 1024, 0, stream>>>(plane, tile_stride, n_tiles, adj, tangent, n_points, part);
  return cudaGetLastError();
}
cudaError_t launch_smallm_reduce(const float* part, int n_blocks, int j0, int nj, float* gw, float* gb, int row_off,
                                 int n_in, cudaStream_t stream) {
  const int total = nj * HID + nj;
  smallm_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(part, n_blocks, j0, nj, gw, gb, row_off, n_in);
  return cudaGetLastError();
}
cudaError_t launch_wn_backward(const WnLayers& layers, int max_rows, cudaStream_t stream) {
  if (layers.n <= 0) return cudaSuccess;
  wn_backward_kernel<<<dim3((max_rows + 7) / 8, layers.n), 256, 0, stream>>>(layers);
  return cudaGetLastError();
}
cudaError_t launch_wn_fold(const WnLayers& layers, int max_rows, cudaStream_t stream) {
  if (layers.n <= 0) return cudaSuccess;
  wn_fold_kernel<<<dim3((max_rows + 7) / 8, layers.n), 256, 0, stream>>>(layers);
  return cudaGetLastError();
}
cudaError_t launch_amax(const float* p, long long n, unsigned int* amax_bits, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const long long blocks = (n + 1023) / 1024;
  amax_kernel<<<static_cast<int>(blocks < 1184 ? blocks : 1184), 256, 0, stream>>>(p, n, amax_bits);
  return cudaGetLastError();
}
cudaError_t launch_scale_from_amax(const unsigned int* amax_bits, float* scale, float target, cudaStream_t stream) {
  scale_from_amax_kernel<<<1, 1, 0, stream>>>(amax_bits, scale, target);
  return cudaGetLastError();
}

}  // namespace es
