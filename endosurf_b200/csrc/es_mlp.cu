// Fused per-point MLP chains of the EndoSurf renderer on sm_100a tensor cores.
//
// One persistent CTA per SM walks 128-row tiles through a chain of 256-wide layers:
//   warp 0      : TMA producer   - streams packed fp16 weight units (16 KiB, cp.async.bulk) L2 -> smem ring
//   warp 1      : MMA issuer     - tcgen05.mma (M128 N256 K16, fp16 x fp16 -> fp32 in TMEM), 3-term hi/lo split
//   warps 2..9  : epilogue       - tcgen05.ld the accumulator, bias + activation (+ forward-mode tangents),
//                                  split to fp16 hi/lo and write the next layer's A operand into the smem ring;
//                                  also evaluates positional encodings, the 3-wide output layers and the outputs.
// The accumulator is double buffered in TMEM (2 x 256 columns) so the MMA of layer l+1 starts on K chunk 0 while
// the epilogue is still converting chunks 1..3 of layer l.  Activations never touch HBM.
//
// Replaces (reference, relative to its repo root): src/renderer/endosurf.py:570-689 (EndoSurfNet queries),
// :692-842 (the three MLPs), src/renderer/encoder.py:40-54, and the autograd.grad calls at :594,:612,:636-650,
// which become forward-mode tangent rows riding through the same GEMMs.
#include "es_common.cuh"
#include "es_program.h"
#include "es_kernels.h"
#include <type_traits>

namespace es {

constexpr int N_EPI_WARPS = 8;
constexpr int N_EPI_THREADS = N_EPI_WARPS * 32;
constexpr int N_THREADS = 64 + N_EPI_THREADS;

// dynamic shared memory carve-up
constexpr int SM_A_OFF = 0;
constexpr int SM_W_OFF = SM_A_OFF + NSLOT * SLOT_BYTES;    // 131072
constexpr int SM_XCH_OFF = SM_W_OFF + NSTAGE * UNIT_BYTES;  // 196608
constexpr int SM_XCH_BYTES = 2 * 2 * TILE_ROWS * 4 * 4;     // [parity][half][row][4] floats = 8192
constexpr int SM_BAR_OFF = SM_XCH_OFF + SM_XCH_BYTES;
constexpr int N_BARS = 2 * NSLOT + 2 * NSTAGE + 4;
constexpr int SM_TMEM_OFF = SM_BAR_OFF + N_BARS * 8;
constexpr int SM_TOTAL = SM_TMEM_OFF + 16;

struct Bars {
  uint64_t* a_full;   // [NSLOT]  epilogue -> MMA   (count N_EPI_THREADS)
  uint64_t* a_empty;  // [NSLOT]  MMA commit -> epilogue
  uint64_t* w_full;   // [NSTAGE] TMA -> MMA
  uint64_t* w_empty;  // [NSTAGE] MMA commit -> TMA
  uint64_t* d_full;   // [2]      MMA commit -> epilogue
  uint64_t* d_empty;  // [2]      epilogue -> MMA  (count N_EPI_THREADS)
};

// ------------------------------------------------------------------------------------------------ activations
template <int ACT>
__device__ __forceinline__ void activate(float z, float& h, float& dh) {
  if (ACT == ACT_RELU) {
    h = fmaxf(z, 0.f);
    dh = z > 0.f ? 1.f : 0.f;
  } else if (ACT == ACT_SOFTPLUS100) {
    // softplus(beta=100): max(z,0) + log1p(exp(-100|z|))/100 ; derivative sigmoid(100 z)
    float e = __expf(-100.f * fabsf(z));
    float l = __logf(1.f + e);
    h = fmaxf(z, 0.f) + 0.01f * l;
    float r = __fdividef(1.f, 1.f + e);
    dh = z >= 0.f ? r : e * r;
  } else {
    h = z;
    dh = 1.f;
  }
}

// ------------------------------------------------------------------------------------------------ per-row state
struct RowState {
  float x[3];    // observed-space point
  float t;       // time
  float xc[3];   // canonical point (valid after the deform tail / = x without deform)
  float g[3];    // colour chain: canonical normal g_c
  float dc[3];   // colour chain: canonical view direction
  long long pt;  // global point index (clamped to a valid one)
  bool valid;    // point index < n_points
  int s;         // tangent mode: 0 primal, 1..3 tangent wrt x_{s-1}; plain mode: 0
};

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// Fill v[32] with columns [32*HALF, 32*HALF+32) of encoder chunk SRC (K order: es_program.h chunk_feat).
// `pos` is the position the encoding is taken of (x for the deform net, x_c otherwise).  Tangent rows (s>0) get
// the derivative of every feature wrt position component s-1.  All feature indices resolve at compile time, so
// only the sin/cos pairs this half needs are evaluated and everything lives in registers.
template <int SRC, int HALF, bool TANGENT>
__device__ __forceinline__ void encode_half(float (&v)[32], const float (&pos)[3], const RowState& rs) {
  float var[10];
  var[0] = pos[0]; var[1] = pos[1]; var[2] = pos[2];
  var[3] = rs.t;
  var[4] = rs.g[0]; var[5] = rs.g[1]; var[6] = rs.g[2];
  var[7] = rs.dc[0]; var[8] = rs.dc[1]; var[9] = rs.dc[2];
  float sn[10][10], cs[10][10];
  static_for<0, 32>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr Feat f = chunk_feat(SRC, 32 * HALF + i);
    if constexpr (f.var >= 0 && f.freq >= 0 && f.is_cos == 0) {
      // reference: torch.sin(x * 2^k), torch.cos(x * 2^k)  (encoder.py:47-50); x*2^k is exact in fp32
      sincosf(var[f.var] * static_cast<float>(1 << f.freq), &sn[f.var][f.freq], &cs[f.var][f.freq]);
    }
  });
  const int s = rs.s;
  static_for<0, 32>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr Feat f = chunk_feat(SRC, 32 * HALF + i);
    if constexpr (f.var < 0) {
      v[i] = 0.f;
    } else {
      float prim, der;
      if constexpr (f.freq < 0) {
        prim = var[f.var];
        der = 1.f;
      } else {
        constexpr float fr = static_cast<float>(1 << f.freq);
        prim = f.is_cos ? cs[f.var][f.freq] : sn[f.var][f.freq];
        der = f.is_cos ? -fr * sn[f.var][f.freq] : fr * cs[f.var][f.freq];
      }
      if constexpr (!TANGENT) {
        v[i] = prim;
      } else if constexpr (f.var < 3) {
        v[i] = (s == 0) ? prim : ((s - 1 == f.var) ? der : 0.f);
      } else {
        v[i] = (s == 0) ? prim : 0.f;
      }
    }
  });
}

// split v[32] into fp16 hi/lo and store as sub-block `half` of A ring slot `slot_base` for row `row`
__device__ __forceinline__ void store_a_half(uint8_t* slot_base, int row, int half, const float (&v)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[8 * g + 2 * j], v[8 * g + 2 * j + 1], hi[j], lo[j]);
    uint8_t* p = slot_base + (4 * half + g) * A_LBO + row * 16;
    *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(p + SLOT_HALF_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
struct EpiCtx {
  uint8_t* smem;
  Bars bars;
  uint32_t tmem_base;
  int* err;
  int row;     // 0..127 (TMEM lane)
  int half;    // which 32 columns of every 64-wide chunk this thread owns
  int lane;
  uint32_t ac;  // A-chunk counter (ring position), identical in all epilogue threads and the MMA warp
  uint32_t g;   // global MMA-layer counter (accumulator buffer = g & 1)
  uint32_t xk;  // cross-half exchange counter
};

// sum a per-row float4 across the two column-half threads of the row (both get the total)
__device__ __forceinline__ float4 cross_half_sum(EpiCtx& c, float4 part) {
  float4* xch = reinterpret_cast<float4*>(c.smem + SM_XCH_OFF);
  const int par = c.xk & 1;
  ++c.xk;
  xch[(par * 2 + c.half) * TILE_ROWS + c.row] = part;
  named_bar_sync(1, N_EPI_THREADS);
  float4 a = xch[(par * 2 + 0) * TILE_ROWS + c.row];
  float4 b = xch[(par * 2 + 1) * TILE_ROWS + c.row];
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// Read this thread's 32 columns of 64-col block `blk` of accumulator buffer `buf`, add bias, activate.
// TANGENT rows (s>0): no bias, multiplied by the primal row's activation derivative (quad shuffle).
template <int ACT, bool TANGENT>
__device__ __forceinline__ void load_act(EpiCtx& c, int buf, int blk, const float* __restrict__ bias, int s,
                                         float (&v)[32]) {
  const int col0 = 64 * blk + 32 * c.half;
  const uint32_t taddr = c.tmem_base + (static_cast<uint32_t>(c.row & ~31) << 16) + buf * HID + col0;
  tmem_ld32(taddr, v);
  tmem_ld_wait();
  const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 bb = __ldg(b4 + q);
    float bj[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = 4 * q + j;
      float z = v[i] + ((!TANGENT || s == 0) ? bj[j] : 0.f);
      float h, dh;
      activate<ACT>(z, h, dh);
      if (TANGENT) {
        float dhp = __shfl_sync(0xffffffffu, dh, c.lane & ~3);
        v[i] = (s == 0) ? h : dhp * v[i];
      } else {
        v[i] = h;
      }
    }
  }
}

template <bool TANGENT>
__device__ __forceinline__ void load_act_dyn(EpiCtx& c, int act, int buf, int blk, const float* bias, int s,
                                             float (&v)[32]) {
  if (act == ACT_RELU) load_act<ACT_RELU, TANGENT>(c, buf, blk, bias, s, v);
  else if (act == ACT_SOFTPLUS100) load_act<ACT_SOFTPLUS100, TANGENT>(c, buf, blk, bias, s, v);
  else load_act<ACT_NONE, TANGENT>(c, buf, blk, bias, s, v);
}

// acc[j] += sum_i v[i] * w[j][col0 + i]   (w row stride 256, uniform loads)
template <int NOUT>
__device__ __forceinline__ void dot_accum(const float (&v)[32], const float* __restrict__ w, int col0,
                                          float (&acc)[4]) {
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
    const float4* w4 = reinterpret_cast<const float4*>(w + j * HID + col0);
    float a = acc[j];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 ww = __ldg(w4 + q);
      a = fmaf(v[4 * q + 0], ww.x, a);
      a = fmaf(v[4 * q + 1], ww.y, a);
      a = fmaf(v[4 * q + 2], ww.z, a);
      a = fmaf(v[4 * q + 3], ww.w, a);
    }
    acc[j] = a;
  }
}

__device__ __forceinline__ void wait_d_full(EpiCtx& c, uint32_t g_layer) {
  mbar_wait(&c.bars.d_full[g_layer & 1], (g_layer >> 1) & 1, c.err, 100 + static_cast<int>(g_layer & 1));
  tc_fence_after();
}
__device__ __forceinline__ void release_d(EpiCtx& c, uint32_t g_layer) {
  tc_fence_before();
  mbar_arrive(&c.bars.d_empty[g_layer & 1]);
}

// Consume the whole accumulator of global layer g_layer through a NOUT-wide fp32 output layer (no MMA):
// out = W_out . act(D + bias) summed over both column halves.  Returns the cross-half total in .x/.y/.z.
template <int NOUT, bool TANGENT>
__device__ __forceinline__ float4 tail_dot(EpiCtx& c, uint32_t g_layer, int act, const float* bias,
                                           const float* w_out, int s) {
  wait_d_full(c, g_layer);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int blk = 0; blk < 4; ++blk) {
    float v[32];
    load_act_dyn<TANGENT>(c, act, g_layer & 1, blk, bias, s, v);
    dot_accum<NOUT>(v, w_out, 64 * blk + 32 * c.half, acc);
  }
  release_d(c, g_layer);
  return cross_half_sum(c, make_float4(acc[0], acc[1], acc[2], acc[3]));
}

template <int SRC, bool TANGENT>
__device__ __forceinline__ void encode_dispatch(EpiCtx& c, float (&v)[32], const float (&pos)[3],
                                                const RowState& rs) {
  if (c.half == 0) encode_half<SRC, 0, TANGENT>(v, pos, rs);
  else encode_half<SRC, 1, TANGENT>(v, pos, rs);
}

template <int CHAIN, bool TANGENT, bool USE_DEFORM>
__global__ void __launch_bounds__(N_THREADS, 1)
mlp_chain_kernel(const __grid_constant__ ChainProg prog, const __grid_constant__ ChainIO io) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  Bars bars;
  {
    uint64_t* b = reinterpret_cast<uint64_t*>(smem + SM_BAR_OFF);
    bars.a_full = b;
    bars.a_empty = b + NSLOT;
    bars.w_full = b + 2 * NSLOT;
    bars.w_empty = b + 2 * NSLOT + NSTAGE;
    bars.d_full = b + 2 * NSLOT + 2 * NSTAGE;
    bars.d_empty = b + 2 * NSLOT + 2 * NSTAGE + 2;
  }
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM_OFF);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(&bars.a_full[i], N_EPI_THREADS);
      mbar_init(&bars.a_empty[i], 1);
    }
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&bars.w_full[i], 1);
      mbar_init(&bars.w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.d_full[i], 1);
      mbar_init(&bars.d_empty[i], N_EPI_THREADS);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int rows_per_tile = TANGENT ? TILE_ROWS / 4 : TILE_ROWS;  // points per tile
  const long long n_tiles = (io.n_points + rows_per_tile - 1) / rows_per_tile;
  int* err = io.err;

  if (warp == 0) {
    // ============================================================== TMA producer
    if (lane == 0) {
      uint32_t wc = 0;
      const int step = (prog.n_terms == 3) ? 1 : 2;  // single-term mode skips the lo units (odd indices)
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int u = 0; u < prog.units_per_tile; u += step) {
          const uint32_t st = wc % NSTAGE;
          mbar_wait(&bars.w_empty[st], ((wc / NSTAGE) & 1) ^ 1, err, 200);
          mbar_arrive_expect_tx(&bars.w_full[st], UNIT_BYTES);
          tma_bulk_g2s(smem + SM_W_OFF + st * UNIT_BYTES, prog.w_units + static_cast<size_t>(u) * UNIT_BYTES,
                       UNIT_BYTES, &bars.w_full[st]);
          ++wc;
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(TILE_ROWS, HID);
      uint32_t wc = 0, ac = 0, g = 0;
      const uint32_t a_base = smem_u32(smem + SM_A_OFF);
      const uint32_t w_base = smem_u32(smem + SM_W_OFF);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < prog.n_layers; ++l, ++g) {
          const LayerProg& L = prog.layer[l];
          const uint32_t d_tmem = tmem_base + (g & 1) * HID;
          mbar_wait(&bars.d_empty[g & 1], ((g >> 1) & 1) ^ 1, err, 300 + static_cast<int>(g & 1));
          tc_fence_after();
          uint32_t accum = 0;
          for (int ck = 0; ck < L.n_chunks; ++ck, ++ac) {
            const uint32_t slot = ac % NSLOT;
            mbar_wait(&bars.a_full[slot], (ac / NSLOT) & 1, err, 310);
            tc_fence_after();
            const uint32_t a_slot = a_base + slot * SLOT_BYTES;
            for (int sb = 0; sb < L.nsub[ck]; ++sb) {
              // ---- hi weight unit: A_hi*B_hi and A_lo*B_hi
              {
                const uint32_t st = wc % NSTAGE;
                mbar_wait(&bars.w_full[st], (wc / NSTAGE) & 1, err, 320);
                tc_fence_after();
                const uint32_t w_st = w_base + st * UNIT_BYTES;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  const uint64_t bd = make_smem_desc(w_st + ks * 2 * B_LBO, B_LBO, B_SBO);
                  const uint32_t a_off = (sb * 4 + ks * 2) * A_LBO;
                  umma_f16_ss(d_tmem, make_smem_desc(a_slot + a_off, A_LBO, A_SBO), bd, idesc, accum);
                  accum = 1;
                  if (prog.n_terms == 3)
                    umma_f16_ss(d_tmem, make_smem_desc(a_slot + SLOT_HALF_BYTES + a_off, A_LBO, A_SBO), bd, idesc,
                                 1);
                }
                umma_commit(&bars.w_empty[st]);
                ++wc;
              }
              // ---- lo weight unit: A_hi*B_lo
              if (prog.n_terms == 3) {
                const uint32_t st = wc % NSTAGE;
                mbar_wait(&bars.w_full[st], (wc / NSTAGE) & 1, err, 321);
                tc_fence_after();
                const uint32_t w_st = w_base + st * UNIT_BYTES;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  const uint64_t bd = make_smem_desc(w_st + ks * 2 * B_LBO, B_LBO, B_SBO);
                  const uint32_t a_off = (sb * 4 + ks * 2) * A_LBO;
                  umma_f16_ss(d_tmem, make_smem_desc(a_slot + a_off, A_LBO, A_SBO), bd, idesc, 1);
                }
                umma_commit(&bars.w_empty[st]);
                ++wc;
              }
            }
            umma_commit(&bars.a_empty[slot]);
          }
          umma_commit(&bars.d_full[g & 1]);
        }
      }
    }
  } else {
    // ============================================================== epilogue warps
    EpiCtx c;
    c.smem = smem;
    c.bars = bars;
    c.tmem_base = tmem_base;
    c.err = err;
    c.lane = lane;
    c.half = (warp - 2) >> 2;
    c.row = (warp & 3) * 32 + lane;
    c.ac = 0;
    c.g = 0;
    c.xk = 0;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      // ---------------------------------------------------------- row state
      RowState rs;
      {
        long long p = TANGENT ? tile * (TILE_ROWS / 4) + (c.row >> 2) : tile * TILE_ROWS + c.row;
        rs.valid = p < io.n_points;
        rs.pt = rs.valid ? p : io.n_points - 1;
        rs.s = TANGENT ? (c.row & 3) : 0;
        rs.t = 0.f;
        rs.g[0] = rs.g[1] = rs.g[2] = 0.f;
        rs.dc[0] = rs.dc[1] = rs.dc[2] = 0.f;
        if constexpr (CHAIN == CHAIN_COLOR) {
          const float* xc = io.x_c + rs.pt * 3;
          const float* gc = io.g_c + rs.pt * 3;
          const float* J = io.jac + rs.pt * 9;
          const float* d = io.dirs + (rs.pt / io.dir_div) * io.dir_stride;
          float dd[3] = {__ldg(d), __ldg(d + 1), __ldg(d + 2)};
          float dcn[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            rs.xc[i] = __ldg(xc + i);
            rs.x[i] = rs.xc[i];
            rs.g[i] = __ldg(gc + i);
            // d_c = J d  (reference endosurf.py:684: bmm(pts_jacobian, d)), J[i][j] = d x_c_i / d x_j
            dcn[i] = io.jac ? (__ldg(J + 3 * i) * dd[0] + __ldg(J + 3 * i + 1) * dd[1] + __ldg(J + 3 * i + 2) * dd[2])
                            : dd[i];
          }
          float nrm = sqrtf(dcn[0] * dcn[0] + dcn[1] * dcn[1] + dcn[2] * dcn[2]) + 1e-10f;
#pragma unroll
          for (int i = 0; i < 3; ++i) rs.dc[i] = dcn[i] / nrm;
        } else {
          const float* xp = io.x + rs.pt * 3;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            rs.x[i] = __ldg(xp + i);
            rs.xc[i] = rs.x[i];
          }
          rs.t = __ldg(io.t + (rs.pt / io.t_div) * io.t_stride);
        }
      }
      float sdf_acc[4] = {0.f, 0.f, 0.f, 0.f};

      for (int l = 0; l < prog.n_layers; ++l, ++c.g) {
        const LayerProg& L = prog.layer[l];
        const float* bias_prev = prog.bias + static_cast<size_t>(l > 0 ? l - 1 : 0) * HID;
        const int act_prev = l > 0 ? prog.layer[l - 1].act : ACT_NONE;
        bool prev_waited = false;

        if (L.pre_op == PRE_DEFORM_TAIL) {
          // deform output layer (3 x 256, fp32 FFMA) -> x_c = x + delta ; tangent rows give dDelta/dx_{s-1}
          float4 r = tail_dot<3, TANGENT>(c, c.g - 1, act_prev, bias_prev, prog.deform_out_w, rs.s);
          float dl[3] = {r.x, r.y, r.z};
          if (!TANGENT) {
#pragma unroll
            for (int i = 0; i < 3; ++i) rs.xc[i] = rs.x[i] + dl[i] + __ldg(prog.deform_out_b + i);
          } else {
            float prim[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              prim[i] = rs.x[i] + dl[i] + __ldg(prog.deform_out_b + i);
              rs.xc[i] = __shfl_sync(0xffffffffu, prim[i], lane & ~3);
            }
            if (c.half == 0 && rs.valid && rs.s > 0 && io.out_jac) {
              // column j = s-1 of J = I + dDelta/dx ; J stored [i][j] row-major
#pragma unroll
              for (int i = 0; i < 3; ++i)
                io.out_jac[rs.pt * 9 + 3 * i + (rs.s - 1)] = dl[i] + ((i == rs.s - 1) ? 1.f : 0.f);
            }
          }
          if (c.half == 0 && rs.valid && rs.s == 0 && io.out_xc) {
#pragma unroll
            for (int i = 0; i < 3; ++i) io.out_xc[rs.pt * 3 + i] = rs.xc[i];
          }
          prev_waited = true;  // (already consumed and released)
        }

        int n_prev_left = 0;
        for (int ck = 0; ck < L.n_chunks; ++ck) n_prev_left += (L.src[ck] == SRC_PREV);

        for (int ck = 0; ck < L.n_chunks; ++ck, ++c.ac) {
          const uint32_t slot = c.ac % NSLOT;
          mbar_wait(&bars.a_empty[slot], ((c.ac / NSLOT) & 1) ^ 1, err, 400);
          uint8_t* slot_base = smem + SM_A_OFF + slot * SLOT_BYTES;
          float v[32];
          const int src = L.src[ck];
          bool active = true;
          if (src == SRC_PREV) {
            if (!prev_waited) {
              wait_d_full(c, c.g - 1);
              prev_waited = true;
            }
            load_act_dyn<TANGENT>(c, act_prev, (c.g - 1) & 1, L.arg[ck], bias_prev, rs.s, v);
            if (L.side_dot) dot_accum<1>(v, prog.sdf_out_w, 64 * L.arg[ck] + 32 * c.half, sdf_acc);
            if (--n_prev_left == 0) release_d(c, c.g - 1);
          } else if (src == SRC_ENC_DEFORM) {
            encode_dispatch<SRC_ENC_DEFORM, TANGENT>(c, v, rs.x, rs);
          } else if (src == SRC_ENC_SDF) {
            encode_dispatch<SRC_ENC_SDF, TANGENT>(c, v, rs.xc, rs);
          } else if (src == SRC_COLOR_A) {
            encode_dispatch<SRC_COLOR_A, false>(c, v, rs.xc, rs);
          } else if (src == SRC_COLOR_B) {
            if (c.half == 0) encode_half<SRC_COLOR_B, 0, false>(v, rs.xc, rs);
            else active = false;
          } else {  // SRC_FEAT
            const float4* f4 = reinterpret_cast<const float4*>(io.feat + rs.pt * HID + 64 * L.arg[ck] + 32 * c.half);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 f = __ldg(f4 + q);
              v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
            }
          }
          if (active) store_a_half(slot_base, c.row, c.half, v);
          fence_proxy_async_smem();
          mbar_arrive(&bars.a_full[slot]);
        }

        if (L.side_dot) {
          // sdf row of the SDF output layer: primal rows -> sdf, tangent rows -> g_c[s-1]
          float4 r = cross_half_sum(c, make_float4(sdf_acc[0], 0.f, 0.f, 0.f));
          if (c.half == 0 && rs.valid) {
            if (rs.s == 0) {
              if (io.out_sdf) io.out_sdf[rs.pt] = r.x + __ldg(prog.sdf_out_b);
            } else if (io.out_gc) {
              io.out_gc[rs.pt * 3 + (rs.s - 1)] = r.x;
            }
          }
        }
      }

      // ---------------------------------------------------------- post op: consume the last accumulator
      const int last = prog.n_layers - 1;
      const float* bias_last = prog.bias + static_cast<size_t>(last) * HID;
      const int act_last = prog.layer[last].act;
      if (prog.post_op == POST_SDF_TAIL) {
        float4 r = tail_dot<1, TANGENT>(c, c.g - 1, act_last, bias_last, prog.sdf_out_w, rs.s);
        if (c.half == 0 && rs.valid) {
          if (rs.s == 0) {
            if (io.out_sdf) io.out_sdf[rs.pt] = r.x + __ldg(prog.sdf_out_b);
          } else if (io.out_gc) {
            io.out_gc[rs.pt * 3 + (rs.s - 1)] = r.x;
          }
        }
      } else if (prog.post_op == POST_COLOR_TAIL) {
        float4 r = tail_dot<3, false>(c, c.g - 1, act_last, bias_last, prog.color_out_w, 0);
        if (c.half == 0 && rs.valid) {
          float o[3] = {r.x, r.y, r.z};
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            float z = o[i] + __ldg(prog.color_out_b + i);
            io.out_rgb[rs.pt * 3 + i] = 1.f / (1.f + expf(-z));
          }
        }
      } else if (prog.post_op == POST_FEAT_OUT) {
        wait_d_full(c, c.g - 1);
#pragma unroll 1
        for (int blk = 0; blk < 4; ++blk) {
          float v[32];
          // feat = D + bias (no activation); tangent rows are not needed
          load_act<ACT_NONE, false>(c, (c.g - 1) & 1, blk, prog.feat_out_b, 0, v);
          if (rs.valid && rs.s == 0) {
            float4* o4 = reinterpret_cast<float4*>(io.out_feat + rs.pt * HID + 64 * blk + 32 * c.half);
#pragma unroll
            for (int q = 0; q < 8; ++q) o4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
        release_d(c, c.g - 1);
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ launchers
template <int CHAIN, bool TANGENT, bool USE_DEFORM>
static cudaError_t launch_one(const ChainProg& prog, const ChainIO& io, int n_sms, cudaStream_t stream) {
  auto kern = mlp_chain_kernel<CHAIN, TANGENT, USE_DEFORM>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
  if (e != cudaSuccess) return e;
  const int pts_per_tile = TANGENT ? TILE_ROWS / 4 : TILE_ROWS;
  long long n_tiles = (io.n_points + pts_per_tile - 1) / pts_per_tile;
  if (n_tiles <= 0) return cudaSuccess;
  int grid = static_cast<int>(n_tiles < n_sms ? n_tiles : n_sms);
  kern<<<grid, N_THREADS, SM_TOTAL, stream>>>(prog, io);
  return cudaGetLastError();
}

cudaError_t launch_mlp_chain(int chain, bool tangent, bool use_deform, const ChainProg& prog, const ChainIO& io,
                             int n_sms, cudaStream_t stream) {
  if (chain == CHAIN_COLOR) return launch_one<CHAIN_COLOR, false, true>(prog, io, n_sms, stream);
  if (chain == CHAIN_SDF) {
    if (tangent) return use_deform ? launch_one<CHAIN_SDF, true, true>(prog, io, n_sms, stream)
                                   : launch_one<CHAIN_SDF, true, false>(prog, io, n_sms, stream);
    return use_deform ? launch_one<CHAIN_SDF, false, true>(prog, io, n_sms, stream)
                      : launch_one<CHAIN_SDF, false, false>(prog, io, n_sms, stream);
  }
  return cudaErrorInvalidValue;
}

int mlp_chain_smem_bytes() { return SM_TOTAL; }

}  // namespace es
