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

// Every 64-column chunk is split into NPART column parts of PCOLS columns; one epilogue warp owns one
// (TMEM lane quadrant, part) pair, i.e. 32 rows x PCOLS columns of every chunk.  16 warps (4 per SM sub-partition)
// hide the TMEM-load / MUFU / shuffle latencies far better than 8 (measured: tensor pipe 31 % -> see profiles/).
constexpr int NPART = 4;
constexpr int PCOLS = CHUNK_K / NPART;  // 16
constexpr int N_EPI_WARPS = 4 * NPART;
constexpr int N_EPI_THREADS = N_EPI_WARPS * 32;
constexpr int N_THREADS = 64 + N_EPI_THREADS;

// dynamic shared memory carve-up
constexpr int SM_A_OFF = 0;
constexpr int SM_W_OFF = SM_A_OFF + NSLOT * SLOT_BYTES;    // 131072
constexpr int SM_XCH_OFF = SM_W_OFF + NSTAGE * UNIT_BYTES;  // 196608
constexpr int SM_XCH_BYTES = NPART * TILE_ROWS * 4 * 4;      // [part][row][4] floats
constexpr int SM_BIAS_OFF = SM_XCH_OFF + SM_XCH_BYTES;       // [MAXL + 1][256] fp32: every layer's bias + feat bias
constexpr int SM_BIAS_BYTES = (MAXL + 1) * HID * 4;
constexpr int PATCH_LD = 20;                                 // padded row length (floats) of an 8 x 16 patch
constexpr int PATCH_FLOATS = 8 * PATCH_LD;
constexpr int SM_PATCH_OFF = SM_BIAS_OFF + SM_BIAS_BYTES;    // per epilogue warp: 2 patches (z|h, act')
constexpr int SM_PATCH_BYTES = N_EPI_WARPS * 2 * PATCH_FLOATS * 4;
constexpr int SM_BAR_OFF = SM_PATCH_OFF + SM_PATCH_BYTES;
constexpr int N_BARS = 2 * NSLOT + 2 * NSTAGE + 4;
constexpr int SM_TMEM_OFF = SM_BAR_OFF + N_BARS * 8;
constexpr int SM_TOTAL = SM_TMEM_OFF + 16;

struct Bars {
  uint64_t* a_full;   // [NSLOT]  epilogue -> MMA   (count N_EPI_WARPS)
  uint64_t* a_empty;  // [NSLOT]  MMA commit -> epilogue
  uint64_t* w_full;   // [NSTAGE] TMA -> MMA
  uint64_t* w_empty;  // [NSTAGE] MMA commit -> TMA
  uint64_t* d_full;   // [2]      MMA commit -> epilogue
  uint64_t* d_empty;  // [2]      epilogue -> MMA  (count N_EPI_WARPS)
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
  float adj[4];  // reverse chains: (o.x, o.y, o.z, r) adjoint of this row's 3-wide output / sdf-row output
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

// does part PART of chunk SRC contain feature (var, freq, is_cos)?
__host__ __device__ constexpr bool part_has(int src, int part, int var, int freq, int is_cos) {
  for (int i = 0; i < PCOLS; ++i) {
    const Feat f = chunk_feat(src, PCOLS * part + i);
    if (f.var == var && f.freq == freq && f.is_cos == is_cos) return true;
  }
  return false;
}

// Fill v[PCOLS] with columns [PCOLS*PART, PCOLS*PART+PCOLS) of encoder chunk SRC (K order: es_program.h chunk_feat).
// `pos` is the position the encoding is taken of (x for the deform net, x_c otherwise).  Tangent rows (s>0) get
// the derivative of every feature wrt position component s-1.  All feature indices resolve at compile time, so
// only the sin/cos pairs this part needs are evaluated and everything lives in registers.
template <int SRC, int PART, bool TANGENT>
__device__ __forceinline__ void encode_part(float (&v)[PCOLS], const float (&pos)[3], const RowState& rs) {
  float var[10];
  var[0] = pos[0]; var[1] = pos[1]; var[2] = pos[2];
  var[3] = rs.t;
  var[4] = rs.g[0]; var[5] = rs.g[1]; var[6] = rs.g[2];
  var[7] = rs.dc[0]; var[8] = rs.dc[1]; var[9] = rs.dc[2];
  float sn[10][10], cs[10][10];
  static_for<0, PCOLS>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr Feat f = chunk_feat(SRC, PCOLS * PART + i);
    if constexpr (f.var >= 0 && f.freq >= 0) {
      // a sin column and its cos partner may land in different parts: evaluate the pair where either is needed,
      // but only once per part (the sin column triggers it if present, else the cos column)
      constexpr bool first = (f.is_cos == 0) || !part_has(SRC, PART, f.var, f.freq, 0);
      if constexpr (first) {
        // reference: torch.sin(x * 2^k), torch.cos(x * 2^k)  (encoder.py:47-50); x*2^k is exact in fp32
        sincosf(var[f.var] * static_cast<float>(1 << f.freq), &sn[f.var][f.freq], &cs[f.var][f.freq]);
      }
    }
  });
  const int s = rs.s;
  static_for<0, PCOLS>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr Feat f = chunk_feat(SRC, PCOLS * PART + i);
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

// split v[PCOLS] into fp16 hi/lo; store as column part `part` of A ring slot `slot_base` for row `row` (if non-null)
// and/or dump the same hi/lo halves to global planes at dhi/dlo (pointers to this row's first column; training stash)
__device__ __forceinline__ void emit_part(uint8_t* slot_base, int row, int part, const float (&v)[PCOLS],
                                          uint16_t* dhi, uint16_t* dlo) {
#pragma unroll
  for (int g = 0; g < PCOLS / 8; ++g) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[8 * g + 2 * j], v[8 * g + 2 * j + 1], hi[j], lo[j]);
    if (slot_base) {
      uint8_t* p = slot_base + ((PCOLS / 8) * part + g) * A_LBO + row * 16;
      *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(p + SLOT_HALF_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (dhi) {
      *reinterpret_cast<uint4*>(dhi + 8 * g) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(dlo + 8 * g) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// read PCOLS values back from fp16 hi/lo planes (value = hi + lo)
__device__ __forceinline__ void load_planes(const uint16_t* phi, const uint16_t* plo, float (&h)[PCOLS]) {
#pragma unroll
  for (int g = 0; g < PCOLS / 8; ++g) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(phi + 8 * g));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(plo + 8 * g));
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[j]));
      const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&bw[j]));
      h[8 * g + 2 * j] = fa.x + fb.x;
      h[8 * g + 2 * j + 1] = fa.y + fb.y;
    }
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
// debug pipeline trace (CTA 0 only, off unless ChainIO::trace is set).  Two recorder threads (the MMA issuer and
// epilogue warp 2 lane 0) write into separate halves of the buffer with private counters: no atomics on the path.
// layout: trace[0] = MMA count, trace[1] = EPI count, then 4000 (clock, code) pairs each.
__device__ __forceinline__ void trace_ev(long long* trace, int code, int who = 0, unsigned* counter = nullptr) {
  if (trace != nullptr && blockIdx.x == 0 && counter != nullptr) {
    const unsigned i = (*counter)++;
    if (i < 4000) {
      long long* base = trace + 2 + who * 8000;
      base[2 * i] = clock64();
      base[2 * i + 1] = code;
      trace[who] = i + 1;
    }
  }
}

struct EpiCtx {
  uint8_t* smem;
  Bars bars;
  uint32_t tmem_base;
  int* err;
  int row;     // 0..127 (TMEM lane)
  int part;    // which PCOLS columns of every 64-wide chunk this thread owns
  int lane;
  uint32_t ac;  // A-chunk counter (ring position), identical in all epilogue threads and the MMA warp
  uint32_t g;   // global MMA-layer counter (accumulator buffer = g & 1)
  uint32_t xk;  // cross-part exchange counter
  long long* trace;
  bool tr;  // this thread records trace events
  unsigned tcount;
  float* patch;  // this warp's two 8 x 16 activation patches (shared memory)
};

// sum a per-row float4 across the NPART column-part threads of the row (every one of them gets the total)
__device__ __forceinline__ float4 cross_part_sum(EpiCtx& c, float4 part) {
  float4* xch = reinterpret_cast<float4*>(c.smem + SM_XCH_OFF);
  xch[c.part * TILE_ROWS + c.row] = part;
  named_bar_sync(1, N_EPI_THREADS);
  float4 t = xch[c.row];
#pragma unroll
  for (int q = 1; q < NPART; ++q) {
    float4 b = xch[q * TILE_ROWS + c.row];
    t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
  }
  named_bar_sync(2, N_EPI_THREADS);  // everyone has read before the buffer is written again
  return t;
}

// Read this thread's PCOLS columns of 64-col block `blk` of accumulator buffer `buf`, add bias (row `bias_row` of the
// smem-staged bias table), activate.  TANGENT rows (s>0): no bias, multiplied by the primal row's activation
// derivative.
//  * Softplus + tangents: only 1 lane in 4 is a primal row, so running the exp/log/rcp chain on every lane wastes
//    3/4 of the MUFU/ALU issue slots.  The 8 primal rows of the warp park their 16 raw values in a per-warp smem patch,
//    every lane then activates 4 of the 128 values (its own point, columns 4s..4s+3), writes h and act' back, and
//    reads what its row needs (primal: h, tangent: act').  ~2x fewer instructions than the per-element shuffle form.
//  * ReLU + tangents: one shuffle of the primal pre-activation per element, then a select.
template <int ACT, bool TANGENT>
__device__ __forceinline__ void load_act(EpiCtx& c, int buf, int blk, int bias_row, int s, float (&v)[PCOLS]) {
  static_assert(PCOLS == 16, "the softplus patch path assigns 4 columns to each of the 4 lanes of a point");
  const int col0 = 64 * blk + PCOLS * c.part;
  const uint32_t taddr = c.tmem_base + (static_cast<uint32_t>(c.row & ~31) << 16) + buf * HID + col0;
  tmem_ld<PCOLS>(taddr, v);
  const float* bias = reinterpret_cast<const float*>(c.smem + SM_BIAS_OFF) + bias_row * HID + col0;
  if constexpr (TANGENT && ACT == ACT_SOFTPLUS100) {
    float* zh = c.patch;
    float* dp = c.patch + PATCH_FLOATS;
    const int p = c.lane >> 2;
    const float4 b4 = *reinterpret_cast<const float4*>(bias + 4 * s);
    tmem_ld_wait();
    if (s == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(zh + p * PATCH_LD + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    __syncwarp();
    const float4 z4 = *reinterpret_cast<const float4*>(zh + p * PATCH_LD + 4 * s);
    float hh[4], dd[4];
    activate<ACT>(z4.x + b4.x, hh[0], dd[0]);
    activate<ACT>(z4.y + b4.y, hh[1], dd[1]);
    activate<ACT>(z4.z + b4.z, hh[2], dd[2]);
    activate<ACT>(z4.w + b4.w, hh[3], dd[3]);
    *reinterpret_cast<float4*>(zh + p * PATCH_LD + 4 * s) = make_float4(hh[0], hh[1], hh[2], hh[3]);
    *reinterpret_cast<float4*>(dp + p * PATCH_LD + 4 * s) = make_float4(dd[0], dd[1], dd[2], dd[3]);
    __syncwarp();
    const float* src = (s == 0 ? zh : dp) + p * PATCH_LD;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 r = *reinterpret_cast<const float4*>(src + 4 * q);
      v[4 * q + 0] = (s == 0) ? r.x : r.x * v[4 * q + 0];
      v[4 * q + 1] = (s == 0) ? r.y : r.y * v[4 * q + 1];
      v[4 * q + 2] = (s == 0) ? r.z : r.z * v[4 * q + 2];
      v[4 * q + 3] = (s == 0) ? r.w : r.w * v[4 * q + 3];
    }
    __syncwarp();  // the patch is rewritten by the next call
  } else {
    float bj[PCOLS];
#pragma unroll
    for (int q = 0; q < PCOLS / 4; ++q) {
      const float4 bb = *reinterpret_cast<const float4*>(bias + 4 * q);
      bj[4 * q] = bb.x; bj[4 * q + 1] = bb.y; bj[4 * q + 2] = bb.z; bj[4 * q + 3] = bb.w;
    }
    tmem_ld_wait();
    if constexpr (TANGENT && ACT == ACT_RELU) {
#pragma unroll
      for (int i = 0; i < PCOLS; ++i) {
        const float zp = __shfl_sync(0xffffffffu, v[i] + bj[i], c.lane & ~3);  // primal pre-activation
        const float w = (s == 0) ? zp : v[i];
        v[i] = zp > 0.f ? w : 0.f;
      }
    } else {
      const float bsel = (!TANGENT || s == 0) ? 1.f : 0.f;
#pragma unroll
      for (int i = 0; i < PCOLS; ++i) {
        float z = fmaf(bj[i], bsel, v[i]);
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
}

// raw accumulator columns (no bias / activation): reverse chains
__device__ __forceinline__ void load_raw(EpiCtx& c, int buf, int blk, float (&v)[PCOLS]) {
  const int col0 = 64 * blk + PCOLS * c.part;
  tmem_ld<PCOLS>(c.tmem_base + (static_cast<uint32_t>(c.row & ~31) << 16) + buf * HID + col0, v);
  tmem_ld_wait();
}

// Activation backward for the reverse (training) chains.  u = adjoint of the post-activation values of a forward
// layer for this row; the forward stash holds those post-activation values h (primal rows) / hdot_j (tangent rows).
//   relu     : zbar = [h_primal > 0] * u                                (all rows; relu'' = 0)
//   softplus : sigma = 1 - exp(-100 h_primal)    (h = softplus(z)  =>  sigma(100 z) = 1 - exp(-100 h))
//              tangent rows: zdotbar_j = sigma * u_j
//              primal row  : zbar = sigma * u + 100 (1 - sigma) * sum_j hdot_j * u_j      (softplus'' = 100 s (1-s))
template <bool TANGENT>
__device__ __forceinline__ void bwd_gate(EpiCtx& c, const uint16_t* shi, const uint16_t* slo, int act, int s,
                                         float (&u)[PCOLS]) {
  float h[PCOLS];
  load_planes(shi, slo, h);
  if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < PCOLS; ++i) {
      const float hp = TANGENT ? __shfl_sync(0xffffffffu, h[i], c.lane & ~3) : h[i];
      u[i] = hp > 0.f ? u[i] : 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < PCOLS; ++i) {
      const float hp = TANGENT ? __shfl_sync(0xffffffffu, h[i], c.lane & ~3) : h[i];
      const float e = __expf(-100.f * hp);  // 1 - sigma
      if (TANGENT) {
        float ct = (s > 0) ? h[i] * u[i] : 0.f;
        ct += __shfl_xor_sync(0xffffffffu, ct, 1);
        ct += __shfl_xor_sync(0xffffffffu, ct, 2);
        const float su = (1.f - e) * u[i];
        u[i] = (s == 0) ? fmaf(100.f * e, ct, su) : su;
      } else {
        u[i] = (1.f - e) * u[i];
      }
    }
  }
}

template <bool TANGENT>
__device__ __forceinline__ void load_act_dyn(EpiCtx& c, int act, int buf, int blk, int bias, int s,
                                             float (&v)[PCOLS]) {
  if (act == ACT_RELU) load_act<ACT_RELU, TANGENT>(c, buf, blk, bias, s, v);
  else if (act == ACT_SOFTPLUS100) load_act<ACT_SOFTPLUS100, TANGENT>(c, buf, blk, bias, s, v);
  else load_act<ACT_NONE, TANGENT>(c, buf, blk, bias, s, v);
}

// acc[j] += sum_i v[i] * w[j][col0 + i]   (w row stride 256, uniform loads)
template <int NOUT>
__device__ __forceinline__ void dot_accum(const float (&v)[PCOLS], const float* __restrict__ w, int col0,
                                          float (&acc)[4]) {
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
    const float4* w4 = reinterpret_cast<const float4*>(w + j * HID + col0);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int q = 0; q < PCOLS / 4; ++q) {
      float4 ww = __ldg(w4 + q);
      a0 = fmaf(v[4 * q + 0], ww.x, a0);
      a1 = fmaf(v[4 * q + 1], ww.y, a1);
      a0 = fmaf(v[4 * q + 2], ww.z, a0);
      a1 = fmaf(v[4 * q + 3], ww.w, a1);
    }
    acc[j] += a0 + a1;
  }
}

__device__ __forceinline__ void wait_d_full(EpiCtx& c, uint32_t g_layer) {
  mbar_wait(&c.bars.d_full[g_layer & 1], (g_layer >> 1) & 1, c.err, 100 + static_cast<int>(g_layer & 1));
  tc_fence_after();
  if (c.tr) trace_ev(c.trace, 4000 + static_cast<int>(g_layer % 100), 1, &c.tcount);  // EPI: accumulator of layer g ready
}
__device__ __forceinline__ void release_d(EpiCtx& c, uint32_t g_layer) {
  tc_fence_before();
  __syncwarp();
  if (c.lane == 0) mbar_arrive(&c.bars.d_empty[g_layer & 1]);
}

// Consume the whole accumulator of global layer g_layer through a NOUT-wide fp32 output layer (no MMA):
// out = W_out . act(D + bias) summed over both column halves.  Returns the cross-half total in .x/.y/.z.
template <int NOUT, bool TANGENT>
__device__ __forceinline__ float4 tail_dot(EpiCtx& c, uint32_t g_layer, int act, int bias,
                                           const float* w_out, int s, uint16_t* st_hi = nullptr,
                                           uint16_t* st_lo = nullptr) {
  wait_d_full(c, g_layer);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int blk = 0; blk < 4; ++blk) {
    float v[PCOLS];
    load_act_dyn<TANGENT>(c, act, g_layer & 1, blk, bias, s, v);
    if (st_hi) {  // training stash of the output layer's input (pointers address this row's column 0)
      const int co = 64 * blk + PCOLS * c.part;
      emit_part(nullptr, 0, 0, v, st_hi + co, st_lo + co);
    }
    dot_accum<NOUT>(v, w_out, 64 * blk + PCOLS * c.part, acc);
  }
  release_d(c, g_layer);
  return cross_part_sum(c, make_float4(acc[0], acc[1], acc[2], acc[3]));
}

template <int SRC, bool TANGENT>
__device__ __forceinline__ void encode_dispatch(EpiCtx& c, float (&v)[PCOLS], const float (&pos)[3],
                                                const RowState& rs) {
  // c.part is warp-uniform: no divergence
  static_for<0, NPART>([&](auto pc) {
    constexpr int P = decltype(pc)::value;
    if (c.part == P) encode_part<SRC, P, TANGENT>(v, pos, rs);
  });
}

template <int CHAIN, bool TANGENT, bool BWD>
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
      mbar_init(&bars.a_full[i], N_EPI_WARPS);   // one elected arrive per epilogue warp
      mbar_init(&bars.a_empty[i], 1);
    }
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&bars.w_full[i], 1);
      mbar_init(&bars.w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.d_full[i], 1);
      mbar_init(&bars.d_empty[i], N_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  if constexpr (!BWD) {  // stage every layer's bias (+ the feature-layer bias in row MAXL) in shared memory
    float* bs = reinterpret_cast<float*>(smem + SM_BIAS_OFF);
    for (int i = threadIdx.x; i < prog.n_layers * HID; i += N_THREADS) bs[i] = __ldg(prog.bias + i);
    for (int i = threadIdx.x; i < HID; i += N_THREADS) bs[MAXL * HID + i] = __ldg(prog.feat_out_b + i);
  }
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
          if ((io.debug_flags & 1) && wc >= NSTAGE) {
            mbar_arrive(&bars.w_full[st]);
          } else {
            mbar_arrive_expect_tx(&bars.w_full[st], UNIT_BYTES);
            tma_bulk_g2s(smem + SM_W_OFF + st * UNIT_BYTES, prog.w_units + static_cast<size_t>(u) * UNIT_BYTES,
                         UNIT_BYTES, &bars.w_full[st]);
          }
          ++wc;
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================== MMA issuer
    // One thread feeds the tensor pipe; it must stay well ahead of the 128 cycles an M128 N256 K16 UMMA takes, so
    // the loop body is a handful of 64-bit adds on precomputed descriptors (a naive loop that rebuilt the
    // descriptors cost ~285 cycles per MMA and capped the tensor pipe at 31 %, profiles/r1_ncu_summary_v1.txt).
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(TILE_ROWS, HID);
      uint32_t wc = 0, ac = 0, g = 0;
      unsigned tcount = 0;
      const uint64_t a_desc0 = make_smem_desc(smem_u32(smem + SM_A_OFF), A_LBO, A_SBO);
      const uint64_t w_desc0 = make_smem_desc(smem_u32(smem + SM_W_OFF), B_LBO, B_SBO);
      constexpr uint64_t A_KS = (2 * A_LBO) >> 4;          // one K=16 step inside a slot plane
      constexpr uint64_t A_LO = SLOT_HALF_BYTES >> 4;      // hi plane -> lo plane
      constexpr uint64_t A_SB = (4 * A_LBO) >> 4;          // one 32-wide sub-block
      constexpr uint64_t W_KS = (2 * B_LBO) >> 4;
      const bool three = prog.n_terms == 3;
      const bool do_mma = !(io.debug_flags & 4);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < prog.n_layers; ++l, ++g) {
          const LayerProg& L = prog.layer[l];
          const int n_chunks = L.n_chunks;
          const uint32_t d_tmem = tmem_base + (g & 1) * HID;
          mbar_wait(&bars.d_empty[g & 1], ((g >> 1) & 1) ^ 1, err, 300 + static_cast<int>(g & 1));
          tc_fence_after();
          trace_ev(io.trace, 1000 + l, 0, &tcount);  // MMA: accumulator free, layer l starts
          uint32_t accum = 0;
          for (int ck = 0; ck < n_chunks; ++ck, ++ac) {
            const uint32_t slot = ac % NSLOT;
            const int nsub = L.nsub[ck];
            mbar_wait(&bars.a_full[slot], (ac / NSLOT) & 1, err, 310);
            tc_fence_after();
            trace_ev(io.trace, 2000 + l * 16 + ck, 0, &tcount);  // MMA: chunk ck of layer l available
            uint64_t a_hi = a_desc0 + static_cast<uint64_t>(slot * (SLOT_BYTES >> 4));
            for (int sb = 0; sb < nsub; ++sb, a_hi += A_SB) {
              // ---- hi weight unit: A_hi*B_hi and A_lo*B_hi
              {
                const uint32_t st = wc % NSTAGE;
                mbar_wait(&bars.w_full[st], (wc / NSTAGE) & 1, err, 320);
                tc_fence_after();
                const uint64_t wd = w_desc0 + static_cast<uint64_t>(st * (UNIT_BYTES >> 4));
                if (do_mma) {
                  umma_f16_ss(d_tmem, a_hi, wd, idesc, accum);
                  if (three) umma_f16_ss(d_tmem, a_hi + A_LO, wd, idesc, 1);
                  umma_f16_ss(d_tmem, a_hi + A_KS, wd + W_KS, idesc, 1);
                  if (three) umma_f16_ss(d_tmem, a_hi + A_LO + A_KS, wd + W_KS, idesc, 1);
                }
                accum = 1;
                umma_commit(&bars.w_empty[st]);
                ++wc;
              }
              // ---- lo weight unit: A_hi*B_lo
              if (three) {
                const uint32_t st = wc % NSTAGE;
                mbar_wait(&bars.w_full[st], (wc / NSTAGE) & 1, err, 321);
                tc_fence_after();
                const uint64_t wd = w_desc0 + static_cast<uint64_t>(st * (UNIT_BYTES >> 4));
                if (do_mma) {
                  umma_f16_ss(d_tmem, a_hi, wd, idesc, 1);
                  umma_f16_ss(d_tmem, a_hi + A_KS, wd + W_KS, idesc, 1);
                }
                umma_commit(&bars.w_empty[st]);
                ++wc;
              }
            }
            umma_commit(&bars.a_empty[slot]);
          }
          umma_commit(&bars.d_full[g & 1]);
          trace_ev(io.trace, 3000 + l, 0, &tcount);  // MMA: all MMAs of layer l issued
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
    c.part = (warp - 2) >> 2;
    c.row = (warp & 3) * 32 + lane;
    c.ac = 0;
    c.g = 0;
    c.xk = 0;
    c.patch = reinterpret_cast<float*>(smem + SM_PATCH_OFF) + (warp - 2) * 2 * PATCH_FLOATS;
    c.trace = io.trace;
    c.tr = (warp == 2 && lane == 0);
    c.tcount = 0;

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
        rs.adj[0] = rs.adj[1] = rs.adj[2] = rs.adj[3] = 0.f;
        if constexpr (BWD) {
          rs.x[0] = rs.x[1] = rs.x[2] = 0.f;
          rs.xc[0] = rs.xc[1] = rs.xc[2] = 0.f;
          if (rs.valid) {  // padding rows carry zero adjoints so they add nothing to the weight gradients
            const float4 a = __ldg(reinterpret_cast<const float4*>(io.adj) + (TANGENT ? rs.pt * 4 + rs.s : rs.pt));
            rs.adj[0] = a.x; rs.adj[1] = a.y; rs.adj[2] = a.z; rs.adj[3] = a.w;
          }
        } else if constexpr (CHAIN == CHAIN_COLOR) {
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
      const size_t row_global = static_cast<size_t>(tile) * TILE_ROWS + c.row;  // stash row

      for (int l = 0; l < prog.n_layers; ++l, ++c.g) {
        const LayerProg& L = prog.layer[l];
        const int bias_prev = l > 0 ? l - 1 : 0;  // row of the smem bias table
        const int act_prev = l > 0 ? prog.layer[l - 1].act : ACT_NONE;
        bool prev_waited = false;

        if (L.pre_op == PRE_DEFORM_TAIL) {
          // deform output layer (3 x 256, fp32 FFMA) -> x_c = x + delta ; tangent rows give dDelta/dx_{s-1}
          // training: the output layer's input goes to stash slot l (this layer has no SRC_PREV chunk of its own)
          uint16_t* sh = io.stash_hi ? io.stash_hi + (static_cast<size_t>(l) * io.stash_rows + row_global) * HID : nullptr;
          uint16_t* sl = io.stash_hi ? io.stash_lo + (static_cast<size_t>(l) * io.stash_rows + row_global) * HID : nullptr;
          float4 r = tail_dot<3, TANGENT>(c, c.g - 1, act_prev, bias_prev, prog.deform_out_w, rs.s, sh, sl);
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
            if (c.part == 0 && rs.valid && rs.s > 0 && io.out_jac) {
              // column j = s-1 of J = I + dDelta/dx ; J stored [i][j] row-major
#pragma unroll
              for (int i = 0; i < 3; ++i)
                io.out_jac[rs.pt * 9 + 3 * i + (rs.s - 1)] = dl[i] + ((i == rs.s - 1) ? 1.f : 0.f);
            }
          }
          if (c.part == 0 && rs.valid && rs.s == 0 && io.out_xc) {
#pragma unroll
            for (int i = 0; i < 3; ++i) io.out_xc[rs.pt * 3 + i] = rs.xc[i];
          }
          prev_waited = true;  // (already consumed and released)
        }

        int n_prev_left = 0;
        for (int ck = 0; ck < L.n_chunks; ++ck) n_prev_left += (L.src[ck] == SRC_PREV || L.src[ck] == SRC_BWD_PREV);

        for (int ck = 0; ck < L.n_chunks; ++ck, ++c.ac) {
          const uint32_t slot = c.ac % NSLOT;
          // Every layer starts after the previous layer's accumulator is complete, i.e. after every earlier MMA has
          // read its A slot: the first NSLOT chunks of a layer never have to wait for a free slot.
          if (ck >= NSLOT) mbar_wait(&bars.a_empty[slot], ((c.ac / NSLOT) & 1) ^ 1, err, 400);
          if (c.tr) trace_ev(io.trace, 6000 + l * 16 + ck, 1, &c.tcount);  // EPI: slot free
          uint8_t* slot_base = smem + SM_A_OFF + slot * SLOT_BYTES;
          float v[PCOLS];
          const int src = L.src[ck];
          bool active = true;
          uint16_t* dump_hi = nullptr;
          uint16_t* dump_lo = nullptr;
          const int col0 = 64 * L.arg[ck] + PCOLS * c.part;
          if (BWD && (src == SRC_BWD_PREV || src == SRC_BWD_OUTER3)) {
            if (src == SRC_BWD_PREV) {
              if (!prev_waited) {
                wait_d_full(c, c.g - 1);
                prev_waited = true;
              }
              load_raw(c, (c.g - 1) & 1, L.arg[ck], v);
              if (L.rank1) {
#pragma unroll
                for (int i = 0; i < PCOLS; ++i) v[i] = fmaf(rs.adj[3], __ldg(prog.sdf_out_w + col0 + i), v[i]);
              }
              if (--n_prev_left == 0) release_d(c, c.g - 1);
            } else {
#pragma unroll
              for (int i = 0; i < PCOLS; ++i)
                v[i] = rs.adj[0] * __ldg(prog.outer3_w + col0 + i) + rs.adj[1] * __ldg(prog.outer3_w + HID + col0 + i) +
                       rs.adj[2] * __ldg(prog.outer3_w + 2 * HID + col0 + i);
            }
            const size_t so = (static_cast<size_t>(L.stash_slot) * io.stash_rows + row_global) * HID + col0;
            bwd_gate<TANGENT>(c, io.stash_hi + so, io.stash_lo + so, L.bwd_act, rs.s, v);
            const size_t zo = (static_cast<size_t>(L.zbar_slot) * io.stash_rows + row_global) * HID + col0;
            dump_hi = io.zbar_hi + zo;
            dump_lo = io.zbar_lo + zo;
          } else if (BWD && src == SRC_ADJ_FEAT) {
            if (rs.valid && rs.s == 0) {
              const float4* f4 = reinterpret_cast<const float4*>(io.adj_feat + rs.pt * HID + col0);
#pragma unroll
              for (int q = 0; q < PCOLS / 4; ++q) {
                float4 f = __ldg(f4 + q);
                v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < PCOLS; ++i) v[i] = 0.f;
            }
          } else if (src == SRC_PREV) {
            if (!prev_waited) {
              wait_d_full(c, c.g - 1);
              prev_waited = true;
            }
            load_act_dyn<TANGENT>(c, act_prev, (c.g - 1) & 1, L.arg[ck], bias_prev, rs.s, v);
            if (L.side_dot) dot_accum<1>(v, prog.sdf_out_w, col0, sdf_acc);
            if (--n_prev_left == 0) release_d(c, c.g - 1);
            if (io.stash_hi) {  // training: keep this layer's input for the reverse pass / weight gradients
              const size_t so = (static_cast<size_t>(l) * io.stash_rows + row_global) * HID + col0;
              dump_hi = io.stash_hi + so;
              dump_lo = io.stash_lo + so;
            }
          } else if (src == SRC_ENC_DEFORM) {
            encode_dispatch<SRC_ENC_DEFORM, TANGENT>(c, v, rs.x, rs);
          } else if (src == SRC_ENC_SDF) {
            encode_dispatch<SRC_ENC_SDF, TANGENT>(c, v, rs.xc, rs);
          } else if (src == SRC_COLOR_A) {
            encode_dispatch<SRC_COLOR_A, false>(c, v, rs.xc, rs);
          } else if (src == SRC_COLOR_B) {
            if (PCOLS * c.part < 32) encode_dispatch<SRC_COLOR_B, false>(c, v, rs.xc, rs);  // 32-wide chunk
            else active = false;
          } else {  // SRC_FEAT
            const float4* f4 =
                reinterpret_cast<const float4*>(io.feat + rs.pt * HID + 64 * L.arg[ck] + PCOLS * c.part);
#pragma unroll
            for (int q = 0; q < PCOLS / 4; ++q) {
              float4 f = __ldg(f4 + q);
              v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
            }
          }
          if (c.tr) trace_ev(io.trace, 7000 + l * 16 + ck, 1, &c.tcount);  // EPI: values ready (tmem + math done)
          if (active) emit_part((io.debug_flags & 2) ? nullptr : slot_base, c.row, c.part, v, dump_hi, dump_lo);
          if (c.tr) trace_ev(io.trace, 8000 + l * 16 + ck, 1, &c.tcount);  // EPI: stores issued
          fence_proxy_async_smem();
          if (c.tr) trace_ev(io.trace, 9000 + l * 16 + ck, 1, &c.tcount);  // EPI: proxy fence done
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.a_full[slot]);  // 512 same-word arrivals serialise; 16 do not
          if (c.tr) trace_ev(io.trace, 5000 + l * 16 + ck, 1, &c.tcount);  // EPI: chunk ck of layer l written
        }

        if (L.side_dot) {
          // sdf row of the SDF output layer: primal rows -> sdf, tangent rows -> g_c[s-1]
          float4 r = cross_part_sum(c, make_float4(sdf_acc[0], 0.f, 0.f, 0.f));
          if (c.part == 0 && rs.valid) {
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
      const int bias_last = last;
      const int act_last = prog.layer[last].act;
      if (prog.post_op == POST_SDF_TAIL) {
        float4 r = tail_dot<1, TANGENT>(c, c.g - 1, act_last, bias_last, prog.sdf_out_w, rs.s);
        if (c.part == 0 && rs.valid) {
          if (rs.s == 0) {
            if (io.out_sdf) io.out_sdf[rs.pt] = r.x + __ldg(prog.sdf_out_b);
          } else if (io.out_gc) {
            io.out_gc[rs.pt * 3 + (rs.s - 1)] = r.x;
          }
        }
      } else if (BWD && prog.post_op == POST_BWD_DUMP) {
        wait_d_full(c, c.g - 1);
#pragma unroll 1
        for (int blk = 0; blk < 4; ++blk) {
          float v[PCOLS];
          const int col0 = 64 * blk + PCOLS * c.part;
          load_raw(c, (c.g - 1) & 1, blk, v);
          const size_t so = (static_cast<size_t>(prog.post_stash_slot) * io.stash_rows + row_global) * HID + col0;
          bwd_gate<TANGENT>(c, io.stash_hi + so, io.stash_lo + so, prog.post_bwd_act, rs.s, v);
          const size_t zo = (static_cast<size_t>(prog.post_zbar_slot) * io.stash_rows + row_global) * HID + col0;
          emit_part(nullptr, 0, 0, v, io.zbar_hi + zo, io.zbar_lo + zo);
        }
        release_d(c, c.g - 1);
      } else if (prog.post_op == POST_COLOR_TAIL) {
        uint16_t* sh = io.stash_hi ? io.stash_hi + (static_cast<size_t>(prog.n_layers) * io.stash_rows + row_global) * HID : nullptr;
        uint16_t* sl = io.stash_hi ? io.stash_lo + (static_cast<size_t>(prog.n_layers) * io.stash_rows + row_global) * HID : nullptr;
        float4 r = tail_dot<3, false>(c, c.g - 1, act_last, bias_last, prog.color_out_w, 0, sh, sl);
        if (c.part == 0 && rs.valid) {
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
          float v[PCOLS];
          // feat = D + bias (no activation); tangent rows are not needed
          load_act<ACT_NONE, false>(c, (c.g - 1) & 1, blk, MAXL, 0, v);  // row MAXL = feature-layer bias
          if (rs.valid && rs.s == 0) {
            float4* o4 = reinterpret_cast<float4*>(io.out_feat + rs.pt * HID + 64 * blk + PCOLS * c.part);
#pragma unroll
            for (int q = 0; q < PCOLS / 4; ++q)
              o4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
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
template <int CHAIN, bool TANGENT, bool BWD>
static cudaError_t launch_one(const ChainProg& prog, const ChainIO& io, int n_sms, cudaStream_t stream) {
  auto kern = mlp_chain_kernel<CHAIN, TANGENT, BWD>;
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
                             int n_sms, cudaStream_t stream, bool bwd) {
  (void)use_deform;  // the layer program already encodes whether a deformation network is present
  if (chain == CHAIN_COLOR)
    return bwd ? launch_one<CHAIN_COLOR, false, true>(prog, io, n_sms, stream)
               : launch_one<CHAIN_COLOR, false, false>(prog, io, n_sms, stream);
  if (chain == CHAIN_SDF) {
    if (bwd) return tangent ? launch_one<CHAIN_SDF, true, true>(prog, io, n_sms, stream) : cudaErrorInvalidValue;
    return tangent ? launch_one<CHAIN_SDF, true, false>(prog, io, n_sms, stream)
                   : launch_one<CHAIN_SDF, false, false>(prog, io, n_sms, stream);
  }
  return cudaErrorInvalidValue;
}

int mlp_chain_smem_bytes() { return SM_TOTAL; }

}  // namespace es
