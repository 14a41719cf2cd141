// Fused per-point MLP chains of the EndoSurf renderer on sm_100a tensor cores.
//
// One persistent CTA per SM walks 128-row tiles through a chain of 256-wide layers:
//   warp 0      : TMA producer   - streams packed fp16 weight units (16 KiB, cp.async.bulk) L2 -> smem ring
//   warp 1      : MMA issuer     - tcgen05.mma (M128 N256 K16, fp16 x fp16 -> fp32 in TMEM), 3-term hi/lo split
//   warps 2..17 : epilogue       - tcgen05.ld the accumulator, bias + activation (+ forward-mode tangents),
//                                  split to fp16 hi/lo and write the next layer's A operand into the smem ring;
//                                  also evaluates positional encodings, the 3-wide output layers and the outputs.
// The accumulator is double buffered in TMEM (2 x 256 columns) so the MMA of layer l+1 starts on K chunk 0 while
// the epilogue is still converting chunks 1..3 of layer l.  Activations never touch HBM.
//
// Tangent mode (geometry chains): a tile is 32 points x 4 streams (primal + d/dx_0..2).  Stream s of point
// (quadrant Q, p) lives in tile row 32Q + 8s + p, so that ONE epilogue thread receives all four streams of a point
// for the same accumulator columns from two tcgen05.ld.16x256b loads (the m16n8 fragment layout: TMEM lanes L and
// L+8 per thread, both 16-lane halves of the warp's quadrant).  The chain rule  hdot = act'(z) * zdot  is then
// thread-local: one activation per (point, column) instead of one per row, no shuffles, no shared-memory exchange.
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
// (TMEM lane quadrant, part) pair, i.e. 32 rows x PCOLS columns of every chunk.
constexpr int NPART = 4;
constexpr int PCOLS = CHUNK_K / NPART;  // 16
constexpr int N_EPI_WARPS = 4 * NPART;
constexpr int N_EPI_THREADS = N_EPI_WARPS * 32;
constexpr int N_THREADS = 64 + N_EPI_THREADS;
constexpr int TILE_PTS_T = TILE_ROWS / 4;  // points per tile in tangent mode

// dynamic shared memory carve-up
constexpr int SM_A_OFF = 0;
constexpr int SM_W_OFF = SM_A_OFF + NSLOT * SLOT_BYTES;
constexpr int SM_XCH_OFF = SM_W_OFF + NSTAGE * UNIT_BYTES;
constexpr int SM_XCH_BYTES = NPART * TILE_ROWS * 4 * 4;      // plain: [part][row][4] floats; tangent: [part][pt][<=12]
constexpr int SM_BIAS_OFF = SM_XCH_OFF + SM_XCH_BYTES;       // [MAXL + 1][256] fp32: every layer's bias + feat bias
constexpr int SM_BIAS_BYTES = (MAXL + 1) * HID * 4;
constexpr int SM_BAR_OFF = SM_BIAS_OFF + SM_BIAS_BYTES;
constexpr int N_BARS = 2 * NSLOT + 2 * NSTAGE + 4;
constexpr int SM_TMEM_OFF = SM_BAR_OFF + N_BARS * 8;
constexpr int SM_TOTAL = SM_TMEM_OFF + 16;
static_assert(SM_TOTAL <= 232448, "shared memory budget (227 KiB per CTA)");
static_assert((NSLOT & (NSLOT - 1)) == 0, "NSLOT must be a power of two");

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
  long long pt;  // global point index (clamped to a valid one)
  bool valid;    // point index < n_points
  int s;         // tangent mode: stream of the row this thread encodes (0 primal, 1..3 d/dx_{s-1}); plain mode: 0
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

// ------------------------------------------------------------------------------------------------ pipeline trace
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
  int quad;    // TMEM lane quadrant this warp may access (warp index % 4)
  int part;    // which PCOLS columns of every 64-wide chunk this warp owns
  int lane;
  int row;     // tile row this thread owns for row-wise work (plain: 32 quad + lane; tangent: 32 quad + 8 (lane&3) + lane/4)
  uint32_t ac;  // A-chunk counter (ring position), identical in all epilogue threads and the MMA warp
  uint32_t g;   // global MMA-layer counter (accumulator buffer = g & 1)
  long long* trace;
  bool tr;  // this thread records trace events
  unsigned tcount;
};

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
// hand a finished A-operand chunk to the MMA warp: make the generic-proxy stores visible to the async proxy, then one
// elected arrive per warp (512 same-word arrivals serialise; 16 do not)
__device__ __forceinline__ void publish_chunk(EpiCtx& c, uint32_t slot, int code) {
  fence_proxy_async_smem();
  __syncwarp();
  if (c.lane == 0) mbar_arrive(&c.bars.a_full[slot]);
  if (c.tr) trace_ev(c.trace, 5000 + code, 1, &c.tcount);  // EPI: chunk written
}

// =================================================================================================================
// plain mode (one row = one point: sdf-query chain, colour chains): thread = TMEM lane, PCOLS columns per chunk
// =================================================================================================================

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
// smem-staged bias table), activate.
template <int ACT>
__device__ __forceinline__ void load_act(EpiCtx& c, int buf, int blk, int bias_row, float (&v)[PCOLS]) {
  const int col0 = 64 * blk + PCOLS * c.part;
  const uint32_t taddr = c.tmem_base + (static_cast<uint32_t>(32 * c.quad) << 16) + buf * HID + col0;
  tmem_ld<PCOLS>(taddr, v);
  const float* bias = reinterpret_cast<const float*>(c.smem + SM_BIAS_OFF) + bias_row * HID + col0;
  float bj[PCOLS];
#pragma unroll
  for (int q = 0; q < PCOLS / 4; ++q) {
    const float4 bb = *reinterpret_cast<const float4*>(bias + 4 * q);
    bj[4 * q] = bb.x; bj[4 * q + 1] = bb.y; bj[4 * q + 2] = bb.z; bj[4 * q + 3] = bb.w;
  }
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < PCOLS; ++i) {
    float h, dh;
    activate<ACT>(v[i] + bj[i], h, dh);
    v[i] = h;
  }
}
__device__ __forceinline__ void load_act_dyn(EpiCtx& c, int act, int buf, int blk, int bias, float (&v)[PCOLS]) {
  if (act == ACT_RELU) load_act<ACT_RELU>(c, buf, blk, bias, v);
  else if (act == ACT_SOFTPLUS100) load_act<ACT_SOFTPLUS100>(c, buf, blk, bias, v);
  else load_act<ACT_NONE>(c, buf, blk, bias, v);
}

// raw accumulator columns (no bias / activation): reverse chains
__device__ __forceinline__ void load_raw(EpiCtx& c, int buf, int blk, float (&v)[PCOLS]) {
  const int col0 = 64 * blk + PCOLS * c.part;
  tmem_ld<PCOLS>(c.tmem_base + (static_cast<uint32_t>(32 * c.quad) << 16) + buf * HID + col0, v);
  tmem_ld_wait();
}

// Activation backward for the reverse (training) chains.  u = adjoint of the post-activation values of a forward
// layer for this row; the forward stash holds those post-activation values h (primal rows) / hdot_j (tangent rows).
//   relu     : zbar = [h_primal > 0] * u                                (all rows; relu'' = 0)
//   softplus : sigma = 1 - exp(-100 h_primal)    (h = softplus(z)  =>  sigma(100 z) = 1 - exp(-100 h))
//              tangent rows: zdotbar_j = sigma * u_j
//              primal row  : zbar = sigma * u + 100 (1 - sigma) * sum_j hdot_j * u_j      (softplus'' = 100 s (1-s))
__device__ __forceinline__ void bwd_gate_plain(const uint16_t* shi, const uint16_t* slo, int act, float (&u)[PCOLS]) {
  float h[PCOLS];
  load_planes(shi, slo, h);
  if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < PCOLS; ++i) u[i] = h[i] > 0.f ? u[i] : 0.f;
  } else {
#pragma unroll
    for (int i = 0; i < PCOLS; ++i) u[i] = (1.f - __expf(-100.f * h[i])) * u[i];
  }
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

// Consume the whole accumulator of global layer g_layer through a NOUT-wide fp32 output layer (no MMA):
// out = W_out . act(D + bias) summed over the column parts.  Returns the total in .x/.y/.z.
template <int NOUT>
__device__ __forceinline__ float4 tail_dot(EpiCtx& c, uint32_t g_layer, int act, int bias, const float* w_out,
                                           uint16_t* st_hi = nullptr, uint16_t* st_lo = nullptr) {
  wait_d_full(c, g_layer);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int blk = 0; blk < 4; ++blk) {
    float v[PCOLS];
    load_act_dyn(c, act, g_layer & 1, blk, bias, v);
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

template <int CHAIN, bool BWD>
__device__ __forceinline__ void epilogue_plain(const ChainProg& prog, const ChainIO& io, EpiCtx& c,
                                               long long n_tiles) {
  const Bars& bars = c.bars;
  uint8_t* smem = c.smem;
  int* err = c.err;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---------------------------------------------------------- row state
    RowState rs;
    float adj[4] = {0.f, 0.f, 0.f, 0.f};  // reverse chains: (o.x, o.y, o.z, r) adjoint of this row's outputs
    {
      const long long p = tile * TILE_ROWS + c.row;
      rs.valid = p < io.n_points;
      rs.pt = rs.valid ? p : io.n_points - 1;
      rs.s = 0;
      rs.t = 0.f;
      rs.g[0] = rs.g[1] = rs.g[2] = 0.f;
      rs.dc[0] = rs.dc[1] = rs.dc[2] = 0.f;
      if constexpr (BWD) {
        rs.x[0] = rs.x[1] = rs.x[2] = 0.f;
        rs.xc[0] = rs.xc[1] = rs.xc[2] = 0.f;
        if (rs.valid) {  // padding rows carry zero adjoints so they add nothing to the weight gradients
          const float4 a = __ldg(reinterpret_cast<const float4*>(io.adj) + rs.pt);
          adj[0] = a.x; adj[1] = a.y; adj[2] = a.z; adj[3] = a.w;
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
      const int last_prev = L.last_prev;
      bool prev_waited = false;

      if (L.pre_op == PRE_DEFORM_TAIL) {
        // deform output layer (3 x 256, fp32 FFMA) -> x_c = x + delta
        float4 r = tail_dot<3>(c, c.g - 1, act_prev, bias_prev, prog.deform_out_w);
        rs.xc[0] = rs.x[0] + r.x + __ldg(prog.deform_out_b + 0);
        rs.xc[1] = rs.x[1] + r.y + __ldg(prog.deform_out_b + 1);
        rs.xc[2] = rs.x[2] + r.z + __ldg(prog.deform_out_b + 2);
        if (c.part == 0 && rs.valid && io.out_xc) {
#pragma unroll
          for (int i = 0; i < 3; ++i) io.out_xc[rs.pt * 3 + i] = rs.xc[i];
        }
        prev_waited = true;  // (already consumed and released)
      }

      for (int ck = 0; ck < L.n_chunks; ++ck, ++c.ac) {
        const uint32_t slot = c.ac & (NSLOT - 1);
        // Every layer starts after the previous layer's accumulator is complete, i.e. after every earlier MMA has
        // read its A slot: the first NSLOT chunks of a layer never have to wait for a free slot.
        if (ck >= NSLOT) mbar_wait(&bars.a_empty[slot], ((c.ac / NSLOT) & 1) ^ 1, err, 400);
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
              for (int i = 0; i < PCOLS; ++i) v[i] = fmaf(adj[3], __ldg(prog.sdf_out_w + col0 + i), v[i]);
            }
            if (ck == last_prev) release_d(c, c.g - 1);
          } else {
#pragma unroll
            for (int i = 0; i < PCOLS; ++i)
              v[i] = adj[0] * __ldg(prog.outer3_w + col0 + i) + adj[1] * __ldg(prog.outer3_w + HID + col0 + i) +
                     adj[2] * __ldg(prog.outer3_w + 2 * HID + col0 + i);
          }
          const size_t so = (static_cast<size_t>(L.stash_slot) * io.stash_rows + row_global) * HID + col0;
          bwd_gate_plain(io.stash_hi + so, io.stash_lo + so, L.bwd_act, v);
          const size_t zo = (static_cast<size_t>(L.zbar_slot) * io.stash_rows + row_global) * HID + col0;
          dump_hi = io.zbar_hi + zo;
          dump_lo = io.zbar_lo + zo;
        } else if (src == SRC_PREV) {
          if (!prev_waited) {
            wait_d_full(c, c.g - 1);
            prev_waited = true;
          }
          load_act_dyn(c, act_prev, (c.g - 1) & 1, L.arg[ck], bias_prev, v);
          if (L.side_dot) dot_accum<1>(v, prog.sdf_out_w, col0, sdf_acc);
          if (ck == last_prev) release_d(c, c.g - 1);
          if (io.stash_hi) {  // training: keep this layer's input for the reverse pass / weight gradients
            const size_t so = (static_cast<size_t>(l) * io.stash_rows + row_global) * HID + col0;
            dump_hi = io.stash_hi + so;
            dump_lo = io.stash_lo + so;
          }
        } else if (src == SRC_ENC_DEFORM) {
          encode_dispatch<SRC_ENC_DEFORM, false>(c, v, rs.x, rs);
        } else if (src == SRC_ENC_SDF) {
          encode_dispatch<SRC_ENC_SDF, false>(c, v, rs.xc, rs);
        } else if (src == SRC_COLOR_A) {
          encode_dispatch<SRC_COLOR_A, false>(c, v, rs.xc, rs);
        } else if (src == SRC_COLOR_B) {
          if (PCOLS * c.part < 32) encode_dispatch<SRC_COLOR_B, false>(c, v, rs.xc, rs);  // 32-wide chunk
          else active = false;
        } else {  // SRC_FEAT
          const float4* f4 = reinterpret_cast<const float4*>(io.feat + rs.pt * HID + 64 * L.arg[ck] + PCOLS * c.part);
#pragma unroll
          for (int q = 0; q < PCOLS / 4; ++q) {
            float4 f = __ldg(f4 + q);
            v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
          }
        }
        if (c.tr) trace_ev(io.trace, 7000 + l * 16 + ck, 1, &c.tcount);  // EPI: values ready (tmem + math done)
        if (active) emit_part((io.debug_flags & 2) ? nullptr : slot_base, c.row, c.part, v, dump_hi, dump_lo);
        publish_chunk(c, slot, l * 16 + ck);
      }

      if (L.side_dot) {
        // sdf row of the SDF output layer
        float4 r = cross_part_sum(c, make_float4(sdf_acc[0], 0.f, 0.f, 0.f));
        if (c.part == 0 && rs.valid && io.out_sdf) io.out_sdf[rs.pt] = r.x + __ldg(prog.sdf_out_b);
      }
    }

    // ---------------------------------------------------------- post op: consume the last accumulator
    const int last = prog.n_layers - 1;
    const int act_last = prog.layer[last].act;
    if (prog.post_op == POST_SDF_TAIL) {
      float4 r = tail_dot<1>(c, c.g - 1, act_last, last, prog.sdf_out_w);
      if (c.part == 0 && rs.valid && io.out_sdf) io.out_sdf[rs.pt] = r.x + __ldg(prog.sdf_out_b);
    } else if (BWD && prog.post_op == POST_BWD_DUMP) {
      wait_d_full(c, c.g - 1);
#pragma unroll 1
      for (int blk = 0; blk < 4; ++blk) {
        float v[PCOLS];
        const int col0 = 64 * blk + PCOLS * c.part;
        load_raw(c, (c.g - 1) & 1, blk, v);
        const size_t so = (static_cast<size_t>(prog.post_stash_slot) * io.stash_rows + row_global) * HID + col0;
        bwd_gate_plain(io.stash_hi + so, io.stash_lo + so, prog.post_bwd_act, v);
        const size_t zo = (static_cast<size_t>(prog.post_zbar_slot) * io.stash_rows + row_global) * HID + col0;
        emit_part(nullptr, 0, 0, v, io.zbar_hi + zo, io.zbar_lo + zo);
      }
      release_d(c, c.g - 1);
    } else if (prog.post_op == POST_COLOR_TAIL) {
      uint16_t* sh = io.stash_hi ? io.stash_hi + (static_cast<size_t>(prog.n_layers) * io.stash_rows + row_global) * HID : nullptr;
      uint16_t* sl = io.stash_hi ? io.stash_lo + (static_cast<size_t>(prog.n_layers) * io.stash_rows + row_global) * HID : nullptr;
      float4 r = tail_dot<3>(c, c.g - 1, act_last, last, prog.color_out_w, sh, sl);
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
        load_act<ACT_NONE>(c, (c.g - 1) & 1, blk, MAXL, v);  // row MAXL = feature-layer bias
        if (rs.valid) {
          float4* o4 = reinterpret_cast<float4*>(io.out_feat + rs.pt * HID + 64 * blk + PCOLS * c.part);
#pragma unroll
          for (int q = 0; q < PCOLS / 4; ++q) o4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
      release_d(c, c.g - 1);
    }
  }
}

// =================================================================================================================
// tangent mode: fragment form.  A thread (quadrant Q, part, lane = 4 p + q) holds, for point p of the quadrant, the
// four streams s = 0..3 (tile rows 32Q + 8s + p) at the four accumulator columns
//     col(c) = 64 blk + 16 part + 8 (c >> 1) + 2 q + (c & 1),   c = 0..3.
// =================================================================================================================
struct Frag {
  float f[4][4];  // [stream][c]
};

__device__ __forceinline__ void load_frag(const EpiCtx& c, int buf, int blk, Frag& F) {
  const uint32_t ta = c.tmem_base + (static_cast<uint32_t>(32 * c.quad) << 16) + buf * HID + 64 * blk + PCOLS * c.part;
  tmem_ld_16x256b_x2(ta, F.f[0][0], F.f[0][1], F.f[1][0], F.f[1][1], F.f[0][2], F.f[0][3], F.f[1][2], F.f[1][3]);
  tmem_ld_16x256b_x2(ta + (16u << 16), F.f[2][0], F.f[2][1], F.f[3][0], F.f[3][1], F.f[2][2], F.f[2][3], F.f[3][2],
                     F.f[3][3]);
  tmem_ld_wait();
}

// bias + activation of the primal stream, chain rule on the three tangent streams
template <int ACT>
__device__ __forceinline__ void act_frag(const EpiCtx& c, int blk, int bias_row, Frag& F) {
  const float* bias = reinterpret_cast<const float*>(c.smem + SM_BIAS_OFF) + bias_row * HID + 64 * blk + PCOLS * c.part +
                      2 * (c.lane & 3);
  const float2 b0 = *reinterpret_cast<const float2*>(bias);
  const float2 b1 = *reinterpret_cast<const float2*>(bias + 8);
  const float b[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float z = F.f[0][i] + b[i];
    if constexpr (ACT == ACT_RELU) {
      const bool m = z > 0.f;
      F.f[0][i] = m ? z : 0.f;
      F.f[1][i] = m ? F.f[1][i] : 0.f;
      F.f[2][i] = m ? F.f[2][i] : 0.f;
      F.f[3][i] = m ? F.f[3][i] : 0.f;
    } else if constexpr (ACT == ACT_SOFTPLUS100) {
      float h, dh;
      activate<ACT_SOFTPLUS100>(z, h, dh);
      F.f[0][i] = h;
      F.f[1][i] *= dh;
      F.f[2][i] *= dh;
      F.f[3][i] *= dh;
    } else {
      F.f[0][i] = z;
    }
  }
}
__device__ __forceinline__ void act_frag_dyn(const EpiCtx& c, int act, int blk, int bias_row, Frag& F) {
  if (act == ACT_RELU) act_frag<ACT_RELU>(c, blk, bias_row, F);
  else if (act == ACT_SOFTPLUS100) act_frag<ACT_SOFTPLUS100>(c, blk, bias_row, F);
  else act_frag<ACT_NONE>(c, blk, bias_row, F);
}

// split to fp16 hi/lo and store: `slot` = this thread's base inside an A ring slot (k-group 2 part, row 32Q + p,
// byte 4q) or null; dhi/dlo = global plane pointers at (row of stream 0, col(0)) or null (training stash / zbar).
__device__ __forceinline__ void emit_frag(const Frag& F, uint8_t* slot, uint16_t* dhi, uint16_t* dlo) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t hi, lo;
      split2(F.f[s][2 * j], F.f[s][2 * j + 1], hi, lo);
      if (slot) {
        *reinterpret_cast<uint32_t*>(slot + j * A_LBO + s * 128) = hi;
        *reinterpret_cast<uint32_t*>(slot + SLOT_HALF_BYTES + j * A_LBO + s * 128) = lo;
      }
      if (dhi) {
        *reinterpret_cast<uint32_t*>(dhi + s * 8 * HID + 8 * j) = hi;
        *reinterpret_cast<uint32_t*>(dlo + s * 8 * HID + 8 * j) = lo;
      }
    }
  }
}

// fragment of fp16 hi/lo planes (value = hi + lo); pointers at (row of stream 0, col(0))
__device__ __forceinline__ void load_planes_frag(const uint16_t* phi, const uint16_t* plo, Frag& H) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(phi + s * 8 * HID + 8 * j));
      const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(plo + s * 8 * HID + 8 * j));
      const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&a));
      const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&b));
      H.f[s][2 * j] = fa.x + fb.x;
      H.f[s][2 * j + 1] = fa.y + fb.y;
    }
  }
}

// activation backward on a fragment (see bwd_gate_plain for the formulas); everything is thread-local
__device__ __forceinline__ void bwd_gate_frag(const uint16_t* shi, const uint16_t* slo, int act, Frag& U) {
  Frag H;
  load_planes_frag(shi, slo, H);
  if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool m = H.f[0][i] > 0.f;
#pragma unroll
      for (int s = 0; s < 4; ++s) U.f[s][i] = m ? U.f[s][i] : 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float e = __expf(-100.f * H.f[0][i]);  // 1 - sigma
      const float sg = 1.f - e;
      const float ct = H.f[1][i] * U.f[1][i] + H.f[2][i] * U.f[2][i] + H.f[3][i] * U.f[3][i];
      U.f[0][i] = fmaf(100.f * e, ct, sg * U.f[0][i]);
      U.f[1][i] *= sg;
      U.f[2][i] *= sg;
      U.f[3][i] *= sg;
    }
  }
}

// acc[s][o] += sum_c F[s][c] * w[o][col(c)]
template <int NOUT>
__device__ __forceinline__ void dot_frag(const Frag& F, const float* __restrict__ w, int colq, float (&acc)[4][NOUT]) {
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    const float2 w0 = __ldg(reinterpret_cast<const float2*>(w + o * HID + colq));
    const float2 w1 = __ldg(reinterpret_cast<const float2*>(w + o * HID + colq + 8));
#pragma unroll
    for (int s = 0; s < 4; ++s)
      acc[s][o] += fmaf(F.f[s][0], w0.x, F.f[s][1] * w0.y) + fmaf(F.f[s][2], w1.x, F.f[s][3] * w1.y);
  }
}

// total of a per-point vector over the 4 column lanes of the quad and the NPART column-part warps; every thread of
// the point gets the result.  NV = 4 or 12.
template <int NV>
__device__ __forceinline__ void point_sum(EpiCtx& c, float (&v)[NV]) {
  static_assert(NV % 4 == 0 && NPART * TILE_PTS_T * NV * 4 <= SM_XCH_BYTES, "exchange buffer");
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1);
    v[i] += __shfl_xor_sync(0xffffffffu, v[i], 2);
  }
  float4* xch = reinterpret_cast<float4*>(c.smem + SM_XCH_OFF);
  const int ptl = 8 * c.quad + (c.lane >> 2);
  if ((c.lane & 3) == 0) {
#pragma unroll
    for (int i = 0; i < NV / 4; ++i)
      xch[(c.part * TILE_PTS_T + ptl) * (NV / 4) + i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  named_bar_sync(1, N_EPI_THREADS);
#pragma unroll
  for (int i = 0; i < NV / 4; ++i) {
    float4 t = xch[ptl * (NV / 4) + i];
#pragma unroll
    for (int q = 1; q < NPART; ++q) {
      const float4 b = xch[(q * TILE_PTS_T + ptl) * (NV / 4) + i];
      t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
    }
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
  named_bar_sync(2, N_EPI_THREADS);  // everyone has read before the buffer is written again
}

// Consume the whole accumulator of layer g_layer through a NOUT-wide fp32 output layer: out[s][o] for the 4 streams.
// st_hi / st_lo: training stash planes at (row of stream 0, column 2q of part 0... ) i.e. + 64 blk + 16 part added here.
template <int NOUT>
__device__ __forceinline__ void tail_frag(EpiCtx& c, uint32_t g_layer, int act, int bias, const float* w_out,
                                          float (&out)[4 * NOUT], uint16_t* st_hi, uint16_t* st_lo) {
  wait_d_full(c, g_layer);
  float acc[4][NOUT];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int o = 0; o < NOUT; ++o) acc[s][o] = 0.f;
  const int colq = PCOLS * c.part + 2 * (c.lane & 3);
#pragma unroll 1
  for (int blk = 0; blk < 4; ++blk) {
    Frag F;
    load_frag(c, g_layer & 1, blk, F);
    act_frag_dyn(c, act, blk, bias, F);
    if (st_hi) emit_frag(F, nullptr, st_hi + 64 * blk + colq, st_lo + 64 * blk + colq);
    dot_frag<NOUT>(F, w_out, 64 * blk + colq, acc);
  }
  release_d(c, g_layer);
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int o = 0; o < NOUT; ++o) out[s * NOUT + o] = acc[s][o];
  point_sum<4 * NOUT>(c, out);
}

template <bool BWD>
__device__ __forceinline__ void epilogue_tangent(const ChainProg& prog, const ChainIO& io, EpiCtx& c,
                                                 long long n_tiles) {
  const Bars& bars = c.bars;
  uint8_t* smem = c.smem;
  int* err = c.err;
  const int q = c.lane & 3, p = c.lane >> 2;
  const int colq = PCOLS * c.part + 2 * q;                                   // col(0) inside a 64-column block
  const int frag_off = 2 * c.part * A_LBO + (32 * c.quad + p) * 16 + 4 * q;  // see emit_frag
  const bool writer = (c.part == 0);

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---------------------------------------------------------- point state (same for the 4 lanes of a quad)
    RowState rs;
    {
      const long long pp = tile * TILE_PTS_T + 8 * c.quad + p;
      rs.valid = pp < io.n_points;
      rs.pt = rs.valid ? pp : io.n_points - 1;
      rs.s = q;  // row-wise work (encodings): this thread is the row of stream q
      rs.t = 0.f;
      rs.g[0] = rs.g[1] = rs.g[2] = 0.f;
      rs.dc[0] = rs.dc[1] = rs.dc[2] = 0.f;
      rs.x[0] = rs.x[1] = rs.x[2] = 0.f;
      if constexpr (!BWD) {
        const float* xp = io.x + rs.pt * 3;
#pragma unroll
        for (int i = 0; i < 3; ++i) rs.x[i] = __ldg(xp + i);
        rs.t = __ldg(io.t + (rs.pt / io.t_div) * io.t_stride);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) rs.xc[i] = rs.x[i];
    }
    float sdf_acc[4][1] = {{0.f}, {0.f}, {0.f}, {0.f}};
    // global plane row of stream 0 of this point (stream s: + 8 s rows)
    const size_t row0 = static_cast<size_t>(tile) * TILE_ROWS + 32 * c.quad + p;

    for (int l = 0; l < prog.n_layers; ++l, ++c.g) {
      const LayerProg& L = prog.layer[l];
      const int bias_prev = l > 0 ? l - 1 : 0;  // row of the smem bias table
      const int act_prev = l > 0 ? prog.layer[l - 1].act : ACT_NONE;
      const int last_prev = L.last_prev;
      bool prev_waited = false;

      if (!BWD && L.pre_op == PRE_DEFORM_TAIL) {
        // deform output layer (3 x 256, fp32 FFMA) -> x_c = x + delta ; tangent streams give dDelta/dx_{s-1}
        // training: the output layer's input goes to stash slot l (this layer has no SRC_PREV chunk of its own)
        uint16_t* sh = io.stash_hi ? io.stash_hi + (static_cast<size_t>(l) * io.stash_rows + row0) * HID : nullptr;
        uint16_t* sl = io.stash_hi ? io.stash_lo + (static_cast<size_t>(l) * io.stash_rows + row0) * HID : nullptr;
        float o[12];
        tail_frag<3>(c, c.g - 1, act_prev, bias_prev, prog.deform_out_w, o, sh, sl);
#pragma unroll
        for (int i = 0; i < 3; ++i) rs.xc[i] = rs.x[i] + o[i] + __ldg(prog.deform_out_b + i);
        if (writer && rs.valid) {
          if (q == 0) {
            if (io.out_xc) {
#pragma unroll
              for (int i = 0; i < 3; ++i) io.out_xc[rs.pt * 3 + i] = rs.xc[i];
            }
          } else if (io.out_jac) {
            // column j = q-1 of J = I + dDelta/dx ; J stored [i][j] row-major
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const float dl = q == 1 ? o[3 + i] : (q == 2 ? o[6 + i] : o[9 + i]);
              io.out_jac[rs.pt * 9 + 3 * i + (q - 1)] = dl + ((i == q - 1) ? 1.f : 0.f);
            }
          }
        }
        prev_waited = true;  // (already consumed and released)
      }

      for (int ck = 0; ck < L.n_chunks; ++ck, ++c.ac) {
        const uint32_t slot = c.ac & (NSLOT - 1);
        // Every layer starts after the previous layer's accumulator is complete, i.e. after every earlier MMA has
        // read its A slot: the first NSLOT chunks of a layer never have to wait for a free slot.
        if (ck >= NSLOT) mbar_wait(&bars.a_empty[slot], ((c.ac / NSLOT) & 1) ^ 1, err, 400);
        uint8_t* slot_base = smem + SM_A_OFF + slot * SLOT_BYTES;
        const bool store_a = !(io.debug_flags & 2);
        const int src = L.src[ck];
        const int blk = L.arg[ck];
        if (src == SRC_PREV || src == SRC_BWD_PREV || src == SRC_BWD_OUTER3 || src == SRC_ADJ_FEAT) {
          Frag F;
          uint16_t* dump_hi = nullptr;
          uint16_t* dump_lo = nullptr;
          if (!BWD) {
            if (!prev_waited) {
              wait_d_full(c, c.g - 1);
              prev_waited = true;
            }
            load_frag(c, (c.g - 1) & 1, blk, F);
            act_frag_dyn(c, act_prev, blk, bias_prev, F);
            if (L.side_dot) dot_frag<1>(F, prog.sdf_out_w, 64 * blk + colq, sdf_acc);
            if (ck == last_prev) release_d(c, c.g - 1);
            if (io.stash_hi) {  // training: keep this layer's input for the reverse pass / weight gradients
              const size_t so = (static_cast<size_t>(l) * io.stash_rows + row0) * HID + 64 * blk + colq;
              dump_hi = io.stash_hi + so;
              dump_lo = io.stash_lo + so;
            }
          } else if (src == SRC_ADJ_FEAT) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
              for (int i = 0; i < 4; ++i) F.f[s][i] = 0.f;
            if (rs.valid) {
              const float* fp = io.adj_feat + rs.pt * HID + 64 * blk + colq;
              const float2 f0 = __ldg(reinterpret_cast<const float2*>(fp));
              const float2 f1 = __ldg(reinterpret_cast<const float2*>(fp + 8));
              F.f[0][0] = f0.x; F.f[0][1] = f0.y; F.f[0][2] = f1.x; F.f[0][3] = f1.y;
            }
          } else {
            // adjoints (o.x, o.y, o.z, r) of the 4 streams' 3-wide / sdf-row outputs; padding points carry zeros
            float4 a[4];
#pragma unroll
            for (int s = 0; s < 4; ++s)
              a[s] = rs.valid ? __ldg(reinterpret_cast<const float4*>(io.adj) + rs.pt * 4 + s)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
            if (src == SRC_BWD_PREV) {
              if (!prev_waited) {
                wait_d_full(c, c.g - 1);
                prev_waited = true;
              }
              load_frag(c, (c.g - 1) & 1, blk, F);
              if (L.rank1) {
                const float2 w0 = __ldg(reinterpret_cast<const float2*>(prog.sdf_out_w + 64 * blk + colq));
                const float2 w1 = __ldg(reinterpret_cast<const float2*>(prog.sdf_out_w + 64 * blk + colq + 8));
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                  F.f[s][0] = fmaf(a[s].w, w0.x, F.f[s][0]);
                  F.f[s][1] = fmaf(a[s].w, w0.y, F.f[s][1]);
                  F.f[s][2] = fmaf(a[s].w, w1.x, F.f[s][2]);
                  F.f[s][3] = fmaf(a[s].w, w1.y, F.f[s][3]);
                }
              }
              if (ck == last_prev) release_d(c, c.g - 1);
            } else {  // SRC_BWD_OUTER3
              float w[3][4];
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                const float2 w0 = __ldg(reinterpret_cast<const float2*>(prog.outer3_w + i * HID + 64 * blk + colq));
                const float2 w1 = __ldg(reinterpret_cast<const float2*>(prog.outer3_w + i * HID + 64 * blk + colq + 8));
                w[i][0] = w0.x; w[i][1] = w0.y; w[i][2] = w1.x; w[i][3] = w1.y;
              }
#pragma unroll
              for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int i = 0; i < 4; ++i) F.f[s][i] = a[s].x * w[0][i] + a[s].y * w[1][i] + a[s].z * w[2][i];
            }
            const size_t so = (static_cast<size_t>(L.stash_slot) * io.stash_rows + row0) * HID + 64 * blk + colq;
            bwd_gate_frag(io.stash_hi + so, io.stash_lo + so, L.bwd_act, F);
            const size_t zo = (static_cast<size_t>(L.zbar_slot) * io.stash_rows + row0) * HID + 64 * blk + colq;
            dump_hi = io.zbar_hi + zo;
            dump_lo = io.zbar_lo + zo;
          }
          if (c.tr) trace_ev(io.trace, 7000 + l * 16 + ck, 1, &c.tcount);  // EPI: values ready (tmem + math done)
          emit_frag(F, store_a ? slot_base + frag_off : nullptr, dump_hi, dump_lo);
        } else {
          // positional encodings: row-wise, this thread is row 32Q + 8q + p (stream q of its point)
          float v[PCOLS];
          if (src == SRC_ENC_DEFORM) encode_dispatch<SRC_ENC_DEFORM, true>(c, v, rs.x, rs);
          else encode_dispatch<SRC_ENC_SDF, true>(c, v, rs.xc, rs);
          if (c.tr) trace_ev(io.trace, 7000 + l * 16 + ck, 1, &c.tcount);
          emit_part(store_a ? slot_base : nullptr, c.row, c.part, v, nullptr, nullptr);
        }
        publish_chunk(c, slot, l * 16 + ck);
      }

      if (!BWD && L.side_dot) {
        // sdf row of the SDF output layer: primal stream -> sdf, tangent streams -> g_c
        float o[4] = {sdf_acc[0][0], sdf_acc[1][0], sdf_acc[2][0], sdf_acc[3][0]};
        point_sum<4>(c, o);
        if (writer && rs.valid) {
          if (q == 0) {
            if (io.out_sdf) io.out_sdf[rs.pt] = o[0] + __ldg(prog.sdf_out_b);
          } else if (io.out_gc) {
            io.out_gc[rs.pt * 3 + (q - 1)] = q == 1 ? o[1] : (q == 2 ? o[2] : o[3]);
          }
        }
      }
    }

    // ---------------------------------------------------------- post op: consume the last accumulator
    const int last = prog.n_layers - 1;
    if (!BWD && prog.post_op == POST_SDF_TAIL) {
      float o[4];
      tail_frag<1>(c, c.g - 1, prog.layer[last].act, last, prog.sdf_out_w, o, nullptr, nullptr);
      if (writer && rs.valid) {
        if (q == 0) {
          if (io.out_sdf) io.out_sdf[rs.pt] = o[0] + __ldg(prog.sdf_out_b);
        } else if (io.out_gc) {
          io.out_gc[rs.pt * 3 + (q - 1)] = q == 1 ? o[1] : (q == 2 ? o[2] : o[3]);
        }
      }
    } else if (BWD && prog.post_op == POST_BWD_DUMP) {
      wait_d_full(c, c.g - 1);
#pragma unroll 1
      for (int blk = 0; blk < 4; ++blk) {
        Frag F;
        load_frag(c, (c.g - 1) & 1, blk, F);
        const size_t so = (static_cast<size_t>(prog.post_stash_slot) * io.stash_rows + row0) * HID + 64 * blk + colq;
        bwd_gate_frag(io.stash_hi + so, io.stash_lo + so, prog.post_bwd_act, F);
        const size_t zo = (static_cast<size_t>(prog.post_zbar_slot) * io.stash_rows + row0) * HID + 64 * blk + colq;
        emit_frag(F, nullptr, io.zbar_hi + zo, io.zbar_lo + zo);
      }
      release_d(c, c.g - 1);
    } else if (!BWD && prog.post_op == POST_FEAT_OUT) {
      wait_d_full(c, c.g - 1);
      const float* fb = reinterpret_cast<const float*>(smem + SM_BIAS_OFF) + MAXL * HID;  // feature-layer bias
#pragma unroll 1
      for (int blk = 0; blk < 4; ++blk) {
        // feat = D + bias (no activation); only the primal stream is needed: lanes L = 32Q + p of this quadrant
        Frag F;
        load_frag(c, (c.g - 1) & 1, blk, F);
        if (rs.valid) {
          const float* b = fb + 64 * blk + colq;
          float* o = io.out_feat + rs.pt * HID + 64 * blk + colq;
          *reinterpret_cast<float2*>(o) = make_float2(F.f[0][0] + b[0], F.f[0][1] + b[1]);
          *reinterpret_cast<float2*>(o + 8) = make_float2(F.f[0][2] + b[8], F.f[0][3] + b[9]);
        }
      }
      release_d(c, c.g - 1);
    }
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
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

  const int rows_per_tile = TANGENT ? TILE_PTS_T : TILE_ROWS;  // points per tile
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
    c.quad = warp & 3;
    c.part = (warp - 2) >> 2;
    c.row = TANGENT ? 32 * c.quad + 8 * (lane & 3) + (lane >> 2) : 32 * c.quad + lane;
    c.ac = 0;
    c.g = 0;
    c.trace = io.trace;
    c.tr = (io.trace != nullptr && warp == 2 && lane == 0 && blockIdx.x == 0);
    c.tcount = 0;
    if constexpr (TANGENT) epilogue_tangent<BWD>(prog, io, c, n_tiles);
    else epilogue_plain<CHAIN, BWD>(prog, io, c, n_tiles);
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
  const int pts_per_tile = TANGENT ? TILE_PTS_T : TILE_ROWS;
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
