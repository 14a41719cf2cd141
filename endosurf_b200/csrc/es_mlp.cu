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
//
// Thread layout: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..17 = epilogue.  The register file is split per SM
// sub-partition (16384 registers; two sub-partitions hold 5 warps) => 96 registers/thread.  (setmaxnreg cannot be
// combined with the out-of-line helpers below: ptxas refuses to allocate ABI calls inside a re-sized region.)
constexpr int NPART = 4;
constexpr int PCOLS = CHUNK_K / NPART;  // 16
constexpr int N_EPI_WARPS = 4 * NPART;
constexpr int N_EPI_THREADS = N_EPI_WARPS * 32;
constexpr int EPI_WARP0 = 2;
constexpr int N_THREADS = EPI_WARP0 * 32 + N_EPI_THREADS;
constexpr int STORE_WARP = EPI_WARP0 + N_EPI_WARPS;  // training launches only: plane-dump warp (bulk smem -> global)
constexpr int N_THREADS_DUMP = N_THREADS + 32;
constexpr int TILE_PTS_T = TILE_ROWS / 4;  // points per tile in tangent mode
constexpr int CHUNK_PLANE_BYTES = SLOT_HALF_BYTES;  // one dumped chunk half: [8 k-groups][128 rows][8 fp16] = 16 KiB

// dynamic shared memory carve-up (byte offsets from the 1024-aligned base)
constexpr int SM_A_OFF = 0;
constexpr int SM_W_OFF = SM_A_OFF + NSLOT * SLOT_BYTES;
constexpr int SM_XCH_OFF = SM_W_OFF + NSTAGE * STAGE_BYTES;
constexpr int SM_XCH_BYTES = NPART * TILE_ROWS * 4 * 4;      // plain: [part][row][4] floats; tangent: [part][pt][<=12]
constexpr int SM_BIAS_OFF = SM_XCH_OFF + SM_XCH_BYTES;       // [MAXL + 1][256] fp32: every layer's bias + feat bias
constexpr int SM_BIAS_BYTES = (MAXL + 1) * HID * 4;
constexpr int SM_BAR_OFF = SM_BIAS_OFF + SM_BIAS_BYTES;
constexpr int BAR_A_FULL = SM_BAR_OFF;                       // [NSLOT]  epilogue -> MMA   (count N_EPI_WARPS)
constexpr int BAR_A_EMPTY = BAR_A_FULL + 8 * NSLOT;          // [NSLOT]  MMA commit -> epilogue
constexpr int BAR_W_FULL = BAR_A_EMPTY + 8 * NSLOT;          // [NSTAGE] TMA -> MMA
constexpr int BAR_W_EMPTY = BAR_W_FULL + 8 * NSTAGE;         // [NSTAGE] MMA commit -> TMA
constexpr int BAR_D_FULL = BAR_W_EMPTY + 8 * NSTAGE;         // [2]      MMA commit -> epilogue
constexpr int BAR_D_EMPTY = BAR_D_FULL + 16;                 // [2]      epilogue -> MMA  (count N_EPI_WARPS)
// CTA pairs: arrivals relayed from the peer CTA into the leader's shared memory (count 1 each)
constexpr int BAR_A_FULL_PEER = BAR_D_EMPTY + 16;            // [NSLOT]
constexpr int BAR_W_FULL_PEER = BAR_A_FULL_PEER + 8 * NSLOT; // [NSTAGE]
constexpr int BAR_D_EMPTY_PEER = BAR_W_FULL_PEER + 8 * NSTAGE;  // [2]
constexpr int SM_TMEM_OFF = BAR_D_EMPTY_PEER + 16;
constexpr int SM_TOTAL = SM_TMEM_OFF + 16;
static_assert(SM_TOTAL <= 232448, "shared memory budget (227 KiB per CTA)");

#ifdef ES_ABLATE  // perf experiments (ES_DEBUG_FLAGS): 1 = no weight copies, 2 = no A stores, 4 = no MMAs,
                  // 8 = no L2 gate prefetch (reverse chains), 16 = no plane-record stores (training launches),
                  // 32 = plane records of every tile land on the CTA's first tile (L2-resident stores)
#define ES_FLAG(io, bit) (((io).debug_flags & (bit)) != 0)
#else
#define ES_FLAG(io, bit) false
#endif

// ------------------------------------------------------------------------------------------------ tile walk
// Which tiles a CTA processes.  Single CTAs stride over the tiles; the two CTAs of a pair take tiles 2p and 2p + 1 of the
// pairs p their cluster strides over, and always run the same number of iterations (the odd one out at the end is a
// ghost tile: every row invalid, nothing read past the records, nothing written).
struct TileWalk {
  long long tile0, step, n_iter, n_tiles;
};
template <bool PAIR>
__device__ __forceinline__ TileWalk make_walk(long long n_tiles) {
  TileWalk w;
  w.n_tiles = n_tiles;
  if constexpr (PAIR) {
    const long long c = cluster_id_x(), C = cluster_nctaid_x();
    const long long n_pairs = (n_tiles + 1) / 2;
    w.tile0 = 2 * c + cluster_ctarank();
    w.step = 2 * C;
    w.n_iter = c < n_pairs ? (n_pairs - c + C - 1) / C : 0;
  } else {
    w.tile0 = blockIdx.x;
    w.step = gridDim.x;
    w.n_iter = static_cast<long long>(blockIdx.x) < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  }
  return w;
}

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

// ------------------------------------------------------------------------------------------------ encoder inputs
struct EncIn {
  float p[3];    // position the encoding is taken of (x for the deform net, x_c otherwise)
  float t;       // time
  float g[3];    // colour chain: canonical normal g_c
  float dc[3];   // colour chain: canonical view direction
  int s;         // tangent mode: stream of the row being encoded (0 primal, 1..3 d/dx_{s-1}); plain mode: 0
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
// Tangent rows (s>0) get the derivative of every feature wrt position component s-1.  All feature indices resolve at compile time, so
// only the sin/cos pairs this part needs are evaluated and everything lives in registers.
template <int SRC, int PART, bool TANGENT>
__device__ __forceinline__ void encode_part(float (&v)[PCOLS], const EncIn& rs) {
  float var[10];
  var[0] = rs.p[0]; var[1] = rs.p[1]; var[2] = rs.p[2];
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

// ------------------------------------------------------------------------------------------------ pipeline trace
// Debug pipeline trace, compiled in with -DES_TRACE only (tools/trace_chain.py rebuilds the library): CTA 0's MMA issuer
// and epilogue warp EPI_WARP0 lane 0 write (clock, code) pairs into separate halves of ChainIO::trace.
// layout: trace[0] = MMA count, trace[1] = EPI count, then 4000 (clock, code) pairs each.
#ifdef ES_TRACE
__device__ __forceinline__ void trace_ev(long long* trace, int code, int who, unsigned* counter) {
  if (trace != nullptr && blockIdx.x == 0) {
    const unsigned i = (*counter)++;
    if (i < 4000) {
      long long* base = trace + 2 + who * 8000;
      base[2 * i] = clock64();
      base[2 * i + 1] = code;
      trace[who] = i + 1;
    }
  }
}
#define TRACE_MMA(code) trace_ev(io.trace, (code), 0, &tcount)
#define TRACE_EPI(code) do { if (c.tr) trace_ev(c.trace, (code), 1, &c.tcount); } while (0)
#else
#define TRACE_MMA(code) do { } while (0)
#define TRACE_EPI(code) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------ epilogue context
// Everything an epilogue thread needs to address shared memory / TMEM, as 32-bit values that stay in registers.
struct Epi {
  uint32_t sm;    // shared-space address of the dynamic shared memory base
  uint32_t tmem;  // tmem_base + (32 quad << 16) + PCOLS part: this warp's lanes and columns inside a 64-column block
  int* err;
  int quad;       // TMEM lane quadrant this warp may access (warp index % 4)
  int part;       // which PCOLS columns of every 64-wide chunk this warp owns
  int lane;
  int row;        // tile row owned for row-wise work (plain: 32 quad + lane; tangent: 32 quad + 8 (lane&3) + lane/4)
  uint32_t ac;    // A-chunk counter (ring position), identical in all epilogue threads and the MMA warp
  uint32_t g;     // global MMA-layer counter (accumulator buffer = g & 1)
#ifdef ES_TRACE
  long long* trace;
  bool tr;
  unsigned tcount;
#endif
};

__device__ __forceinline__ void wait_d_full(Epi& c, uint32_t g_layer) {
  mbar_wait_sa(c.sm + BAR_D_FULL + 8 * (g_layer & 1), (g_layer >> 1) & 1, c.err, 100);
  tc_fence_after();
  TRACE_EPI(4000 + static_cast<int>(g_layer % 100));  // EPI: accumulator of layer g ready
}
__device__ __forceinline__ void release_d(const Epi& c, uint32_t g_layer) {
  tc_fence_before();
  __syncwarp();
  if (c.lane == 0) mbar_arrive_sa(c.sm + BAR_D_EMPTY + 8 * (g_layer & 1));
}
// Claim the next A ring slot.  Every layer starts after the previous layer's accumulator is complete, i.e. after every
// earlier MMA has read its A slot: the first NSLOT chunks of a layer never have to wait for a free slot.
// In training launches (DUMP) the plane-dump warp reads the slots as well and dump-only chunks are interleaved, so
// the wait is unconditional there.
template <bool DUMP>
__device__ __forceinline__ uint32_t claim_slot(const Epi& c, int ck) {
  const uint32_t slot = c.ac % NSLOT;
  if (DUMP || ck >= NSLOT) mbar_wait_sa(c.sm + BAR_A_EMPTY + 8 * slot, ((c.ac / NSLOT) & 1) ^ 1, c.err, 400);
  return slot;
}
// Hand a finished A-operand chunk to the MMA warp: make the generic-proxy stores visible to the async proxy, then one
// elected arrive per warp (512 same-word arrivals serialise; 16 do not).
__device__ __forceinline__ void publish_chunk(Epi& c, uint32_t slot, int code) {
  fence_proxy_async_smem();
  __syncwarp();
  if (c.lane == 0) mbar_arrive_sa(c.sm + BAR_A_FULL + 8 * slot);
  TRACE_EPI(5000 + code);  // EPI: chunk written
}

// ------------------------------------------------------------------------------------------------ row-form stores
// Split v[PCOLS] into fp16 hi/lo and store them as this row's 16-byte units of k-groups 2 part, 2 part + 1 of an A
// ring slot (sa = slot base + 2 part A_LBO + 16 row).
__device__ __forceinline__ void emit_row(uint32_t sa, const float (&v)[PCOLS]) {
  static_for<0, PCOLS / 8>([&](auto gc) {
    constexpr int g = decltype(gc)::value;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[8 * g + 2 * j], v[8 * g + 2 * j + 1], hi[j], lo[j]);
    sts128<g * A_LBO>(sa, hi[0], hi[1], hi[2], hi[3]);
    sts128<g * A_LBO + SLOT_HALF_BYTES>(sa, lo[0], lo[1], lo[2], lo[3]);
  });
}

// Read this row's PCOLS values of a dumped plane chunk (value = hi + lo; plo may be null).  phi / plo point at the
// chunk half + 2 part A_LBO + 16 row, i.e. the layout emit_row wrote into the ring slot.
__device__ __forceinline__ void load_planes(const uint8_t* phi, const uint8_t* plo, float (&h)[PCOLS]) {
#pragma unroll
  for (int g = 0; g < PCOLS / 8; ++g) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(phi + g * A_LBO));
    const uint4 b = plo ? __ldg(reinterpret_cast<const uint4*>(plo + g * A_LBO)) : make_uint4(0u, 0u, 0u, 0u);
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

// raw hi-plane words of load_planes (the plo == nullptr case), fetched before the accumulator wait and unpacked at the
// use: 8 live registers instead of 16 floats
__device__ __forceinline__ void load_gate_raw(const uint8_t* phi, uint4 (&g)[PCOLS / 8]) {
#pragma unroll
  for (int k = 0; k < PCOLS / 8; ++k) g[k] = __ldg(reinterpret_cast<const uint4*>(phi + k * A_LBO));
}
__device__ __forceinline__ void unpack_gate_raw(const uint4 (&g)[PCOLS / 8], float (&h)[PCOLS]) {
#pragma unroll
  for (int k = 0; k < PCOLS / 8; ++k) {
    const uint32_t w[4] = {g[k].x, g[k].y, g[k].z, g[k].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
      h[8 * k + 2 * j] = f.x;
      h[8 * k + 2 * j + 1] = f.y;
    }
  }
}

// SRC_PLANE: copy this row's 16-byte units (k-groups 2 part, 2 part + 1; hi and lo halves) of a dumped chunk into the
// ring slot.  off = 2 part A_LBO + 16 row (the same offset in the chunk and in the slot half).
__device__ __forceinline__ void copy_plane_row(uint32_t slot_sa, const uint8_t* chi, const uint8_t* clo, uint32_t off) {
#pragma unroll
  for (int g = 0; g < PCOLS / 8; ++g) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(chi + off + g * A_LBO));
    const uint4 b = clo ? __ldg(reinterpret_cast<const uint4*>(clo + off + g * A_LBO)) : make_uint4(0u, 0u, 0u, 0u);
    if (g == 0) {
      sts128<0>(slot_sa + off, a.x, a.y, a.z, a.w);
      sts128<SLOT_HALF_BYTES>(slot_sa + off, b.x, b.y, b.z, b.w);
    } else {
      sts128<A_LBO>(slot_sa + off, a.x, a.y, a.z, a.w);
      sts128<A_LBO + SLOT_HALF_BYTES>(slot_sa + off, b.x, b.y, b.z, b.w);
    }
  }
}

// SRC_PLANE with both halves present: the chunk comes straight from the record by two 16 KiB bulk copies into the ring
// slot.  One thread issues them and its arrive on the slot's full barrier carries the transaction bytes; the other
// epilogue warps only arrive.  (Replaces publish_chunk for this chunk.)
__device__ __forceinline__ void reload_chunk_bulk(const Epi& c, uint32_t slot, const uint8_t* chi, const uint8_t* clo) {
  const uint32_t bar = c.sm + BAR_A_FULL + 8 * slot;
  __syncwarp();
  if (c.lane == 0) {
    if (c.quad == 0 && c.part == 0) {
      const uint32_t slot_sa = c.sm + SM_A_OFF + slot * SLOT_BYTES;
      mbar_arrive_expect_tx_sa(bar, 2 * CHUNK_PLANE_BYTES);
      tma_bulk_g2s_sa(slot_sa, chi, CHUNK_PLANE_BYTES, bar);
      tma_bulk_g2s_sa(slot_sa + SLOT_HALF_BYTES, clo, CHUNK_PLANE_BYTES, bar);
    } else {
      mbar_arrive_sa(bar);
    }
  }
}

// ------------------------------------------------------------------------------------------------ encoder chunks
// Out of line on purpose: the sin/cos tables need many registers and run only 2-3 times per tile; keeping them out of
// the chunk loop keeps the hot path's register allocation tight.  Writes one row (16 columns of part `part`) of the
// chunk into the A ring slot at shared address slot_sa.
template <bool TANGENT>
static __device__ __noinline__ void encode_chunk(uint32_t slot_sa, int src, int part, int row, float p0, float p1,
                                                 float p2, float t, float g0, float g1, float g2, float d0, float d1,
                                                 float d2, int s) {
  EncIn e;
  e.p[0] = p0; e.p[1] = p1; e.p[2] = p2;
  e.t = t;
  e.g[0] = g0; e.g[1] = g1; e.g[2] = g2;
  e.dc[0] = d0; e.dc[1] = d1; e.dc[2] = d2;
  e.s = s;
  float v[PCOLS];
  // part is warp-uniform: no divergence
  static_for<0, NPART>([&](auto pc) {
    constexpr int P = decltype(pc)::value;
    if (part == P) {
      if (src == SRC_ENC_DEFORM) encode_part<SRC_ENC_DEFORM, P, TANGENT>(v, e);
      else if (src == SRC_ENC_SDF) encode_part<SRC_ENC_SDF, P, TANGENT>(v, e);
      else if constexpr (!TANGENT) {
        if (src == SRC_COLOR_A) encode_part<SRC_COLOR_A, P, false>(v, e);
        else if constexpr (PCOLS * P < 32) encode_part<SRC_COLOR_B, P, false>(v, e);  // 32-wide chunk
      }
    }
  });
  if (src == SRC_COLOR_B && PCOLS * part >= 32) {
    // no weights for these columns; keep the slot (and the training record made from it) free of stale values
#pragma unroll
    for (int i = 0; i < PCOLS; ++i) v[i] = 0.f;
  }
  emit_row(slot_sa + 2 * part * A_LBO + row * 16, v);
}

// =================================================================================================================
// plain mode (one row = one point: sdf-query chain, colour chains): thread = TMEM lane, PCOLS columns per chunk
// =================================================================================================================

// sum a per-row float4 across the NPART column-part threads of the row (every one of them gets the total)
__device__ __forceinline__ float4 cross_part_sum_row(const Epi& c, int row, float4 part) {
  const uint32_t xch = c.sm + SM_XCH_OFF;
  sts128<0>(xch + (c.part * TILE_ROWS + row) * 16, __float_as_uint(part.x), __float_as_uint(part.y),
            __float_as_uint(part.z), __float_as_uint(part.w));
  named_bar_sync(1, N_EPI_THREADS);
  float4 t = lds128f(xch + row * 16);
#pragma unroll
  for (int q = 1; q < NPART; ++q) {
    const float4 b = lds128f(xch + (q * TILE_ROWS + row) * 16);
    t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
  }
  named_bar_sync(2, N_EPI_THREADS);  // everyone has read before the buffer is written again
  return t;
}
__device__ __forceinline__ float4 cross_part_sum(const Epi& c, float4 part) { return cross_part_sum_row(c, c.row, part); }

// Read this thread's PCOLS columns of 64-col block `blk` of accumulator buffer `buf`, add bias (row `bias_row` of the
// smem-staged bias table), activate.
template <int ACT>
__device__ __forceinline__ void load_act(const Epi& c, int buf, int blk, int bias_row, float (&v)[PCOLS]) {
  tmem_ld<PCOLS>(c.tmem + buf * HID + 64 * blk, v);
  const uint32_t bias = c.sm + SM_BIAS_OFF + (bias_row * HID + 64 * blk + PCOLS * c.part) * 4;
  float bj[PCOLS];
#pragma unroll
  for (int q = 0; q < PCOLS / 4; ++q) {
    const float4 bb = lds128f(bias + 16 * q);
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
__device__ __forceinline__ void load_act_dyn(const Epi& c, int act, int buf, int blk, int bias, float (&v)[PCOLS]) {
  if (act == ACT_RELU) load_act<ACT_RELU>(c, buf, blk, bias, v);
  else if (act == ACT_SOFTPLUS100) load_act<ACT_SOFTPLUS100>(c, buf, blk, bias, v);
  else load_act<ACT_NONE>(c, buf, blk, bias, v);
}

// raw accumulator columns (no bias / activation): reverse chains
__device__ __forceinline__ void load_raw(const Epi& c, int buf, int blk, float (&v)[PCOLS]) {
  tmem_ld<PCOLS>(c.tmem + buf * HID + 64 * blk, v);
  tmem_ld_wait();
}

// Activation backward for the reverse (training) chains.  u = adjoint of the post-activation values of a forward
// layer for this row; the forward stash holds those post-activation values h (primal rows) / hdot_j (tangent rows).
//   relu     : zbar = [h_primal > 0] * u                                (all rows; relu'' = 0)
//   softplus : sigma = 1 - exp(-100 h_primal)    (h = softplus(z)  =>  sigma(100 z) = 1 - exp(-100 h))
//              tangent rows: zdotbar_j = sigma * u_j
//              primal row  : zbar = sigma * u + 100 (1 - sigma) * sum_j hdot_j * u_j      (softplus'' = 100 s (1-s))
// h: the gating activations of this row (load_planes), fetched BEFORE the accumulator wait so that the HBM / L2
// latency of the record hides behind the MMAs that are still draining
__device__ __forceinline__ void bwd_gate_plain(const float (&h)[PCOLS], int act, float (&u)[PCOLS]) {
  if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < PCOLS; ++i) u[i] = h[i] > 0.f ? u[i] : 0.f;
  } else {
#pragma unroll
    for (int i = 0; i < PCOLS; ++i) u[i] = (1.f - __expf(-100.f * h[i])) * u[i];
  }
}

// acc[j] += sum_i v[i] * w[j][i]   (w row stride 256, uniform loads)
template <int NOUT>
__device__ __forceinline__ void dot_accum(const float (&v)[PCOLS], const float* __restrict__ w, float (&acc)[4]) {
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
    const float4* w4 = reinterpret_cast<const float4*>(w + j * HID);
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
// out = W_out . act(D + bias) summed over the column parts.  Out of line (once or twice per tile).
// keep: training - the layer's input also goes through the A ring as 4 dump-only chunks (the caller advances c.ac).
template <int NOUT>
static __device__ __noinline__ float4 tail_dot(Epi c, uint32_t g_layer, int act, int bias, const float* w_out,
                                               bool keep) {
  wait_d_full(c, g_layer);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int blk = 0; blk < 4; ++blk) {
    float v[PCOLS];
    load_act_dyn(c, act, g_layer & 1, blk, bias, v);
    const int co = 64 * blk + PCOLS * c.part;
    if (keep) {
      const uint32_t slot = claim_slot<true>(c, blk);
      emit_row(c.sm + SM_A_OFF + slot * SLOT_BYTES + 2 * c.part * A_LBO + c.row * 16, v);
      publish_chunk(c, slot, 900 + blk);
      ++c.ac;
    }
    dot_accum<NOUT>(v, w_out + co, acc);
  }
  release_d(c, g_layer);
  return cross_part_sum(c, make_float4(acc[0], acc[1], acc[2], acc[3]));
}

// ------------------------------------------------------------------------------------------------ input adjoints
// Adjoint of the network-input features of one row.  E[i] = d loss / d (column PCOLS*PART + i of encoder chunk SRC of
// this row); the row is stream rs.s of its point (0 primal, j+1 = d/dx_j tangent row, whose entries are the first
// derivatives of the features, so its adjoint meets the second derivatives).  Accumulates d loss / d var[0..9]
// (es_program.h Feat: 0..2 position, 3 time, 4..6 g_c, 7..9 d_c) into xb.
template <int SRC, int PART>
__device__ __forceinline__ void inadj_part(const float (&E)[PCOLS], const EncIn& rs, float (&xb)[10]) {
  float var[10];
  var[0] = rs.p[0]; var[1] = rs.p[1]; var[2] = rs.p[2];
  var[3] = rs.t;
  var[4] = rs.g[0]; var[5] = rs.g[1]; var[6] = rs.g[2];
  var[7] = rs.dc[0]; var[8] = rs.dc[1]; var[9] = rs.dc[2];
  float sn[10][10], cs[10][10];
  static_for<0, PCOLS>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr Feat f = chunk_feat(SRC, PCOLS * PART + i);
    if constexpr (f.var >= 0 && f.freq >= 0) {
      constexpr bool first = (f.is_cos == 0) || !part_has(SRC, PART, f.var, f.freq, 0);
      if constexpr (first)
        sincosf(var[f.var] * static_cast<float>(1 << f.freq), &sn[f.var][f.freq], &cs[f.var][f.freq]);
    }
  });
  const int s = rs.s;
  static_for<0, PCOLS>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr Feat f = chunk_feat(SRC, PCOLS * PART + i);
    if constexpr (f.var >= 0) {
      float d1, d2;
      if constexpr (f.freq < 0) {
        d1 = 1.f;
        d2 = 0.f;
      } else {
        constexpr float fr = static_cast<float>(1 << f.freq);
        d1 = f.is_cos ? -fr * sn[f.var][f.freq] : fr * cs[f.var][f.freq];
        d2 = f.is_cos ? -(fr * fr) * cs[f.var][f.freq] : -(fr * fr) * sn[f.var][f.freq];
      }
      if constexpr (f.var < 3) xb[f.var] += (s == 0) ? E[i] * d1 : ((s - 1 == f.var) ? E[i] * d2 : 0.f);
      else xb[f.var] += (s == 0) ? E[i] * d1 : 0.f;
    }
  });
}

// Read this row's PCOLS raw accumulator columns of part `part` of 64-col block `blk` (= chunk `src` of the network
// input) and push them through inadj_part.  Out of line like encode_chunk (sin/cos tables, once per tile).
static __device__ __noinline__ void inadj_chunk(uint32_t taddr, int src, int part, float p0, float p1, float p2,
                                                float g0, float g1, float g2, float d0, float d1, float d2, int s,
                                                float* xb_out) {
  EncIn e;
  e.p[0] = p0; e.p[1] = p1; e.p[2] = p2;
  e.t = 0.f;
  e.g[0] = g0; e.g[1] = g1; e.g[2] = g2;
  e.dc[0] = d0; e.dc[1] = d1; e.dc[2] = d2;
  e.s = s;
  float E[PCOLS];
  tmem_ld<PCOLS>(taddr, E);
  tmem_ld_wait();
  float xb[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) xb[i] = xb_out[i];
  static_for<0, NPART>([&](auto pc) {
    constexpr int P = decltype(pc)::value;
    if (part == P) {
      if (src == SRC_ENC_SDF) inadj_part<SRC_ENC_SDF, P>(E, e, xb);
      else if (src == SRC_COLOR_A) inadj_part<SRC_COLOR_A, P>(E, e, xb);
      else if constexpr (PCOLS * P < 32) {
        if (src == SRC_COLOR_B) inadj_part<SRC_COLOR_B, P>(E, e, xb);
      }
    }
  });
#pragma unroll
  for (int i = 0; i < 10; ++i) xb_out[i] = xb[i];
}

// ------------------------------------------------------------------------------------------------ gate prefetch
// The reverse chains gate every accumulator chunk with the forward record of the same rows.  Loaded on demand those
// words come from HBM (1500-2500 cycles, pipeline trace of round 2: 3200 cycles per chunk against 1536 cycles of MMAs):
// every thread therefore prefetches the lines it will read one layer ahead into L2 (no registers, no shared memory).
__device__ __forceinline__ bool is_gate_src(int src) { return src == SRC_BWD_PREV || src == SRC_BWD_OUTER3; }
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// off = the thread's offset inside a chunk half (fragment form: frag_off, ns = streams fetched; row form: row_off, ns = 0)
__device__ __forceinline__ void prefetch_chunk_gate(const uint8_t* chunk, int ns) {
  if (ns == 0) {  // row form: load_planes
#pragma unroll
    for (int g = 0; g < PCOLS / 8; ++g) prefetch_l2(chunk + g * A_LBO);
  } else {        // fragment form: load_planes_frag
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      prefetch_l2(chunk + j * A_LBO);
      if (ns == 4) {
#pragma unroll
        for (int s = 1; s < 4; ++s) prefetch_l2(chunk + j * A_LBO + s * 128);
      }
    }
  }
}
__device__ __forceinline__ void prefetch_layer_gates(const uint8_t* gate_tile, const LayerProg& L, uint32_t off,
                                                     bool frag) {
  const int ns = frag ? (L.bwd_act == ACT_RELU ? 1 : 4) : 0;
  for (int ck = 0; ck < L.n_chunks; ++ck) {
    const int src = L.src[ck];
    if (src == SRC_BWD_PREV || src == SRC_BWD_OUTER3)
      prefetch_chunk_gate(gate_tile + static_cast<size_t>(L.gate_base + L.arg[ck]) * CHUNK_PLANE_BYTES + off, ns);
  }
}
// what follows layer l of a reverse chain: layer l + 1, or (after the last layer) the dump-only result chunks of this
// tile and the first layer of the CTA's next tile
__device__ __forceinline__ void prefetch_next_gates(const ChainProg& prog, const ChainIO& io, const uint8_t* gate_tile,
                                                    int l, long long next_tile, long long n_tiles, uint32_t off,
                                                    bool frag) {
  if (l + 1 < prog.n_layers) {
    prefetch_layer_gates(gate_tile, prog.layer[l + 1], off, frag);
    return;
  }
  if (prog.post_op == POST_BWD_DUMP) {
    const int ns = frag ? (prog.post_bwd_act == ACT_RELU ? 1 : 4) : 0;
    for (int blk = 0; blk < 4; ++blk)
      prefetch_chunk_gate(gate_tile + static_cast<size_t>(prog.post_gate_base + blk) * CHUNK_PLANE_BYTES + off, ns);
  }
  if (next_tile < n_tiles)
    prefetch_layer_gates(io.gate_hi + static_cast<size_t>(next_tile) * prog.n_gate * CHUNK_PLANE_BYTES, prog.layer[0],
                         off, frag);
}

template <int CHAIN, bool BWD, bool STASH>
__device__ __forceinline__ void epilogue_plain(const ChainProg& prog, const ChainIO& io, Epi& c, const TileWalk tw) {
  constexpr bool DUMP = BWD || STASH;
  const float scale = (BWD && io.scale) ? __ldg(io.scale) : 1.f;
  const uint32_t row_off = 2 * c.part * A_LBO + c.row * 16;  // this thread's 16-byte units inside a chunk half
  long long tile = tw.tile0;
  for (long long it = 0; it < tw.n_iter; ++it, tile += tw.step) {
    const long long tile_r = tile < tw.n_tiles ? tile : tw.n_tiles - 1;  // record index that is safe to read
    // ---------------------------------------------------------- row state
    float xc[3] = {0.f, 0.f, 0.f};        // canonical point (after the deform tail / = x without deform)
    float adj[4] = {0.f, 0.f, 0.f, 0.f};  // reverse chains: (o.x, o.y, o.z, r) adjoint of this row's outputs
    const long long p_raw = tile * TILE_ROWS + c.row;
    const bool valid = p_raw < io.n_points;
    const long long pt = valid ? p_raw : io.n_points - 1;
    if constexpr (BWD) {
      if (valid && io.adj) {  // padding rows carry zero adjoints so they add nothing to the weight gradients
        const float4 a = __ldg(reinterpret_cast<const float4*>(io.adj) + pt);
        adj[0] = a.x * scale; adj[1] = a.y * scale; adj[2] = a.z * scale; adj[3] = a.w * scale;
      }
    } else if constexpr (CHAIN == CHAIN_COLOR) {
#pragma unroll
      for (int i = 0; i < 3; ++i) xc[i] = __ldg(io.x_c + pt * 3 + i);
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) xc[i] = __ldg(io.x + pt * 3 + i);
    }
    float sdf_acc[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (!BWD && CHAIN == CHAIN_COLOR) {
      // the geometry feature rows (1 KiB per point, read on demand by the SRC_FEAT chunks of layer 0 and of the skip
      // layer) of the CTA's NEXT tile go to L2 now: a whole tile of MMAs later the loads no longer wait for HBM
      const long long pn = (tile + tw.step) * TILE_ROWS + c.row;
      if (pn < io.n_points && io.feat && !io.feat_rec) {
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) prefetch_l2(io.feat + pn * HID + 64 * blk + PCOLS * c.part);
      }
    }
    const uint8_t* gate_tile =
        BWD ? io.gate_hi + static_cast<size_t>(tile_r) * prog.n_gate * CHUNK_PLANE_BYTES : nullptr;
    const uint8_t* gate_tile_lo =
        (BWD && prog.gate_use_lo) ? io.gate_lo + static_cast<size_t>(tile_r) * prog.n_gate * CHUNK_PLANE_BYTES : nullptr;

    for (int l = 0; l < prog.n_layers; ++l, ++c.g) {
      const LayerProg& L = prog.layer[l];
      const int bias_prev = l > 0 ? l - 1 : 0;  // row of the smem bias table
      const int act_prev = l > 0 ? prog.layer[l - 1].act : ACT_NONE;
      const int last_prev = L.last_prev;
      const int n_chunks = L.n_chunks;
      bool prev_waited = false;
      if constexpr (BWD) {
        if (io.gate_hi && !ES_FLAG(io, 8))
          prefetch_next_gates(prog, io, gate_tile, l, tile + tw.step, tw.n_tiles, row_off, false);
      }

      if (!BWD && L.pre_op == PRE_DEFORM_TAIL) {
        // deform output layer (3 x 256, fp32 FFMA) -> x_c = x + delta
        float4 r = tail_dot<3>(c, c.g - 1, act_prev, bias_prev, prog.deform_out_w, false);
        xc[0] += r.x + __ldg(prog.deform_out_b + 0);
        xc[1] += r.y + __ldg(prog.deform_out_b + 1);
        xc[2] += r.z + __ldg(prog.deform_out_b + 2);
        if (c.part == 0 && valid && io.out_xc) {
#pragma unroll
          for (int i = 0; i < 3; ++i) io.out_xc[pt * 3 + i] = xc[i];
        }
        prev_waited = true;  // (already consumed and released)
      }

      for (int ck = 0; ck < n_chunks; ++ck, ++c.ac) {
        const uint32_t slot = claim_slot<DUMP>(c, ck);
        const uint32_t slot_sa = c.sm + SM_A_OFF + slot * SLOT_BYTES;
        const uint32_t row_sa = slot_sa + row_off;
        const int src = L.src[ck];
        const int col0 = 64 * L.arg[ck] + PCOLS * c.part;
        if (BWD && (src == SRC_BWD_PREV || src == SRC_BWD_OUTER3)) {
          float v[PCOLS];
          // gating activations first (hi-only records: raw words), so that their L2 latency overlaps the accumulator
          // wait and the TMEM load
          const size_t go = static_cast<size_t>(L.gate_base + L.arg[ck]) * CHUNK_PLANE_BYTES + row_off;
          uint4 graw[PCOLS / 8];
          if (!gate_tile_lo) load_gate_raw(gate_tile + go, graw);
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
          {
            float hg[PCOLS];
            if (gate_tile_lo) load_planes(gate_tile + go, gate_tile_lo + go, hg);
            else unpack_gate_raw(graw, hg);
            bwd_gate_plain(hg, L.bwd_act, v);
          }
          emit_row(row_sa, v);
        } else if (BWD && src == SRC_PLANE) {
          const uint8_t* chi =
              io.plane_hi + (static_cast<size_t>(tile_r) * prog.n_plane + L.arg[ck]) * CHUNK_PLANE_BYTES;
          const uint8_t* clo = prog.plane_lo[ck] != NO_DUMP
                                   ? io.plane_lo + (static_cast<size_t>(tile_r) * prog.n_plane_lo + prog.plane_lo[ck]) *
                                                       CHUNK_PLANE_BYTES
                                   : nullptr;
          if (clo) {
            reload_chunk_bulk(c, slot, chi, clo);
            continue;
          }
          copy_plane_row(slot_sa, chi, clo, row_off);
        } else if (src == SRC_PREV) {
          float v[PCOLS];
          if (!prev_waited) {
            wait_d_full(c, c.g - 1);
            prev_waited = true;
          }
          load_act_dyn(c, act_prev, (c.g - 1) & 1, L.arg[ck], bias_prev, v);
          if (L.side_dot) dot_accum<1>(v, prog.sdf_out_w + col0, sdf_acc);
          if (ck == last_prev) release_d(c, c.g - 1);
          TRACE_EPI(7000 + l * 16 + ck);  // EPI: values ready (tmem + math done)
          emit_row(row_sa, v);
        } else if (src == SRC_FEAT && io.feat_rec) {
          // the chunk as the geometry chain left it: two 16 KiB bulk copies straight into the ring slot
          const uint8_t* chi = io.feat_rec + (static_cast<size_t>(tile_r) * 4 + L.arg[ck]) * (2 * CHUNK_PLANE_BYTES);
          reload_chunk_bulk(c, slot, chi, chi + CHUNK_PLANE_BYTES);
          continue;
        } else if (src == SRC_FEAT) {
          float v[PCOLS];
          const float4* f4 = reinterpret_cast<const float4*>(io.feat + pt * HID + col0);
#pragma unroll
          for (int q = 0; q < PCOLS / 4; ++q) {
            float4 f = __ldg(f4 + q);
            v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
          }
          emit_row(row_sa, v);
        } else if (src == SRC_ENC_DEFORM) {
          const float* xp = io.x + pt * 3;
          encode_chunk<false>(slot_sa, src, c.part, c.row, __ldg(xp), __ldg(xp + 1), __ldg(xp + 2),
                              __ldg(io.t + (pt / io.t_div) * io.t_stride), 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0);
        } else if (src == SRC_ENC_SDF) {
          encode_chunk<false>(slot_sa, src, c.part, c.row, xc[0], xc[1], xc[2], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0);
        } else {  // SRC_COLOR_A / SRC_COLOR_B: [enc10(x_c), g_c, enc4(d_c)]
          const float* gc = io.g_c + pt * 3;
          const float* d = io.dirs + (pt / io.dir_div) * io.dir_stride;
          const float dd[3] = {__ldg(d), __ldg(d + 1), __ldg(d + 2)};
          float dcn[3];
          if (io.jac) {
            // d_c = J d  (reference endosurf.py:684: bmm(pts_jacobian, d)), J[i][j] = d x_c_i / d x_j
            const float* J = io.jac + pt * 9;
#pragma unroll
            for (int i = 0; i < 3; ++i)
              dcn[i] = __ldg(J + 3 * i) * dd[0] + __ldg(J + 3 * i + 1) * dd[1] + __ldg(J + 3 * i + 2) * dd[2];
          } else {
#pragma unroll
            for (int i = 0; i < 3; ++i) dcn[i] = dd[i];
          }
          const float nrm = sqrtf(dcn[0] * dcn[0] + dcn[1] * dcn[1] + dcn[2] * dcn[2]) + 1e-10f;
          encode_chunk<false>(slot_sa, src, c.part, c.row, xc[0], xc[1], xc[2], 0.f, __ldg(gc), __ldg(gc + 1),
                              __ldg(gc + 2), dcn[0] / nrm, dcn[1] / nrm, dcn[2] / nrm, 0);
        }
        publish_chunk(c, slot, l * 16 + ck);
      }

      if (!BWD && L.side_dot) {
        // sdf row of the SDF output layer
        float4 r = cross_part_sum(c, make_float4(sdf_acc[0], 0.f, 0.f, 0.f));
        if (c.part == 0 && valid && io.out_sdf) io.out_sdf[pt] = r.x + __ldg(prog.sdf_out_b);
      }
    }

    // ---------------------------------------------------------- post op: consume the last accumulator
    const int last = prog.n_layers - 1;
    const int act_last = prog.layer[last].act;
    if (!BWD && prog.post_op == POST_SDF_TAIL) {
      float4 r = tail_dot<1>(c, c.g - 1, act_last, last, prog.sdf_out_w, false);
      if (c.part == 0 && valid && io.out_sdf) io.out_sdf[pt] = r.x + __ldg(prog.sdf_out_b);
    } else if (BWD && prog.post_op == POST_BWD_DUMP) {
      // adjoint of the first layer's pre-activation: feeds no MMA of this launch, goes out as 4 dump-only chunks
      wait_d_full(c, c.g - 1);
#pragma unroll 1
      for (int blk = 0; blk < 4; ++blk, ++c.ac) {
        float v[PCOLS], hg[PCOLS];
        const size_t go = static_cast<size_t>(prog.post_gate_base + blk) * CHUNK_PLANE_BYTES + row_off;
        uint4 graw[PCOLS / 8];
        if (!gate_tile_lo) load_gate_raw(gate_tile + go, graw);  // before the TMEM load: latencies overlap
        load_raw(c, (c.g - 1) & 1, blk, v);
        if (gate_tile_lo) load_planes(gate_tile + go, gate_tile_lo + go, hg);
        else unpack_gate_raw(graw, hg);
        bwd_gate_plain(hg, prog.post_bwd_act, v);
        const uint32_t slot = claim_slot<true>(c, blk);
        emit_row(c.sm + SM_A_OFF + slot * SLOT_BYTES + row_off, v);
        publish_chunk(c, slot, 950 + blk);
      }
      release_d(c, c.g - 1);
    } else if (!BWD && prog.post_op == POST_COLOR_TAIL) {
      float4 r = tail_dot<3>(c, c.g - 1, act_last, last, prog.color_out_w, STASH);
      if constexpr (STASH) c.ac += 4;
      if (c.part == 0 && valid) {
        float o[3] = {r.x, r.y, r.z};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float z = o[i] + __ldg(prog.color_out_b + i);
          io.out_rgb[pt * 3 + i] = 1.f / (1.f + expf(-z));
        }
      }
    } else if (BWD && prog.post_op == POST_FEAT_BAR) {
      // d loss / d feat = zbar_0 W_0[:, feat] + zbar_skip W_skip[:, feat] / sqrt 2 (colour net), unscaled again
      wait_d_full(c, c.g - 1);
      const float inv = 1.f / scale;
      float m = 0.f;
#pragma unroll 1
      for (int blk = 0; blk < 4; ++blk) {
        float v[PCOLS];
        load_raw(c, (c.g - 1) & 1, blk, v);
#pragma unroll
        for (int i = 0; i < PCOLS; ++i) {
          v[i] *= inv;
          m = fmaxf(m, fabsf(v[i]));
        }
        if (valid) {
          float4* o = reinterpret_cast<float4*>(io.feat_bar + pt * HID + 64 * blk + PCOLS * c.part);
#pragma unroll
          for (int q = 0; q < PCOLS / 4; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
      release_d(c, c.g - 1);
      if (!valid) m = 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (c.lane == 0 && io.amax_bits) atomicMax(io.amax_bits, __float_as_uint(m));
    } else if (BWD && prog.post_op == POST_INADJ_COLOR) {
      // adjoint of [enc10(x_c), g_c, enc4(d_c)] (accumulator columns 0..95 in the kernel's chunk order) pushed back to
      // x_c, g_c and, through d_c = normalize(J d) (endosurf.py:684-685), to J
      wait_d_full(c, c.g - 1);
      const float inv = 1.f / scale;
      float xb[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) xb[i] = 0.f;
      float px[3], gcv[3], dd[3], u[3], dcn[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        px[i] = __ldg(io.x_c + pt * 3 + i);
        gcv[i] = __ldg(io.g_c + pt * 3 + i);
        dd[i] = __ldg(io.dirs + (pt / io.dir_div) * io.dir_stride + i);
      }
      if (io.jac) {
        const float* J = io.jac + pt * 9;
#pragma unroll
        for (int i = 0; i < 3; ++i)
          u[i] = __ldg(J + 3 * i) * dd[0] + __ldg(J + 3 * i + 1) * dd[1] + __ldg(J + 3 * i + 2) * dd[2];
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) u[i] = dd[i];
      }
      const float un = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
      const float nrm = un + 1e-10f;
#pragma unroll
      for (int i = 0; i < 3; ++i) dcn[i] = u[i] / nrm;
      const uint32_t ta = c.tmem + ((c.g - 1) & 1) * HID;
      inadj_chunk(ta, SRC_COLOR_A, c.part, px[0], px[1], px[2], gcv[0], gcv[1], gcv[2], dcn[0], dcn[1], dcn[2], 0, xb);
      if (PCOLS * c.part < 32)
        inadj_chunk(ta + 64, SRC_COLOR_B, c.part, px[0], px[1], px[2], gcv[0], gcv[1], gcv[2], dcn[0], dcn[1], dcn[2],
                    0, xb);
      release_d(c, c.g - 1);
      const float4 sx = cross_part_sum(c, make_float4(xb[0], xb[1], xb[2], 0.f));
      const float4 sg = cross_part_sum(c, make_float4(xb[4], xb[5], xb[6], 0.f));
      const float4 sd = cross_part_sum(c, make_float4(xb[7], xb[8], xb[9], 0.f));
      if (c.part == 0 && valid) {
        float4* as = reinterpret_cast<float4*>(io.adj_sdf) + pt * 4;
        const float gb[3] = {sg.x * inv, sg.y * inv, sg.z * inv};
#pragma unroll
        for (int j = 0; j < 3; ++j) as[1 + j].w += gb[j];
        if (io.adj_deform) {
          float4* ad = reinterpret_cast<float4*>(io.adj_deform) + pt * 4;
          ad[0].x += sx.x * inv;
          ad[0].y += sx.y * inv;
          ad[0].z += sx.z * inv;
          // d_c = u / (|u| + eps):  ubar = dbar / (|u| + eps) - u (u . dbar) / (|u| (|u| + eps)^2);  Jbar[i][j] = ubar_i d_j
          const float db[3] = {sd.x * inv, sd.y * inv, sd.z * inv};
          const float udb = u[0] * db[0] + u[1] * db[1] + u[2] * db[2];
          const float k = udb / (fmaxf(un, 1e-30f) * nrm * nrm);
          float ub[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) ub[i] = db[i] / nrm - u[i] * k;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            ad[1 + j].x += ub[0] * dd[j];
            ad[1 + j].y += ub[1] * dd[j];
            ad[1 + j].z += ub[2] * dd[j];
          }
        }
      }
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

__device__ __forceinline__ void load_frag(uint32_t ta, Frag& F) {
  tmem_ld_16x256b_x2(ta, F.f[0][0], F.f[0][1], F.f[1][0], F.f[1][1], F.f[0][2], F.f[0][3], F.f[1][2], F.f[1][3]);
  tmem_ld_16x256b_x2(ta + (16u << 16), F.f[2][0], F.f[2][1], F.f[3][0], F.f[3][1], F.f[2][2], F.f[2][3], F.f[3][2],
                     F.f[3][3]);
  tmem_ld_wait();
}

// bias (shared address of this thread's col(0) entry) + activation of the primal stream, chain rule on the tangents
template <int ACT>
__device__ __forceinline__ void act_frag(uint32_t bias_sa, Frag& F) {
  const float2 b0 = lds64f(bias_sa);
  const float2 b1 = lds64f(bias_sa + 32);
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
__device__ __forceinline__ void act_frag_dyn(int act, uint32_t bias_sa, Frag& F) {
  if (act == ACT_RELU) act_frag<ACT_RELU>(bias_sa, F);
  else if (act == ACT_SOFTPLUS100) act_frag<ACT_SOFTPLUS100>(bias_sa, F);
  else act_frag<ACT_NONE>(bias_sa, F);
}

// split to fp16 hi/lo and store into an A ring slot; sa = this thread's base inside the slot (k-group 2 part,
// row 32Q + p, byte 4q)
__device__ __forceinline__ void emit_frag(const Frag& F, uint32_t sa) {
  static_for<0, 4>([&](auto sc) {
    constexpr int s = decltype(sc)::value;
    static_for<0, 2>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      uint32_t hi, lo;
      split2(F.f[s][2 * j], F.f[s][2 * j + 1], hi, lo);
      sts32<j * A_LBO + s * 128>(sa, hi);
      sts32<SLOT_HALF_BYTES + j * A_LBO + s * 128>(sa, lo);
    });
  });
}

// fragment of a dumped plane chunk (value = hi + lo, plo may be null); pointers at the chunk half + the same
// per-thread offset as in the ring slot (k-group 2 part, row 32Q + p, byte 4q).  A warp-wide 4-byte load covers 8
// consecutive rows x 16 bytes = one full 128-byte line.
// NS = 1: only the primal stream is fetched (ReLU gates need nothing else: relu'' = 0), the tangent entries of H are
// left untouched - a quarter of the record's lines.
template <int NS>
__device__ __forceinline__ void load_planes_frag_n(const uint8_t* phi, const uint8_t* plo, Frag& H) {
#pragma unroll
  for (int s = 0; s < NS; ++s) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(phi + j * A_LBO + s * 128));
      const uint32_t b = plo ? __ldg(reinterpret_cast<const uint32_t*>(plo + j * A_LBO + s * 128)) : 0u;
      const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&a));
      const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&b));
      H.f[s][2 * j] = fa.x + fb.x;
      H.f[s][2 * j + 1] = fa.y + fb.y;
    }
  }
}
__device__ __forceinline__ void load_planes_frag(const uint8_t* phi, const uint8_t* plo, Frag& H, int n_streams = 4) {
  if (n_streams == 1) load_planes_frag_n<1>(phi, plo, H);
  else load_planes_frag_n<4>(phi, plo, H);
}

// hi-plane gate words of a fragment, raw (index 2 s + j).  They are requested before the accumulator wait / TMEM load and
// unpacked only at their use, so that the first instruction that needs them comes after the TMEM load has been issued
// (warp-stall sampling of the reverse SDF chain: 14 % of all samples sat on the unpack that followed the loads directly).
__device__ __forceinline__ void fetch_gate_raw(const uint8_t* phi, int ns, uint32_t (&g)[8]) {
#pragma unroll
  for (int j = 0; j < 2; ++j) g[j] = __ldg(reinterpret_cast<const uint32_t*>(phi + j * A_LBO));
  if (ns == 4) {
#pragma unroll
    for (int s = 1; s < 4; ++s)
#pragma unroll
      for (int j = 0; j < 2; ++j) g[2 * s + j] = __ldg(reinterpret_cast<const uint32_t*>(phi + j * A_LBO + s * 128));
  }
}
__device__ __forceinline__ float2 unpack_h2(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
// bwd_gate_frag on raw hi-plane words
__device__ __forceinline__ void bwd_gate_frag_raw(const uint32_t (&g)[8], int act, Frag& U) {
  if (act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float2 h = unpack_h2(g[j]);
      const bool m0 = h.x > 0.f, m1 = h.y > 0.f;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        U.f[s][2 * j] = m0 ? U.f[s][2 * j] : 0.f;
        U.f[s][2 * j + 1] = m1 ? U.f[s][2 * j + 1] : 0.f;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float2 h0 = unpack_h2(g[j]), h1 = unpack_h2(g[2 + j]), h2 = unpack_h2(g[4 + j]), h3 = unpack_h2(g[6 + j]);
      const float hp[2] = {h0.x, h0.y}, t1[2] = {h1.x, h1.y}, t2[2] = {h2.x, h2.y}, t3[2] = {h3.x, h3.y};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = 2 * j + e;
        const float ex = __expf(-100.f * hp[e]);  // 1 - sigma
        const float sg = 1.f - ex;
        const float ct = t1[e] * U.f[1][i] + t2[e] * U.f[2][i] + t3[e] * U.f[3][i];
        U.f[0][i] = fmaf(100.f * ex, ct, sg * U.f[0][i]);
        U.f[1][i] *= sg;
        U.f[2][i] *= sg;
        U.f[3][i] *= sg;
      }
    }
  }
}

// activation backward on a fragment (see bwd_gate_plain for the formulas); everything is thread-local
__device__ __forceinline__ void bwd_gate_frag(const Frag& H, int act, Frag& U) {
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

// acc[s][o] += sum_c F[s][c] * w[o][col(c)]      (w points at col(0))
template <int NOUT>
__device__ __forceinline__ void dot_frag(const Frag& F, const float* __restrict__ w, float (&acc)[4][NOUT]) {
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    const float2 w0 = __ldg(reinterpret_cast<const float2*>(w + o * HID));
    const float2 w1 = __ldg(reinterpret_cast<const float2*>(w + o * HID + 8));
#pragma unroll
    for (int s = 0; s < 4; ++s)
      acc[s][o] += fmaf(F.f[s][0], w0.x, F.f[s][1] * w0.y) + fmaf(F.f[s][2], w1.x, F.f[s][3] * w1.y);
  }
}

// total of a per-point vector over the 4 column lanes of the quad and the NPART column-part warps; every thread of
// the point gets the result.  NV = 4 or 12.
template <int NV>
__device__ __forceinline__ void point_sum(const Epi& c, float (&v)[NV]) {
  static_assert(NV % 4 == 0 && NPART * TILE_PTS_T * NV * 4 <= SM_XCH_BYTES, "exchange buffer");
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1);
    v[i] += __shfl_xor_sync(0xffffffffu, v[i], 2);
  }
  const uint32_t xch = c.sm + SM_XCH_OFF;
  const int ptl = 8 * c.quad + (c.lane >> 2);
  if ((c.lane & 3) == 0) {
#pragma unroll
    for (int i = 0; i < NV / 4; ++i)
      sts128<0>(xch + ((c.part * TILE_PTS_T + ptl) * (NV / 4) + i) * 16, __float_as_uint(v[4 * i]),
                __float_as_uint(v[4 * i + 1]), __float_as_uint(v[4 * i + 2]), __float_as_uint(v[4 * i + 3]));
  }
  named_bar_sync(1, N_EPI_THREADS);
#pragma unroll
  for (int i = 0; i < NV / 4; ++i) {
    float4 t = lds128f(xch + (ptl * (NV / 4) + i) * 16);
#pragma unroll
    for (int q = 1; q < NPART; ++q) {
      const float4 b = lds128f(xch + ((q * TILE_PTS_T + ptl) * (NV / 4) + i) * 16);
      t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
    }
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
  named_bar_sync(2, N_EPI_THREADS);  // everyone has read before the buffer is written again
}

// Consume the whole accumulator of layer g_layer through a NOUT-wide fp32 output layer: out[s * NOUT + o] for the 4
// streams.  keep: training - the layer's input also goes through the A ring as 4 dump-only chunks (the caller
// advances c.ac); frag_off = this thread's fragment offset inside a ring slot.  Out of line.
template <int NOUT>
static __device__ __noinline__ void tail_frag(Epi c, uint32_t g_layer, int act, int bias, const float* w_out,
                                              bool keep, uint32_t frag_off, float* out) {
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
    load_frag(c.tmem + (g_layer & 1) * HID + 64 * blk, F);
    act_frag_dyn(act, c.sm + SM_BIAS_OFF + (bias * HID + 64 * blk + colq) * 4, F);
    if (keep) {
      const uint32_t slot = claim_slot<true>(c, blk);
      emit_frag(F, c.sm + SM_A_OFF + slot * SLOT_BYTES + frag_off);
      publish_chunk(c, slot, 900 + blk);
      ++c.ac;
    }
    dot_frag<NOUT>(F, w_out + 64 * blk + colq, acc);
  }
  release_d(c, g_layer);
  float o[4 * NOUT];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int k = 0; k < NOUT; ++k) o[s * NOUT + k] = acc[s][k];
  point_sum<4 * NOUT>(c, o);
#pragma unroll
  for (int i = 0; i < 4 * NOUT; ++i) out[i] = o[i];
}

template <bool BWD, bool STASH>
__device__ __forceinline__ void epilogue_tangent(const ChainProg& prog, const ChainIO& io, Epi& c, const TileWalk tw) {
  constexpr bool DUMP = BWD || STASH;
  const int q = c.lane & 3, p = c.lane >> 2;
  const int colq = PCOLS * c.part + 2 * q;                                        // col(0) inside a 64-column block
  const uint32_t frag_off = 2 * c.part * A_LBO + (32 * c.quad + p) * 16 + 4 * q;  // see emit_frag
  const uint32_t bias_sa0 = c.sm + SM_BIAS_OFF + colq * 4;
  const bool writer = (c.part == 0);
  const float scale = (BWD && io.scale) ? __ldg(io.scale) : 1.f;

  long long tile = tw.tile0;
  for (long long it = 0; it < tw.n_iter; ++it, tile += tw.step) {
    const long long tile_r = tile < tw.n_tiles ? tile : tw.n_tiles - 1;  // record index that is safe to read
    // ---------------------------------------------------------- point state (same for the 4 lanes of a quad)
    const long long p_raw = tile * TILE_PTS_T + 8 * c.quad + p;
    const bool valid = p_raw < io.n_points;
    const long long pt = valid ? p_raw : io.n_points - 1;
    float xc[3] = {0.f, 0.f, 0.f};
    if constexpr (!BWD) {
#pragma unroll
      for (int i = 0; i < 3; ++i) xc[i] = __ldg(io.x + pt * 3 + i);
    }
    float sdf_acc[4][1] = {{0.f}, {0.f}, {0.f}, {0.f}};
    if constexpr (BWD) {
      // reverse SDF chain: the adjoint of the geometry feature (fp32 [P,256], read on demand by the SRC_ADJ_FEAT chunks
      // of layer 0) of the CTA's next tile goes to L2 now
      const long long pn = (tile + tw.step) * TILE_PTS_T + 8 * c.quad + p;
      if (io.adj_feat && pn < io.n_points) {
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) prefetch_l2(io.adj_feat + pn * HID + 64 * blk + colq);
      }
    }
    const uint8_t* gate_tile =
        BWD ? io.gate_hi + static_cast<size_t>(tile_r) * prog.n_gate * CHUNK_PLANE_BYTES : nullptr;
    const uint8_t* gate_tile_lo =
        (BWD && prog.gate_use_lo) ? io.gate_lo + static_cast<size_t>(tile_r) * prog.n_gate * CHUNK_PLANE_BYTES : nullptr;

    for (int l = 0; l < prog.n_layers; ++l, ++c.g) {
      const LayerProg& L = prog.layer[l];
      const int bias_prev = l > 0 ? l - 1 : 0;  // row of the smem bias table
      const int act_prev = l > 0 ? prog.layer[l - 1].act : ACT_NONE;
      const int last_prev = L.last_prev;
      const int n_chunks = L.n_chunks;
      const bool side_dot = L.side_dot != 0;
      bool prev_waited = false;
      if constexpr (BWD) {
        if (!ES_FLAG(io, 8)) prefetch_next_gates(prog, io, gate_tile, l, tile + tw.step, tw.n_tiles, frag_off, true);
      }

      if (!BWD && L.pre_op == PRE_DEFORM_TAIL) {
        // deform output layer (3 x 256, fp32 FFMA) -> x_c = x + delta ; tangent streams give dDelta/dx_{s-1}
        // training: the output layer's input leaves as 4 dump-only chunks (this layer has no SRC_PREV chunk of its own)
        float o[12];
        tail_frag<3>(c, c.g - 1, act_prev, bias_prev, prog.deform_out_w, STASH, frag_off, o);
        TRACE_EPI(8000 + l);  // EPI: output-layer tail done
        if constexpr (STASH) c.ac += 4;
#pragma unroll
        for (int i = 0; i < 3; ++i) xc[i] += o[i] + __ldg(prog.deform_out_b + i);
        if (writer && valid) {
          if (q == 0) {
            if (io.out_xc) {
#pragma unroll
              for (int i = 0; i < 3; ++i) io.out_xc[pt * 3 + i] = xc[i];
            }
          } else if (io.out_jac) {
            // column j = q-1 of J = I + dDelta/dx ; J stored [i][j] row-major
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const float dl = q == 1 ? o[3 + i] : (q == 2 ? o[6 + i] : o[9 + i]);
              io.out_jac[pt * 9 + 3 * i + (q - 1)] = dl + ((i == q - 1) ? 1.f : 0.f);
            }
          }
        }
        prev_waited = true;  // (already consumed and released)
        TRACE_EPI(9000 + l);  // EPI: tail outputs stored
      }

      for (int ck = 0; ck < n_chunks; ++ck, ++c.ac) {
        const uint32_t slot = claim_slot<DUMP>(c, ck);
        if constexpr (DUMP) TRACE_EPI(6000 + l * 16 + ck);  // EPI: ring slot free
        const uint32_t slot_sa = c.sm + SM_A_OFF + slot * SLOT_BYTES;
        const int src = L.src[ck];
        const int blk = L.arg[ck];
        if constexpr (!BWD) {
          if (src == SRC_PREV) {
            if (!prev_waited) {
              wait_d_full(c, c.g - 1);
              prev_waited = true;
            }
            Frag F;
            load_frag(c.tmem + ((c.g - 1) & 1) * HID + 64 * blk, F);
            act_frag_dyn(act_prev, bias_sa0 + (bias_prev * HID + 64 * blk) * 4, F);
            if (side_dot) dot_frag<1>(F, prog.sdf_out_w + 64 * blk + colq, sdf_acc);
            if (ck == last_prev) release_d(c, c.g - 1);
            TRACE_EPI(7000 + l * 16 + ck);  // EPI: values ready (tmem + math done)
            emit_frag(F, slot_sa + frag_off);
          } else {
            // positional encodings: row-wise, this thread is row 32Q + 8q + p (stream q of its point)
            if (src == SRC_ENC_DEFORM) {
              const float* xp = io.x + pt * 3;
              encode_chunk<true>(slot_sa, src, c.part, c.row, __ldg(xp), __ldg(xp + 1), __ldg(xp + 2),
                                 __ldg(io.t + (pt / io.t_div) * io.t_stride), 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, q);
            } else {
              encode_chunk<true>(slot_sa, src, c.part, c.row, xc[0], xc[1], xc[2], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, q);
            }
          }
        } else {
          Frag F;
          if (src == SRC_PLANE) {
            const uint8_t* chi = io.plane_hi + (static_cast<size_t>(tile_r) * prog.n_plane + blk) * CHUNK_PLANE_BYTES;
            const uint8_t* clo = prog.plane_lo[ck] != NO_DUMP
                                     ? io.plane_lo + (static_cast<size_t>(tile_r) * prog.n_plane_lo + prog.plane_lo[ck]) *
                                                         CHUNK_PLANE_BYTES
                                     : nullptr;
            if (clo) {
              reload_chunk_bulk(c, slot, chi, clo);
              continue;
            }
            copy_plane_row(slot_sa, chi, clo, 2 * c.part * A_LBO + c.row * 16);
          } else if (src == SRC_ADJ_FEAT) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
              for (int i = 0; i < 4; ++i) F.f[s][i] = 0.f;
            if (valid) {
              const float* fp = io.adj_feat + pt * HID + 64 * blk + colq;
              const float2 f0 = __ldg(reinterpret_cast<const float2*>(fp));
              const float2 f1 = __ldg(reinterpret_cast<const float2*>(fp + 8));
              F.f[0][0] = f0.x * scale; F.f[0][1] = f0.y * scale; F.f[0][2] = f1.x * scale; F.f[0][3] = f1.y * scale;
            }
            emit_frag(F, slot_sa + frag_off);
          } else {
            // gating activations first (requested, not yet used): their latency (L2, see prefetch_next_gates) overlaps
            // the accumulator wait and the TMEM load
            Frag HG;
            uint32_t graw[8];
            const size_t go = static_cast<size_t>(L.gate_base + blk) * CHUNK_PLANE_BYTES + frag_off;
            if (gate_tile_lo) load_planes_frag(gate_tile + go, gate_tile_lo + go, HG, L.bwd_act == ACT_RELU ? 1 : 4);
            else fetch_gate_raw(gate_tile + go, L.bwd_act == ACT_RELU ? 1 : 4, graw);
            if (src == SRC_BWD_PREV) {
              if (!prev_waited) {
                wait_d_full(c, c.g - 1);
                prev_waited = true;
              }
              load_frag(c.tmem + ((c.g - 1) & 1) * HID + 64 * blk, F);
              if (L.rank1) {
                // + adjoint of the sdf-row output (r) times the sdf row of the output layer; padding points: zero
                const float2 w0 = __ldg(reinterpret_cast<const float2*>(prog.sdf_out_w + 64 * blk + colq));
                const float2 w1 = __ldg(reinterpret_cast<const float2*>(prog.sdf_out_w + 64 * blk + colq + 8));
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                  const float aw = valid ? __ldg(io.adj + (pt * 4 + s) * 4 + 3) * scale : 0.f;
                  F.f[s][0] = fmaf(aw, w0.x, F.f[s][0]);
                  F.f[s][1] = fmaf(aw, w0.y, F.f[s][1]);
                  F.f[s][2] = fmaf(aw, w1.x, F.f[s][2]);
                  F.f[s][3] = fmaf(aw, w1.y, F.f[s][3]);
                }
              }
              if (ck == last_prev) release_d(c, c.g - 1);
            } else {  // SRC_BWD_OUTER3: adjoints (o.x, o.y, o.z) of the 4 streams' 3-wide outputs times W_out
              float w[3][4];
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                const float2 w0 = __ldg(reinterpret_cast<const float2*>(prog.outer3_w + i * HID + 64 * blk + colq));
                const float2 w1 = __ldg(reinterpret_cast<const float2*>(prog.outer3_w + i * HID + 64 * blk + colq + 8));
                w[i][0] = w0.x; w[i][1] = w0.y; w[i][2] = w1.x; w[i][3] = w1.y;
              }
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const float4 a = valid ? __ldg(reinterpret_cast<const float4*>(io.adj) + pt * 4 + s)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 4; ++i) F.f[s][i] = (a.x * w[0][i] + a.y * w[1][i] + a.z * w[2][i]) * scale;
              }
            }
            if (gate_tile_lo) bwd_gate_frag(HG, L.bwd_act, F);
            else bwd_gate_frag_raw(graw, L.bwd_act, F);
            emit_frag(F, slot_sa + frag_off);
          }
        }
        publish_chunk(c, slot, l * 16 + ck);
      }

      if (!BWD && side_dot) {
        // sdf row of the SDF output layer: primal stream -> sdf, tangent streams -> g_c
        float o[4] = {sdf_acc[0][0], sdf_acc[1][0], sdf_acc[2][0], sdf_acc[3][0]};
        point_sum<4>(c, o);
        if (writer && valid) {
          if (q == 0) {
            if (io.out_sdf) io.out_sdf[pt] = o[0] + __ldg(prog.sdf_out_b);
          } else if (io.out_gc) {
            io.out_gc[pt * 3 + (q - 1)] = q == 1 ? o[1] : (q == 2 ? o[2] : o[3]);
          }
        }
      }
    }

    // ---------------------------------------------------------- post op: consume the last accumulator
    const int last = prog.n_layers - 1;
    if (!BWD && prog.post_op == POST_SDF_TAIL) {
      float o[4];
      tail_frag<1>(c, c.g - 1, prog.layer[last].act, last, prog.sdf_out_w, false, frag_off, o);
      if (writer && valid) {
        if (q == 0) {
          if (io.out_sdf) io.out_sdf[pt] = o[0] + __ldg(prog.sdf_out_b);
        } else if (io.out_gc) {
          io.out_gc[pt * 3 + (q - 1)] = q == 1 ? o[1] : (q == 2 ? o[2] : o[3]);
        }
      }
    } else if (BWD && prog.post_op == POST_BWD_DUMP) {
      // adjoint of the first layer's pre-activation: feeds no MMA of this launch, goes out as 4 dump-only chunks
      bool waited = false;
#pragma unroll 1
      for (int blk = 0; blk < 4; ++blk, ++c.ac) {
        Frag F, HG;
        uint32_t graw[8];
        const size_t go = static_cast<size_t>(prog.post_gate_base + blk) * CHUNK_PLANE_BYTES + frag_off;
        if (gate_tile_lo) load_planes_frag(gate_tile + go, gate_tile_lo + go, HG, prog.post_bwd_act == ACT_RELU ? 1 : 4);
        else fetch_gate_raw(gate_tile + go, prog.post_bwd_act == ACT_RELU ? 1 : 4, graw);
        if (!waited) {
          wait_d_full(c, c.g - 1);
          waited = true;
        }
        load_frag(c.tmem + ((c.g - 1) & 1) * HID + 64 * blk, F);
        if (gate_tile_lo) bwd_gate_frag(HG, prog.post_bwd_act, F);
        else bwd_gate_frag_raw(graw, prog.post_bwd_act, F);
        const uint32_t slot = claim_slot<true>(c, blk);
        emit_frag(F, c.sm + SM_A_OFF + slot * SLOT_BYTES + frag_off);
        publish_chunk(c, slot, 950 + blk);
      }
      release_d(c, c.g - 1);
    } else if (!BWD && prog.post_op == POST_FEAT_OUT) {
      wait_d_full(c, c.g - 1);
#pragma unroll 1
      for (int blk = 0; blk < 4; ++blk) {
        // feat = D + bias (row MAXL of the bias table, no activation); only the primal stream is needed
        Frag F;
        load_frag(c.tmem + ((c.g - 1) & 1) * HID + 64 * blk, F);
        const float2 b0 = lds64f(bias_sa0 + (MAXL * HID + 64 * blk) * 4);
        const float2 b1 = lds64f(bias_sa0 + (MAXL * HID + 64 * blk) * 4 + 32);
        const float f0 = F.f[0][0] + b0.x, f1 = F.f[0][1] + b0.y, f2 = F.f[0][2] + b1.x, f3 = F.f[0][3] + b1.y;
        if (valid && io.out_feat) {
          float* o = io.out_feat + pt * HID + 64 * blk + colq;
          *reinterpret_cast<float2*>(o) = make_float2(f0, f1);
          *reinterpret_cast<float2*>(o + 8) = make_float2(f2, f3);
        }
        if (valid && io.out_feat_rec) {
          // colour tile pt / 128, row pt % 128; columns (16 part + 2q, +1) sit in k-group 2 part at byte 4q, columns
          // (16 part + 8 + 2q, +1) in k-group 2 part + 1: a warp stores 8 rows x 16 B = one full line per instruction
          uint8_t* rec = io.out_feat_rec + (static_cast<size_t>(pt / TILE_ROWS) * 4 + blk) * (2 * CHUNK_PLANE_BYTES) +
                         2 * c.part * A_LBO + (pt % TILE_ROWS) * 16 + 4 * q;
          uint32_t h0, l0, h1, l1;
          split2(f0, f1, h0, l0);
          split2(f2, f3, h1, l1);
          *reinterpret_cast<uint32_t*>(rec) = h0;
          *reinterpret_cast<uint32_t*>(rec + A_LBO) = h1;
          *reinterpret_cast<uint32_t*>(rec + CHUNK_PLANE_BYTES) = l0;
          *reinterpret_cast<uint32_t*>(rec + CHUNK_PLANE_BYTES + A_LBO) = l1;
        }
      }
      release_d(c, c.g - 1);
      TRACE_EPI(8000 + 99);  // EPI: feature rows stored
    } else if (BWD && prog.post_op == POST_INADJ_SDF) {
      // adjoint of the enc6(x_c) rows (accumulator columns 0..63 in the kernel's chunk order; primal + tangent rows)
      // pushed back to x_c.  Row form: this thread is TMEM lane 32Q + lane = stream lane>>3 of point lane&7.
      wait_d_full(c, c.g - 1);
      const int s_row = c.lane >> 3;
      const long long pr = tile * TILE_PTS_T + 8 * c.quad + (c.lane & 7);
      const bool vr = pr < io.n_points;
      const long long ptr_ = vr ? pr : io.n_points - 1;
      float xb[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) xb[i] = 0.f;
      inadj_chunk(c.tmem + ((c.g - 1) & 1) * HID, SRC_ENC_SDF, c.part, __ldg(io.x_c + ptr_ * 3),
                  __ldg(io.x_c + ptr_ * 3 + 1), __ldg(io.x_c + ptr_ * 3 + 2), 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, s_row, xb);
      release_d(c, c.g - 1);
      // sum over the 4 stream rows of the point (lanes p, p + 8, p + 16, p + 24), then over the column parts
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        xb[i] += __shfl_xor_sync(0xffffffffu, xb[i], 8);
        xb[i] += __shfl_xor_sync(0xffffffffu, xb[i], 16);
      }
      const float4 sx = cross_part_sum_row(c, 32 * c.quad + c.lane, make_float4(xb[0], xb[1], xb[2], 0.f));
      if (writer && vr && s_row == 0 && io.adj_deform) {
        const float inv = 1.f / scale;
        float4* ad = reinterpret_cast<float4*>(io.adj_deform) + ptr_ * 4;
        ad[0].x += sx.x * inv;
        ad[0].y += sx.y * inv;
        ad[0].z += sx.z * inv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
// dump-only chunk groups of a layer / of the tile end (see ChainProg::pre_dump / post_dump)
template <bool BWD, bool STASH>
__device__ __forceinline__ int pre_dump_chunks(const LayerProg& L) {
  return (!BWD && STASH && L.pre_op == PRE_DEFORM_TAIL) ? 4 : 0;
}
template <bool BWD, bool STASH>
__device__ __forceinline__ int post_dump_chunks(const ChainProg& prog) {
  if (BWD) return prog.post_op == POST_BWD_DUMP ? 4 : 0;
  return (STASH && prog.post_op == POST_COLOR_TAIL) ? 4 : 0;
}

template <int CHAIN, bool TANGENT, bool BWD, bool STASH, bool PAIR>
__global__ void __launch_bounds__((BWD || STASH) ? N_THREADS_DUMP : N_THREADS, 1)
mlp_chain_kernel(const __grid_constant__ ChainProg prog, const __grid_constant__ ChainIO io) {
  constexpr bool DUMP = BWD || STASH;
  constexpr int NT = DUMP ? N_THREADS_DUMP : N_THREADS;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t sm = smem_u32(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM_OFF);
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;  // 0 = leader: issues the MMAs of the pair

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(reinterpret_cast<uint64_t*>(smem + BAR_A_FULL) + i, N_EPI_WARPS);  // one elected arrive per epilogue warp
      // slot free = the MMAs that read it have completed (+ the plane-dump warp has read it, training launches)
      mbar_init(reinterpret_cast<uint64_t*>(smem + BAR_A_EMPTY) + i, DUMP ? 2 : 1);
      mbar_init(reinterpret_cast<uint64_t*>(smem + BAR_A_FULL_PEER) + i, 1);
    }
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(reinterpret_cast<uint64_t*>(smem + BAR_W_FULL) + i, 1);
      mbar_init(reinterpret_cast<uint64_t*>(smem + BAR_W_EMPTY) + i, 1);
      mbar_init(reinterpret_cast<uint64_t*>(smem + BAR_W_FULL_PEER) + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(reinterpret_cast<uint64_t*>(smem + BAR_D_FULL) + i, 1);
      mbar_init(reinterpret_cast<uint64_t*>(smem + BAR_D_EMPTY) + i, N_EPI_WARPS);
      mbar_init(reinterpret_cast<uint64_t*>(smem + BAR_D_EMPTY_PEER) + i, 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc2<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  if constexpr (!BWD) {  // stage every layer's bias (+ the feature-layer bias in row MAXL) in shared memory
    float* bs = reinterpret_cast<float*>(smem + SM_BIAS_OFF);
    for (int i = threadIdx.x; i < prog.n_layers * HID; i += NT) bs[i] = __ldg(prog.bias + i);
    for (int i = threadIdx.x; i < HID; i += NT) bs[MAXL * HID + i] = __ldg(prog.feat_out_b + i);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // both CTAs' barriers and TMEM exist before anything crosses the pair
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int rows_per_tile = TANGENT ? TILE_PTS_T : TILE_ROWS;  // points per tile
  const long long n_tiles = (io.n_points + rows_per_tile - 1) / rows_per_tile;
  const TileWalk tw = make_walk<PAIR>(n_tiles);
  int* err = io.err;

  if (warp < EPI_WARP0) {
    if (warp == 0 && lane == 0) {
      // ============================================================== TMA producer
      // pair: each CTA streams ITS half of every unit (128 of the 256 weight rows, packed contiguously)
      // one ring stage = one 32-wide K sub-block: its hi unit and (3-term mode) its lo unit, adjacent in the stream
      uint32_t wc = 0;
      const bool three = prog.n_terms == 3;
      const uint32_t ubytes = static_cast<uint32_t>(prog.unit_bytes);
      const uint32_t cbytes = PAIR ? ubytes / 2 : ubytes;
      const uint8_t* src0 = prog.w_units + (PAIR ? rank * cbytes : 0);
      const bool w_hint = (io.store_hint & 4) != 0;  // weight units with an L2 evict_last policy
      const uint64_t w_pol = l2_policy_evict_last();
      for (long long it = 0; it < tw.n_iter; ++it) {
        for (int u = 0; u < prog.units_per_tile; u += 2) {
          const uint32_t st = wc % NSTAGE;
          mbar_wait_sa(sm + BAR_W_EMPTY + 8 * st, ((wc / NSTAGE) & 1) ^ 1, err, 200);
          uint64_t* full = reinterpret_cast<uint64_t*>(smem + BAR_W_FULL) + st;
          if (ES_FLAG(io, 1) && wc >= NSTAGE) {
            mbar_arrive(full);
          } else {
            mbar_arrive_expect_tx(full, three ? 2 * cbytes : cbytes);
            uint8_t* dst = smem + SM_W_OFF + st * STAGE_BYTES;
            if (w_hint) {
              tma_bulk_g2s_hint(dst, src0 + static_cast<size_t>(u) * ubytes, cbytes, full, w_pol);
              if (three)
                tma_bulk_g2s_hint(dst + UNIT_BYTES, src0 + static_cast<size_t>(u + 1) * ubytes, cbytes, full, w_pol);
            } else {
              tma_bulk_g2s(dst, src0 + static_cast<size_t>(u) * ubytes, cbytes, full);
              if (three) tma_bulk_g2s(dst + UNIT_BYTES, src0 + static_cast<size_t>(u + 1) * ubytes, cbytes, full);
            }
          }
          ++wc;
        }
      }
    } else if (warp == 1 && rank == 0) {
      // ============================================================== MMA issuer
      // The tensor pipe retires one M128 N256 K16 UMMA every 128 cycles, and this warp shares its SM sub-partition's
      // issue slots with four busy epilogue warps, so the loop must cost only a few instructions per MMA
      // (tools/mma_noise.py: an 18-instruction issue loop falls from 181 to 260 cycles/MMA under epilogue-like load).
      // The WHOLE warp walks the loop, so every value is warp-uniform and lives in uniform registers - no
      // elect/R2UR broadcast sequence in front of each UTCHMMA - and one elected lane issues.  Ring positions are
      // wrapped counters, not divisions.
      // Pair: the leader's warp issues M = 256 MMAs for both CTAs; what the peer has ready arrives through the *_PEER
      // barriers (relayed by the peer's otherwise idle warp 1), completions go to both CTAs by multicast commits.
      const uint32_t idesc = make_idesc_f16(PAIR ? 2 * TILE_ROWS : TILE_ROWS, prog.n_mma);
#ifdef ES_TRACE
      unsigned tcount = 0;
#endif
      // descriptors as low words (address field in 16-byte units + LBO); the high words are compile-time constants
      constexpr uint32_t A_HI32 = smem_desc_hi(A_SBO), W_HI32 = smem_desc_hi(B_SBO);
      // bytes between K core matrices of a weight unit = 16 x the weight rows this CTA holds
      const uint32_t b_lbo = static_cast<uint32_t>(PAIR ? prog.n_mma / 2 : prog.n_mma) * 16;
      const uint32_t a_desc0 = smem_desc_lo(sm + SM_A_OFF, A_LBO);
      const uint32_t w_desc0 = smem_desc_lo(sm + SM_W_OFF, b_lbo);
      constexpr uint32_t A_KS = (2 * A_LBO) >> 4;          // one K=16 step inside a slot plane
      constexpr uint32_t A_LO = SLOT_HALF_BYTES >> 4;      // hi plane -> lo plane
      constexpr uint32_t A_SB = (4 * A_LBO) >> 4;          // one 32-wide sub-block
      const uint32_t W_KS = (2 * b_lbo) >> 4;
      static_assert(((SM_W_OFF + NSTAGE * STAGE_BYTES) >> 4) < 0x4000, "descriptor address field");
      const bool three = prog.n_terms == 3;
      const bool do_mma = !ES_FLAG(io, 4);
      const bool leader = elect_one_sync();
      uint32_t st = 0, w_par = 0;      // weight ring stage and its phase parity
      uint32_t slot = 0, a_par = 0;    // A ring slot and its phase parity
      uint32_t g = 0;                  // global layer counter (accumulator buffer g & 1)
      auto mma = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t acc) {
        if constexpr (PAIR) umma2_f16_ss_lo<A_HI32, W_HI32>(d, a, b, idesc, acc);
        else umma_f16_ss_lo<A_HI32, W_HI32>(d, a, b, idesc, acc);
      };
      auto commit = [&](uint32_t bar_sa) {
        if constexpr (PAIR) umma2_commit_sa(bar_sa);
        else umma_commit_sa(bar_sa);
      };
      // dump-only chunks (training): nothing to multiply, only the slot hand-shake
      auto skip_chunks = [&](int n) {
        for (int i = 0; i < n; ++i) {
          mbar_wait_sa(sm + BAR_A_FULL + 8 * slot, a_par, err, 311);
          // (the peer relays these chunks too, so that the phases of A_FULL_PEER stay one per ring pass)
          if constexpr (PAIR) mbar_wait_sa(sm + BAR_A_FULL_PEER + 8 * slot, a_par, err, 313);
          if (leader) mbar_arrive_sa(sm + BAR_A_EMPTY + 8 * slot);
          if (++slot == NSLOT) { slot = 0; a_par ^= 1; }
        }
      };
      for (long long it = 0; it < tw.n_iter; ++it) {
        for (int l = 0; l < prog.n_layers; ++l, ++g) {
          const LayerProg& L = prog.layer[l];
          const int n_chunks = L.n_chunks;
          if constexpr (DUMP) skip_chunks(pre_dump_chunks<BWD, STASH>(L));
          const uint32_t d_tmem = tmem_base + (g & 1) * HID;
          mbar_wait_sa(sm + BAR_D_EMPTY + 8 * (g & 1), ((g >> 1) & 1) ^ 1, err, 300);
          if constexpr (PAIR) mbar_wait_sa(sm + BAR_D_EMPTY_PEER + 8 * (g & 1), (g >> 1) & 1, err, 301);
          tc_fence_after();
          TRACE_MMA(1000 + l);  // MMA: accumulator free, layer l starts
          uint32_t accum = 0;
          for (int ck = 0; ck < n_chunks; ++ck) {
            const int nsub = L.nsub[ck];
            mbar_wait_sa(sm + BAR_A_FULL + 8 * slot, a_par, err, 310);
            if constexpr (PAIR) mbar_wait_sa(sm + BAR_A_FULL_PEER + 8 * slot, a_par, err, 312);
            tc_fence_after();
            TRACE_MMA(2000 + l * 16 + ck);  // MMA: chunk ck of layer l available
            uint32_t a_hi = a_desc0 + slot * (SLOT_BYTES >> 4);
            for (int sb = 0; sb < nsub; ++sb, a_hi += A_SB) {
              // ---- one stage per sub-block: hi unit (A_hi*B_hi, A_lo*B_hi) and lo unit (A_hi*B_lo)
              mbar_wait_sa(sm + BAR_W_FULL + 8 * st, w_par, err, 320);
              if constexpr (PAIR) mbar_wait_sa(sm + BAR_W_FULL_PEER + 8 * st, w_par, err, 322);
              tc_fence_after();
              const uint32_t wd = w_desc0 + st * (STAGE_BYTES >> 4);
              if (leader) {
                if (do_mma) {
                  mma(d_tmem, a_hi, wd, accum);
                  if (three) mma(d_tmem, a_hi + A_LO, wd, 1);
                  mma(d_tmem, a_hi + A_KS, wd + W_KS, 1);
                  if (three) {
                    mma(d_tmem, a_hi + A_LO + A_KS, wd + W_KS, 1);
                    mma(d_tmem, a_hi, wd + (UNIT_BYTES >> 4), 1);
                    mma(d_tmem, a_hi + A_KS, wd + (UNIT_BYTES >> 4) + W_KS, 1);
                  }
                }
                commit(sm + BAR_W_EMPTY + 8 * st);
              }
              accum = 1;
              if (++st == NSTAGE) { st = 0; w_par ^= 1; }
            }
            if (leader) commit(sm + BAR_A_EMPTY + 8 * slot);
            if (++slot == NSLOT) { slot = 0; a_par ^= 1; }
          }
          if (leader) commit(sm + BAR_D_FULL + 8 * (g & 1));
          TRACE_MMA(3000 + l);  // MMA: all MMAs of layer l issued
        }
        if constexpr (DUMP) skip_chunks(post_dump_chunks<BWD, STASH>(prog));
      }
      __syncwarp();
    } else if (PAIR && warp == 1) {
      // ============================================================== relay (peer CTA of a pair)
      // Walks the waits of the leader's MMA warp in the same order and forwards this CTA's side of each of them
      // (accumulator released, A chunk published, weight half landed) to the leader's *_PEER barriers.
      if (lane == 0) {
        uint32_t st = 0, w_par = 0, slot = 0, a_par = 0, g = 0;
        const uint32_t r_a = mapa_u32(sm + BAR_A_FULL_PEER, 0), r_w = mapa_u32(sm + BAR_W_FULL_PEER, 0),
                       r_d = mapa_u32(sm + BAR_D_EMPTY_PEER, 0);
        auto skip_chunks = [&](int n) {
          for (int i = 0; i < n; ++i) {
            mbar_wait_sa(sm + BAR_A_FULL + 8 * slot, a_par, err, 411);
            mbar_arrive_cluster(r_a + 8 * slot);
            mbar_arrive_sa(sm + BAR_A_EMPTY + 8 * slot);
            if (++slot == NSLOT) { slot = 0; a_par ^= 1; }
          }
        };
        for (long long it = 0; it < tw.n_iter; ++it) {
          for (int l = 0; l < prog.n_layers; ++l, ++g) {
            const LayerProg& L = prog.layer[l];
            if constexpr (DUMP) skip_chunks(pre_dump_chunks<BWD, STASH>(L));
            mbar_wait_sa(sm + BAR_D_EMPTY + 8 * (g & 1), ((g >> 1) & 1) ^ 1, err, 400);
            mbar_arrive_cluster(r_d + 8 * (g & 1));
            for (int ck = 0; ck < L.n_chunks; ++ck) {
              mbar_wait_sa(sm + BAR_A_FULL + 8 * slot, a_par, err, 410);
              mbar_arrive_cluster(r_a + 8 * slot);
              if (++slot == NSLOT) { slot = 0; a_par ^= 1; }
              for (int u = 0; u < L.nsub[ck]; ++u) {
                mbar_wait_sa(sm + BAR_W_FULL + 8 * st, w_par, err, 420);
                mbar_arrive_cluster(r_w + 8 * st);
                if (++st == NSTAGE) { st = 0; w_par ^= 1; }
              }
            }
          }
          if constexpr (DUMP) skip_chunks(post_dump_chunks<BWD, STASH>(prog));
        }
      }
      __syncwarp();
    }
  } else if (DUMP && warp == STORE_WARP) {
    // ============================================================== plane-dump warp (training launches)
    // Walks the same chunk sequence as the epilogue and the MMA warp.  A chunk that is kept leaves as one 16 KiB bulk
    // copy per half straight out of the ring slot: no epilogue instructions, full-line HBM writes, and the record in
    // global memory has the ring-slot layout (see LayerProg::dump).
    if (lane == 0) {
      uint32_t slot = 0, a_par = 0;
      const bool hint = (io.store_hint & 1) || ((io.store_hint & 2) && !BWD);  // bit 1: forward (stash) launches only
      const uint64_t pol = l2_policy_evict_first();
      auto store = [&](uint8_t* dst, uint32_t src_sa) {
        if (hint) tma_bulk_s2g_hint(dst, src_sa, CHUNK_PLANE_BYTES, pol);
        else tma_bulk_s2g(dst, src_sa, CHUNK_PLANE_BYTES);
      };
      auto handle = [&](long long tile, int idx_hi, int idx_lo) {
        mbar_wait_sa(sm + BAR_A_FULL + 8 * slot, a_par, err, 500);
        if (ES_FLAG(io, 32) && tile < tw.n_tiles) tile = blockIdx.x;  // ablation: records stay L2 resident (148 tiles)
        const uint32_t slot_sa = sm + SM_A_OFF + slot * SLOT_BYTES;
        bool any = false;
        if (tile < tw.n_tiles && !ES_FLAG(io, 16)) {  // (a pair's ghost tile keeps nothing)
          if (idx_hi != NO_DUMP && io.dump_hi) {
            store(io.dump_hi + (static_cast<size_t>(tile) * prog.n_dump + idx_hi) * CHUNK_PLANE_BYTES, slot_sa);
            any = true;
          }
          if (idx_lo != NO_DUMP && io.dump_lo) {
            store(io.dump_lo + (static_cast<size_t>(tile) * prog.n_dump_lo + idx_lo) * CHUNK_PLANE_BYTES,
                  slot_sa + SLOT_HALF_BYTES);
            any = true;
          }
        }
        if (any) {
          tma_bulk_commit();
          tma_bulk_wait_read0();
        }
        mbar_arrive_sa(sm + BAR_A_EMPTY + 8 * slot);
        if (++slot == NSLOT) { slot = 0; a_par ^= 1; }
      };
      long long tile = tw.tile0;
      for (long long it = 0; it < tw.n_iter; ++it, tile += tw.step) {
        for (int l = 0; l < prog.n_layers; ++l) {
          const LayerProg& L = prog.layer[l];
          const int npre = pre_dump_chunks<BWD, STASH>(L);
          for (int i = 0; i < npre; ++i)
            handle(tile, prog.pre_dump == NO_DUMP ? NO_DUMP : prog.pre_dump + i,
                   prog.pre_dump_lo == NO_DUMP ? NO_DUMP : prog.pre_dump_lo + i);
          for (int ck = 0; ck < L.n_chunks; ++ck) handle(tile, L.dump[ck], L.dump_lo[ck]);
        }
        const int npost = post_dump_chunks<BWD, STASH>(prog);
        for (int i = 0; i < npost; ++i)
          handle(tile, prog.post_dump == NO_DUMP ? NO_DUMP : prog.post_dump + i,
                 prog.post_dump_lo == NO_DUMP ? NO_DUMP : prog.post_dump_lo + i);
      }
      tma_bulk_wait0();
    }
  } else {
    // ============================================================== epilogue warps
    Epi c;
    c.sm = sm;
    c.err = err;
    c.lane = lane;
    c.quad = warp & 3;
    c.part = (warp - EPI_WARP0) >> 2;
    c.tmem = tmem_base + (static_cast<uint32_t>(32 * c.quad) << 16) + PCOLS * c.part;
    c.row = TANGENT ? 32 * c.quad + 8 * (lane & 3) + (lane >> 2) : 32 * c.quad + lane;
    c.ac = 0;
    c.g = 0;
#ifdef ES_TRACE
    c.trace = io.trace;
    c.tr = (io.trace != nullptr && warp == EPI_WARP0 && lane == 0 && blockIdx.x == 0);
    c.tcount = 0;
#endif
    if constexpr (TANGENT) epilogue_tangent<BWD, STASH>(prog, io, c, tw);
    else epilogue_plain<CHAIN, BWD, STASH>(prog, io, c, tw);
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other still depends on it
  else __syncthreads();
  if (warp == 1) {
    if constexpr (PAIR) tmem_dealloc2<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ launchers
template <int CHAIN, bool TANGENT, bool BWD, bool STASH, bool PAIR>
static cudaError_t launch_impl(const ChainProg& prog, const ChainIO& io, int n_sms, cudaStream_t stream) {
  auto kern = mlp_chain_kernel<CHAIN, TANGENT, BWD, STASH, PAIR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
  if (e != cudaSuccess) return e;
  const int pts_per_tile = TANGENT ? TILE_PTS_T : TILE_ROWS;
  long long n_tiles = (io.n_points + pts_per_tile - 1) / pts_per_tile;
  if (n_tiles <= 0) return cudaSuccess;
  const int threads = (BWD || STASH) ? N_THREADS_DUMP : N_THREADS;
  if constexpr (PAIR) {
    const long long n_pairs = (n_tiles + 1) / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * (n_sms / 2));
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = SM_TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent grid = the clusters that are resident at once: a pair needs both SMs of a TPC, and a part with
    // harvested SMs has fewer complete TPCs than n_sms / 2 (a second wave would double the kernel time)
    static int max_clusters = -1;
    if (max_clusters < 0) {
      int mc = 0;
      if (cudaOccupancyMaxActiveClusters(&mc, kern, &cfg) != cudaSuccess || mc <= 0) mc = n_sms / 2;
      max_clusters = mc < n_sms / 2 ? mc : n_sms / 2;
      (void)cudaGetLastError();
    }
    const int clusters = static_cast<int>(n_pairs < max_clusters ? n_pairs : max_clusters);
    cfg.gridDim = dim3(2 * clusters);
    return cudaLaunchKernelEx(&cfg, kern, prog, io);
  } else {
    int grid = static_cast<int>(n_tiles < n_sms ? n_tiles : n_sms);
    kern<<<grid, threads, SM_TOTAL, stream>>>(prog, io);
    return cudaGetLastError();
  }
}
template <int CHAIN, bool TANGENT, bool BWD, bool STASH>
static cudaError_t launch_one(const ChainProg& prog, const ChainIO& io, int n_sms, cudaStream_t stream, bool pair) {
  return pair ? launch_impl<CHAIN, TANGENT, BWD, STASH, true>(prog, io, n_sms, stream)
              : launch_impl<CHAIN, TANGENT, BWD, STASH, false>(prog, io, n_sms, stream);
}

cudaError_t launch_mlp_chain(int chain, bool tangent, bool pair, const ChainProg& prog, const ChainIO& io,
                             int n_sms, cudaStream_t stream, bool bwd) {
  // pair: CTA pairs (cta_group::2); the weight units must have been packed in the pair layout (es_pack.cu)
  const bool stash = !bwd && io.dump_hi != nullptr;
  if (prog.n_mma < 16 || prog.n_mma > HID || prog.n_mma % 16 != 0 || prog.unit_bytes != prog.n_mma * SUB_K * 2)
    return cudaErrorInvalidValue;
  if (pair && prog.n_mma != HID) return cudaErrorInvalidValue;
  if (chain == CHAIN_COLOR) {
    if (bwd) return launch_one<CHAIN_COLOR, false, true, false>(prog, io, n_sms, stream, pair);
    return stash ? launch_one<CHAIN_COLOR, false, false, true>(prog, io, n_sms, stream, pair)
                 : launch_one<CHAIN_COLOR, false, false, false>(prog, io, n_sms, stream, pair);
  }
  if (chain == CHAIN_SDF) {
    if (bwd)
      return tangent ? launch_one<CHAIN_SDF, true, true, false>(prog, io, n_sms, stream, pair) : cudaErrorInvalidValue;
    if (!tangent) return launch_one<CHAIN_SDF, false, false, false>(prog, io, n_sms, stream, pair);
    return stash ? launch_one<CHAIN_SDF, true, false, true>(prog, io, n_sms, stream, pair)
                 : launch_one<CHAIN_SDF, true, false, false>(prog, io, n_sms, stream, pair);
  }
  return cudaErrorInvalidValue;
}

int mlp_chain_smem_bytes() { return SM_TOTAL; }

}  // namespace es
