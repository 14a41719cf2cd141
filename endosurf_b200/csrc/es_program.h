// Layer "program" shared by host (es_api.cu) and device (es_mlp.cu): how one 128-row tile walks through a chain of
// 256-wide MLP layers on the tensor cores.  Also the single source of truth for the K ordering of the
// positional-encoding chunks (the host packs weight columns in exactly this order).
#pragma once
#include <stdint.h>

namespace es {

constexpr int HID = 256;           // hidden width every MMA layer is padded to (N of the UMMA)
constexpr int TILE_ROWS = 128;     // UMMA M
constexpr int CHUNK_K = 64;        // K extent of one A-operand ring slot
constexpr int SUB_K = 32;          // K extent of one weight unit
constexpr int SLOT_HALF_BYTES = TILE_ROWS * CHUNK_K * 2;  // 16 KiB: hi or lo plane of a slot
constexpr int SLOT_BYTES = 2 * SLOT_HALF_BYTES;           // 32 KiB
#ifndef ES_NSLOT
#define ES_NSLOT 3
#endif
#ifndef ES_NSTAGE
#define ES_NSTAGE 3
#endif
constexpr int NSLOT = ES_NSLOT;     // A-operand ring slots (32 KiB each)
constexpr int UNIT_BYTES = HID * SUB_K * 2;               // 16 KiB: 256 x 32 fp16 (hi or lo)
constexpr int STAGE_BYTES = 2 * UNIT_BYTES;                // one weight ring stage: the hi and the lo unit of a 32-wide K sub-block
constexpr int NSTAGE = ES_NSTAGE;  // weight ring stages (32 KiB each): must cover L2 latency x 43 B/clk (>= 96 KiB)
// canonical no-swizzle K-major layout: [k-group of 8][row][8 elements]
constexpr int A_LBO = TILE_ROWS * 16;  // 2048  bytes between K core matrices
constexpr int A_SBO = 128;             //        bytes between 8-row groups
constexpr int B_LBO = HID * 16;        // 4096
constexpr int B_SBO = 128;

constexpr int MAXL = 20;  // MMA layers in one chain
constexpr int MAXC = 10;  // K chunks in one layer

// chunk sources (who fills an A-operand ring slot)
enum : uint8_t {
  SRC_PREV = 0,     // act(D[prev layer][:, 64*arg .. +64) + bias)
  SRC_ENC_DEFORM,   // enc6(x) (+) enc6(t), 52 used of 64
  SRC_ENC_SDF,      // enc6(x_c), 39 used of 64
  SRC_COLOR_A,      // colour-net input, first 64 of [enc10(x_c), g_c, enc4(d_c)] in kernel order
  SRC_COLOR_B,      // remaining 30 (one 32-wide sub-block)
  SRC_FEAT,         // geo_feat[:, 64*arg .. +64) read from global memory
  // reverse (training) chains: the "activations" are adjoints, the weights are packed transposed
  SRC_ADJ_FEAT,     // d loss / d feat [P,256] (primal rows; tangent rows are zero)
  SRC_BWD_PREV,     // activation-backward of the previous reverse layer's accumulator using the forward stash
  SRC_BWD_OUTER3,   // same, but the incoming adjoint is adj.xyz . W_out (the 3-wide output layer, no MMA)
  SRC_PLANE,        // copy of a 64-column fp16 hi/lo plane chunk written by an earlier launch (arg = chunk index)
};
enum : uint8_t { ACT_NONE = 0, ACT_RELU = 1, ACT_SOFTPLUS100 = 2 };
enum : uint8_t { PRE_NONE = 0, PRE_DEFORM_TAIL = 1 };
enum : uint8_t {
  POST_NONE = 0, POST_SDF_TAIL = 1, POST_FEAT_OUT = 2, POST_COLOR_TAIL = 3, POST_BWD_DUMP = 4,
  // input-adjoint launches of the training backward (one MMA layer whose chunks are SRC_PLANE copies of zbar planes)
  POST_INADJ_SDF = 5,    // adjoint of enc6(x_c) rows (primal + tangent)      -> d loss / d x_c
  POST_INADJ_COLOR = 6,  // adjoint of [enc10(x_c), g_c, enc4(d_c)]           -> d loss / d (x_c, g_c, J)
  POST_FEAT_BAR = 7,     // adjoint of the geometry feature (colour-net input) -> feat_bar [P,256] fp32
};
constexpr uint8_t NO_DUMP = 0xFF;

struct LayerProg {
  uint8_t n_chunks;
  uint8_t act;       // activation of THIS layer's output (applied by the consumer of its accumulator)
  uint8_t pre_op;    // epilogue work before this layer's first chunk
  uint8_t side_dot;  // 1: while converting the previous layer's output also accumulate the sdf row
  uint8_t last_prev; // index of the last chunk that reads the previous accumulator (SRC_PREV / SRC_BWD_PREV), 0xFF: none
  uint8_t src[MAXC];
  uint8_t arg[MAXC];
  uint8_t nsub[MAXC];  // number of 32-wide K sub-blocks in the chunk that carry weights (1 or 2)
  // reverse chains only
  uint8_t bwd_act;     // activation being differentiated (ACT_RELU / ACT_SOFTPLUS100)
  uint8_t rank1;       // 1: add adj.w * sdf_out_w[col] to the incoming adjoint (sdf row of the output layer)
  // training planes.  Every A-operand chunk a training launch builds in shared memory can be kept: a dedicated warp
  // bulk-copies the 16 KiB hi (and optionally lo) half of the ring slot to global memory as is, so a plane chunk in
  // HBM is [k-group of 8 columns][128 tile rows][8 fp16] - at once the K-major A operand of the chains (SRC_PLANE
  // reload) and the MN-major operand of the weight-gradient kernel (es_wgrad.cu), both by plain 1-D bulk copies.
  uint8_t dump[MAXC];     // per chunk: index of the chunk in this launch's hi plane record of a tile (NO_DUMP: none)
  uint8_t dump_lo[MAXC];  // same for the lo plane record (NO_DUMP: lo half not kept)
  uint8_t gate_base;      // reverse: chunk index (forward stash record) of column block 0 of the gating activations
};

// host helper: fill LayerProg::last_prev after the chunk list is complete
inline void finish_layer(LayerProg& g) {
  g.last_prev = 0xFF;
  for (int k = 0; k < g.n_chunks; ++k)
    if (g.src[k] == SRC_PREV || g.src[k] == SRC_BWD_PREV) g.last_prev = static_cast<uint8_t>(k);
}

struct ChainProg {
  int32_t n_layers;
  int32_t units_per_tile;  // 16 KiB weight units streamed per tile (hi and lo counted separately)
  int32_t n_terms;         // 3: fp16x3 split (fp32 parity); 1: single fp16 pass
  int32_t post_op;
  LayerProg layer[MAXL];
  const uint8_t* w_units;  // packed weight units, consumption order
  const float* bias;       // [n_layers][256]
  const float* deform_out_w;  // [3][256]  (deform last layer, fp32)
  const float* deform_out_b;  // [3]
  const float* sdf_out_w;     // [256]     (sdf last layer row 0)
  const float* sdf_out_b;     // [1]
  const float* feat_out_b;    // [256]     (sdf last layer rows 1..256 bias)
  const float* color_out_w;   // [3][256]
  const float* color_out_b;   // [3]
  const float* outer3_w;      // reverse chains: [3][256] output-layer weights for SRC_BWD_OUTER3
  // POST_BWD_DUMP: last activation-backward of a reverse chain (its result feeds no MMA, only the weight grads)
  int32_t post_gate_base, post_bwd_act;
  // dump-only chunk groups (4 chunks each, no MMA): the inputs of the 3-wide output layers in the forward training
  // chains (PRE_DEFORM_TAIL / POST_COLOR_TAIL) and the POST_BWD_DUMP result.  First chunk index, NO_DUMP = not kept.
  uint8_t pre_dump, pre_dump_lo, post_dump, post_dump_lo;
  int32_t n_dump, n_dump_lo;    // chunks per tile in this launch's hi / lo plane records
  int32_t n_gate;               // chunks per tile in the forward stash record read by the reverse gates
  int32_t gate_use_lo;          // gates read hi + lo (the lo record has the same indexing as the hi record)
  int32_t n_plane, n_plane_lo;  // chunks per tile of the SRC_PLANE source records (lo index = dump_lo of the source)
  uint8_t plane_lo[MAXC];       // SRC_PLANE: lo-record chunk index of chunk ck of layer 0 (NO_DUMP: lo = 0)
  // MMA shape of this launch (input-adjoint launches use narrower N)
  int32_t n_mma;                // UMMA N (256 for the chains)
  int32_t unit_bytes;           // bytes of one weight unit = n_mma * 32 * 2
};

// ------------------------------------------------------------------------------------------------------------------
// K ordering of the encoder chunks.  A feature is (var, freq, is_cos): var 0..2 = position x/y/z, 3 = time,
// 4..6 = canonical normal g_c, 7..9 = canonical view direction d_c; freq -1 = identity, else sin/cos(2^freq * v).
// var -1 = zero padding.
struct Feat {
  int8_t var, freq, is_cos;
};
__host__ __device__ constexpr Feat feat_pad() { return Feat{-1, -1, 0}; }
// identity of a 3-vector group + sin/cos blocks, laid out [id(3)] then per freq [sin(3), cos(3)]
__host__ __device__ constexpr Feat enc3_feat(int base_var, int idx /*0.. 3+6L*/) {
  if (idx < 3) return Feat{static_cast<int8_t>(base_var + idx), -1, 0};
  int b = (idx - 3) / 3, c = (idx - 3) % 3;
  return Feat{static_cast<int8_t>(base_var + c), static_cast<int8_t>(b / 2), static_cast<int8_t>(b % 2)};
}
__host__ __device__ constexpr Feat enc1_feat(int var, int idx /*0.. 1+2L*/) {
  if (idx < 1) return Feat{static_cast<int8_t>(var), -1, 0};
  int b = idx - 1;
  return Feat{static_cast<int8_t>(var), static_cast<int8_t>(b / 2), static_cast<int8_t>(b % 2)};
}
// chunk column -> feature
__host__ __device__ constexpr Feat chunk_feat(int src, int col) {
  if (src == SRC_ENC_DEFORM) {
    // half 0: x id + x f0..3 (27) + t id + t f0,1 (5)     half 1: x f4,5 (12) + t f2..5 (8) + pad 12
    if (col < 27) return enc3_feat(0, col);
    if (col < 32) return enc1_feat(3, col - 27);
    if (col < 44) return enc3_feat(0, 27 + (col - 32));
    if (col < 52) return enc1_feat(3, 5 + (col - 44));
    return feat_pad();
  }
  if (src == SRC_ENC_SDF) {
    // half 0: x id + f0..3 (27) + pad 5                   half 1: f4,5 (12) + pad 20
    if (col < 27) return enc3_feat(0, col);
    if (col < 32) return feat_pad();
    if (col < 44) return enc3_feat(0, 27 + (col - 32));
    return feat_pad();
  }
  if (src == SRC_COLOR_A) {
    // half 0: x id + f0..3 (27) + g_c (3) + d id[0,1] (2)  half 1: x f4..8 (30) + d id[2] + pad
    if (col < 27) return enc3_feat(0, col);
    if (col < 30) return Feat{static_cast<int8_t>(4 + (col - 27)), -1, 0};
    if (col < 32) return Feat{static_cast<int8_t>(7 + (col - 30)), -1, 0};
    if (col < 62) return enc3_feat(0, 27 + (col - 32));
    if (col < 63) return Feat{9, -1, 0};
    return feat_pad();
  }
  if (src == SRC_COLOR_B) {
    // x f9 (6) + d f0..3 (24) + pad 2   (only sub-block 0 exists)
    if (col < 6) return enc3_feat(0, 57 + col);
    if (col < 30) return enc3_feat(7, 3 + (col - 6));
    return feat_pad();
  }
  return feat_pad();
}

}  // namespace es
