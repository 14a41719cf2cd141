// C ABI of endosurf_b200 (include/endosurf_b200.h): context, weight packing, layer programs and the orchestration
// of render_rays.  Host-side only; all device work is in es_mlp.cu / es_rays.cu / es_pack.cu / es_probe.cu.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/endosurf_b200.h"
#include "es_kernels.h"
#include "es_program.h"

using namespace es;

namespace {

struct LayerPack {
  std::vector<int> colmap;  // K order of the packed layer -> reference weight column (-1 = zero)
  int* colmap_dev = nullptr;
  int k_total = 0;          // multiple of 32
  int n_out = 0, n_in = 0;  // reference weight shape (rows used from row_off)
  int row_off = 0;
  float scale = 1.f;
  long long unit_off = 0;   // first 16 KiB unit in the chain's weight stream
};

struct NetPlan {
  int in_dim = 0;                 // reference input width
  std::vector<int> out_dims;      // reference per-layer out dims
  std::vector<int> in_dims;       // reference per-layer in dims
  std::vector<LayerPack> packs;   // MMA layers of this net (hidden layers [+ feat layer for sdf])
  std::vector<LayerProg> progs;
};

}  // namespace

struct es_ctx {
  es_net_config cfg{};
  int device = 0;
  int n_sms = 0;
  std::string err;
  long long launches = 0;
  NetPlan plan[3];
  bool loaded[3] = {false, false, false};
  // device buffers
  uint8_t* geom_units = nullptr;   // deform hidden | sdf hidden | sdf feat
  long long geom_units_n = 0, sdfq_units_n = 0;
  uint8_t* color_units = nullptr;
  long long color_units_n = 0;
  float* geom_bias = nullptr;      // [(Ld + Ls + 1), 256]
  float* color_bias = nullptr;     // [Lc, 256]
  float* small = nullptr;          // deform_out_w[768] deform_out_b[4] sdf_out_w[256] sdf_out_b[4] feat_out_b[256]
                                   // color_out_w[768] color_out_b[4]
  int* err_dev = nullptr;
  int* err_host = nullptr;         // pinned mirror of err_dev, refreshed by es_poll_error
  // grow-only scratch
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  ChainProg prog_sdfq{}, prog_geom{}, prog_color{};
  // reverse (training) chains: transposed weight units + programs, index = ES_NET_*
  uint8_t* rev_units[3] = {nullptr, nullptr, nullptr};
  int rev_layers[3] = {0, 0, 0};
  ChainProg prog_rev[3]{};
  // input-adjoint launches: 0 = sdf enc rows (N 64), 1 = colour feature (N 256), 2 = colour enc/g_c/d_c (N 128)
  uint8_t* inadj_units[3] = {nullptr, nullptr, nullptr};
  int* inadj_cols_dev[3] = {nullptr, nullptr, nullptr};   // [2][n_mma] source columns of layer 0 / the skip layer
  ChainProg prog_inadj[3]{};
  // plane records of the training chains (chunks of 16 KiB per tile, es_program.h LayerProg::dump)
  struct FwdLayout {
    int n_dump = 0;
    int in_idx[MAXL][MAXC];   // record index of input chunk ck of program layer l (duplicates resolved)
    int tail = -1;            // first of the 4 dump-only chunks (deform tail / colour tail)
  } lay_geom, lay_color;
  // weight-norm fold targets: effective weights owned by the context (es_load_network_wn)
  float* w_eff[3][16] = {};
  // loss scales of the three reverse chains (device): amax bit patterns and the power-of-two scales
  unsigned int* amax_dev = nullptr;   // [3]
  float* scale_dev = nullptr;         // [3]
  float* eik_den_dev = nullptr;       // [1] scratch
  // cached weight-gradient work lists, keyed by (points, plane mode)
  struct WgradPlan {
    long long n = -1;
    int full = 0;
    WgradItem* items_dev = nullptr;
    int n_items = 0;
    std::vector<WgradJob> jobs;   // pointers patched per call
    std::vector<int> job_net;     // ES_NET_* of every job
    std::vector<int> job_layer;
    bool paired = false;          // launch order = (M half 0, M half 1) pairs: run as 2-CTA clusters
    int* colmaps_dev = nullptr;
  };
  std::vector<WgradPlan> wplans;
  // training planes: 0 (default) = fp16 hi planes only (1-term weight gradients, hi-only activation gates) plus the
  // lo halves the input-adjoint launches need; 1 = every lo plane (3-term weight gradients, exact gates)
  int full_planes = 0;
  int wgrad_lbo = 0, wgrad_sbo = 0;   // debug override of the MN-major descriptor strides (0: built-in)
  int pair_mode = 0;                  // standard 256-wide chains run on CTA pairs (cta_group::2)
  int wgrad_pairs = 1;                // weight-gradient CTAs of the two M halves share their B tile (cluster multicast)
  float scale_target = 1024.f;        // the largest adjoint entering a reverse chain is scaled to about this:
                                      // 64x below the fp16 maximum (conversions saturate), and small adjoints
                                      // stay above the fp16 subnormal floor (tools/diag_grad_precision.py)
  // optional per-kernel timing (es_profile_*)
  bool profiling = false;
  int debug_flags = 0;
  long long* trace_dev = nullptr;  // debug pipeline trace buffer (es_debug_trace)
  int trace_kind = -1;
  int store_hint = 1;              // L2 policies (ChainIO::store_hint): plane-record stores evict_first by default -
                                   // the 19 GB record stream of a training launch no longer evicts the packed weights
                                   // (geometry chain 12.7 -> 11.5 ms, profiles/r2_experiments.txt)
  int feat_records = 1;            // geometry feature travels to the colour chain as fp16 plane records (0: fp32 rows)
  struct Timed {
    int kind;
    long long points;
    cudaEvent_t e0, e1;
  };
  std::vector<Timed> timed;
};

namespace {

constexpr int SM_DEFORM_W = 0, SM_DEFORM_B = 768, SM_SDF_W = 772, SM_SDF_B = 1028, SM_FEAT_B = 1032,
              SM_COLOR_W = 1288, SM_COLOR_B = 2056, SM_TOTAL_F = 2060;

int fail(es_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
int cuda_fail(es_ctx* c, cudaError_t e, const char* where) {
  if (c) c->err = std::string(where) + ": " + cudaGetErrorString(e);
  return static_cast<int>(e);
}
#define CU(call)                                                   \
  do {                                                             \
    cudaError_t e__ = (call);                                      \
    if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call);     \
  } while (0)

// reference input column of an encoder feature for each network (es_program.h Feat)
int ref_column(const es_net_config& cfg, int net, Feat f) {
  if (f.var < 0) return -1;
  auto enc3 = [](int base, int c, int freq, int is_cos) {
    return freq < 0 ? base + c : base + 3 + 6 * freq + 3 * is_cos + c;
  };
  if (net == ES_NET_DEFORM) {
    const int nx = 3 + 6 * cfg.multires_deform_pos;
    if (f.var < 3) return f.freq < cfg.multires_deform_pos ? enc3(0, f.var, f.freq, f.is_cos) : -1;
    if (f.var == 3) {
      if (f.freq >= cfg.multires_deform_time) return -1;
      return f.freq < 0 ? nx : nx + 1 + 2 * f.freq + f.is_cos;
    }
    return -1;
  }
  if (net == ES_NET_SDF) {
    if (f.var < 3) return f.freq < cfg.multires_sdf_pos ? enc3(0, f.var, f.freq, f.is_cos) : -1;
    return -1;
  }
  // colour: [enc(x_c) | g_c(3) | enc(d_c) | feat(256)]
  const int nx = 3 + 6 * cfg.multires_color_pos;
  if (f.var < 3) return f.freq < cfg.multires_color_pos ? enc3(0, f.var, f.freq, f.is_cos) : -1;
  if (f.var >= 4 && f.var <= 6) return nx + (f.var - 4);
  if (f.var >= 7 && f.var <= 9) return f.freq < cfg.multires_color_dir ? enc3(nx + 3, f.var - 7, f.freq, f.is_cos) : -1;
  return -1;
}

void chunk_cols(const es_net_config& cfg, int net, int src, int base, int width, std::vector<int>& out) {
  for (int k = 0; k < width; ++k) {
    int rc = ref_column(cfg, net, chunk_feat(src, k));
    out.push_back(rc < 0 ? -1 : base + rc);
  }
}

// Build the MMA-layer plan of one network.
void build_net_plan(const es_net_config& cfg, int net, NetPlan& P) {
  const int L = cfg.n_layers, H = cfg.hidden_dim, skip = cfg.skip_layer;
  const int nx_d = 3 + 6 * cfg.multires_deform_pos, nt_d = 1 + 2 * cfg.multires_deform_time;
  const int nx_s = 3 + 6 * cfg.multires_sdf_pos;
  const int nx_c = 3 + 6 * cfg.multires_color_pos, nd_c = 3 + 6 * cfg.multires_color_dir;
  P.in_dim = net == ES_NET_DEFORM ? nx_d + nt_d : net == ES_NET_SDF ? nx_s : nx_c + 3 + nd_c + H;
  const int final_out = net == ES_NET_SDF ? 1 + H : 3;
  P.out_dims.assign(L, H);
  P.in_dims.assign(L, H);
  for (int l = 0; l < L; ++l) {
    if (net == ES_NET_DEFORM) {  // build_mlp_idr (utils.py:63-111): the layer before a skip emits H - in_dim
      P.in_dims[l] = l == 0 ? P.in_dim : H;
      P.out_dims[l] = l == L - 1 ? final_out : (l + 1 == skip ? H - P.in_dim : H);
    } else {                     // build_mlp_nerf (utils.py:11-60): the skip layer takes H + in_dim
      P.in_dims[l] = l == 0 ? P.in_dim : (l == skip ? H + P.in_dim : H);
      P.out_dims[l] = l == L - 1 ? final_out : H;
    }
  }
  const uint8_t act = net == ES_NET_SDF ? ACT_SOFTPLUS100 : ACT_RELU;
  const int n_mma = (L - 1) + (net == ES_NET_SDF ? 1 : 0);
  P.packs.assign(n_mma, LayerPack());
  P.progs.assign(n_mma, LayerProg());
  for (int l = 0; l < n_mma; ++l) {
    LayerPack& K = P.packs[l];
    LayerProg& G = P.progs[l];
    std::memset(&G, 0, sizeof(G));
    K.n_in = P.in_dims[l];
    K.n_out = P.out_dims[l];
    K.row_off = 0;
    G.act = act;
    int nc = 0;
    auto add_chunk = [&](uint8_t src, uint8_t arg, uint8_t nsub) {
      G.src[nc] = src;
      G.arg[nc] = arg;
      G.nsub[nc] = nsub;
      ++nc;
    };
    auto add_input_chunks = [&](int base) {
      if (net == ES_NET_DEFORM) {
        add_chunk(SRC_ENC_DEFORM, 0, 2);
        chunk_cols(cfg, net, SRC_ENC_DEFORM, base, 64, K.colmap);
      } else if (net == ES_NET_SDF) {
        add_chunk(SRC_ENC_SDF, 0, 2);
        chunk_cols(cfg, net, SRC_ENC_SDF, base, 64, K.colmap);
      } else {
        add_chunk(SRC_COLOR_A, 0, 2);
        chunk_cols(cfg, net, SRC_COLOR_A, base, 64, K.colmap);
        add_chunk(SRC_COLOR_B, 0, 1);
        chunk_cols(cfg, net, SRC_COLOR_B, base, 32, K.colmap);
        for (int c = 0; c < 4; ++c) {
          add_chunk(SRC_FEAT, static_cast<uint8_t>(c), 2);
          for (int k = 0; k < 64; ++k) K.colmap.push_back(base + nx_c + 3 + nd_c + 64 * c + k);
        }
      }
    };
    auto add_prev_chunks = [&](int n_valid) {
      for (int c = 0; c < 4; ++c) {
        add_chunk(SRC_PREV, static_cast<uint8_t>(c), 2);
        for (int k = 0; k < 64; ++k) K.colmap.push_back(64 * c + k < n_valid ? 64 * c + k : -1);
      }
    };
    if (l == 0) {
      add_input_chunks(0);
    } else if (l == skip && l < L - 1) {
      // cat([h, input]) / sqrt(2): h occupies the first out_dims[l-1] columns
      add_prev_chunks(P.out_dims[l - 1]);
      add_input_chunks(P.out_dims[l - 1]);
      K.scale = static_cast<float>(1.0 / std::sqrt(2.0));
    } else {
      add_prev_chunks(P.out_dims[l - 1]);
    }
    if (net == ES_NET_SDF && l == L - 1) {  // feat rows 1..256 of the output layer; row 0 is the side dot
      K.row_off = 1;
      K.n_out = H;
      G.act = ACT_NONE;
      G.side_dot = 1;
    }
    G.n_chunks = static_cast<uint8_t>(nc);
    finish_layer(G);
    K.k_total = static_cast<int>(K.colmap.size());
  }
}

long long plan_units(const NetPlan& P, int n_layers_used) {
  long long u = 0;
  for (int l = 0; l < n_layers_used; ++l) u += 2 * (P.packs[l].k_total / SUB_K);
  return u;
}

int ensure_ws(es_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return 0;
  if (ctx->ws) {
    CU(cudaDeviceSynchronize());
    CU(cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
  }
  size_t want = bytes + bytes / 8;
  CU(cudaMalloc(&ctx->ws, want));
  ctx->ws_bytes = want;
  return 0;
}

struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(uint8_t* b) : base(b) {}
  template <class T>
  T* take(size_t n) {
    off = (off + 255) & ~static_cast<size_t>(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

void build_chain_programs(es_ctx* ctx) {
  const es_net_config& cfg = ctx->cfg;
  const NetPlan& D = ctx->plan[ES_NET_DEFORM];
  const NetPlan& S = ctx->plan[ES_NET_SDF];
  const NetPlan& C = ctx->plan[ES_NET_COLOR];
  const int Ld = cfg.use_deform ? static_cast<int>(D.progs.size()) : 0;
  const int Ls = static_cast<int>(S.progs.size());  // hidden + feat
  auto common = [&](ChainProg& p) {
    p.n_terms = cfg.precision_terms;
    p.n_mma = HID;
    p.unit_bytes = UNIT_BYTES;
    p.pre_dump = p.pre_dump_lo = p.post_dump = p.post_dump_lo = NO_DUMP;
    for (int l = 0; l < MAXL; ++l)
      for (int k = 0; k < MAXC; ++k) p.layer[l].dump[k] = p.layer[l].dump_lo[k] = NO_DUMP;
    for (int k = 0; k < MAXC; ++k) p.plane_lo[k] = NO_DUMP;
    p.deform_out_w = ctx->small + SM_DEFORM_W;
    p.deform_out_b = ctx->small + SM_DEFORM_B;
    p.sdf_out_w = ctx->small + SM_SDF_W;
    p.sdf_out_b = ctx->small + SM_SDF_B;
    p.feat_out_b = ctx->small + SM_FEAT_B;
    p.color_out_w = ctx->small + SM_COLOR_W;
    p.color_out_b = ctx->small + SM_COLOR_B;
  };
  // geometry chain (with feat layer) and its prefix, the sdf query chain
  ChainProg g{};
  common(g);
  int n = 0;
  for (int l = 0; l < Ld; ++l) g.layer[n++] = D.progs[l];
  for (int l = 0; l < Ls; ++l) {
    g.layer[n] = S.progs[l];
    if (l == 0 && Ld > 0) g.layer[n].pre_op = PRE_DEFORM_TAIL;
    ++n;
  }
  g.n_layers = n;
  g.w_units = ctx->geom_units;
  g.bias = ctx->geom_bias;
  g.units_per_tile = static_cast<int>(ctx->geom_units_n);
  g.post_op = POST_FEAT_OUT;
  ChainProg c{};
  common(c);
  c.n_layers = static_cast<int>(C.progs.size());
  for (int l = 0; l < c.n_layers; ++l) c.layer[l] = C.progs[l];
  c.w_units = ctx->color_units;
  c.bias = ctx->color_bias;
  c.units_per_tile = static_cast<int>(ctx->color_units_n);
  c.post_op = POST_COLOR_TAIL;

  // ---- plane records of the forward training chains: every input chunk of every layer is kept once (the skip
  // layers re-read the network input, whose chunks are already in the record from layer 0), plus the inputs of the
  // 3-wide output layers as 4 dump-only chunks
  const int skip = cfg.skip_layer;
  auto assign_forward = [&](ChainProg& p, es_ctx::FwdLayout& lay, int n_first /*layers of the first net*/) {
    int idx = 0;
    lay.tail = -1;
    for (int l = 0; l < p.n_layers; ++l) {
      LayerProg& G = p.layer[l];
      if (G.pre_op == PRE_DEFORM_TAIL) {
        lay.tail = idx;
        p.pre_dump = static_cast<uint8_t>(idx);
        idx += 4;
      }
      const int net_l = l < n_first ? l : l - n_first;   // layer index inside its network
      const int l0 = l < n_first ? 0 : n_first;          // program index of that network's layer 0
      int dup = 0;                                        // next input chunk of layer 0 this skip layer repeats
      for (int k = 0; k < MAXC; ++k) G.dump[k] = G.dump_lo[k] = NO_DUMP;
      for (int k = 0; k < G.n_chunks; ++k) {
        const bool is_input = G.src[k] != SRC_PREV;
        if (is_input && net_l == skip && net_l != 0) {
          lay.in_idx[l][k] = lay.in_idx[l0][dup++];
        } else {
          lay.in_idx[l][k] = idx;
          G.dump[k] = static_cast<uint8_t>(idx++);
        }
      }
    }
    if (p.post_op == POST_COLOR_TAIL) {
      lay.tail = idx;
      p.post_dump = static_cast<uint8_t>(idx);
      idx += 4;
    }
    lay.n_dump = idx;
    p.n_dump = idx;
  };
  assign_forward(g, ctx->lay_geom, Ld);
  assign_forward(c, ctx->lay_color, c.n_layers);
  ctx->prog_geom = g;
  ChainProg q = g;
  q.n_layers = n - 1;
  q.units_per_tile = static_cast<int>(ctx->sdfq_units_n);
  q.post_op = POST_SDF_TAIL;
  ctx->prog_sdfq = q;
  ctx->prog_color = c;

  // ---- reverse chains (training backward): standard 4-chunk 256x256 layers on transposed weights.  The layer that
  // carries W_m^T takes zbar_m (adjoint of forward layer m's pre-activation) as its A operand: chunk ck of reverse
  // layer r is record chunk 4 r + ck of the chain's zbar record, the POST_BWD_DUMP result (zbar_0) is 4 n .. 4 n + 3.
  const int L = cfg.n_layers;
  auto rev_layer = [&](LayerProg& G, uint8_t src, int gate_base, int r, uint8_t act, bool rank1) {
    std::memset(&G, 0, sizeof(G));
    G.n_chunks = 4;
    for (int k = 0; k < MAXC; ++k) G.dump[k] = G.dump_lo[k] = NO_DUMP;
    for (int k = 0; k < 4; ++k) {
      G.src[k] = src;
      G.arg[k] = static_cast<uint8_t>(k);
      G.nsub[k] = 2;
      G.dump[k] = static_cast<uint8_t>(4 * r + k);
    }
    G.gate_base = static_cast<uint8_t>(gate_base);
    G.bwd_act = act;
    G.rank1 = rank1 ? 1 : 0;
    finish_layer(G);
  };
  for (int net = 0; net < 3; ++net) {
    if (net == ES_NET_DEFORM && !cfg.use_deform) continue;
    ChainProg r{};
    common(r);
    r.bias = ctx->geom_bias;
    r.w_units = ctx->rev_units[net];
    r.post_op = POST_BWD_DUMP;
    const es_ctx::FwdLayout& lay = net == ES_NET_COLOR ? ctx->lay_color : ctx->lay_geom;
    const int l0 = net == ES_NET_SDF ? Ld : 0;  // program index of this network's layer 0 in its forward chain
    // record index of the activations a_{m+1} = act(z_m) that gate zbar_m: the SRC_PREV chunks of forward layer m+1
    auto gate_of = [&](int m) {
      if (m + 1 == L - 1 && net != ES_NET_SDF) return lay.tail;  // input of the 3-wide output layer
      return lay.in_idx[l0 + m + 1][0];
    };
    int nr = 0;
    if (net == ES_NET_SDF) {
      rev_layer(r.layer[nr], SRC_ADJ_FEAT, 0, nr, ACT_SOFTPLUS100, false);   // S_{L-1}^T (feature rows)
      ++nr;
      for (int m = L - 2; m >= 1; --m, ++nr)                                  // S_m^T
        rev_layer(r.layer[nr], SRC_BWD_PREV, gate_of(m), nr, ACT_SOFTPLUS100, m == L - 2);
      r.post_bwd_act = ACT_SOFTPLUS100;
    } else {
      for (int m = L - 2; m >= 1; --m, ++nr)
        rev_layer(r.layer[nr], m == L - 2 ? SRC_BWD_OUTER3 : SRC_BWD_PREV, gate_of(m), nr, ACT_RELU, false);
      r.post_bwd_act = ACT_RELU;
      r.outer3_w = net == ES_NET_DEFORM ? ctx->small + SM_DEFORM_W : ctx->small + SM_COLOR_W;
    }
    r.post_gate_base = gate_of(0);
    r.post_dump = static_cast<uint8_t>(4 * nr);
    r.n_layers = nr;
    r.n_dump = 4 * (nr + 1);
    r.n_gate = lay.n_dump;
    r.units_per_tile = nr * 16;
    ctx->rev_layers[net] = nr;
    ctx->prog_rev[net] = r;
  }

  // ---- input-adjoint launches: ONE MMA layer, K = 512 = [zbar_0 | zbar_skip] reloaded from the zbar record,
  // B = the input columns of W_0 and W_skip / sqrt 2 (es_load_network packs them)
  for (int k = 0; k < 3; ++k) {
    const int net = k == 0 ? ES_NET_SDF : ES_NET_COLOR;
    ChainProg r{};
    common(r);
    r.bias = ctx->geom_bias;
    r.w_units = ctx->inadj_units[k];
    r.n_mma = k == 0 ? 64 : (k == 1 ? 256 : 128);
    r.unit_bytes = r.n_mma * SUB_K * 2;
    r.post_op = k == 0 ? POST_INADJ_SDF : (k == 1 ? POST_FEAT_BAR : POST_INADJ_COLOR);
    const int nr = ctx->rev_layers[net];
    const int r_skip = (net == ES_NET_SDF ? 1 : 0) + (L - 2 - skip);  // reverse layer whose input is zbar_skip
    LayerProg& G = r.layer[0];
    std::memset(&G, 0, sizeof(G));
    for (int j = 0; j < MAXC; ++j) G.dump[j] = G.dump_lo[j] = NO_DUMP;
    const bool has_skip = skip > 0 && skip < L - 1;
    G.n_chunks = has_skip ? 8 : 4;
    for (int j = 0; j < G.n_chunks; ++j) {
      G.src[j] = SRC_PLANE;
      G.nsub[j] = 2;
      G.arg[j] = static_cast<uint8_t>(j < 4 ? 4 * nr + j : 4 * r_skip + (j - 4));
    }
    finish_layer(G);
    r.n_layers = 1;
    r.n_plane = 4 * (nr + 1);
    r.units_per_tile = G.n_chunks * 4;
    ctx->prog_inadj[k] = r;
  }
}

}  // namespace

extern "C" {

int es_create(es_ctx** out, const es_net_config* cfg) {
  if (!out || !cfg) return ES_E_BADARG;
  *out = nullptr;
  es_ctx* ctx = new es_ctx();
  ctx->cfg = *cfg;
  const es_net_config& c = ctx->cfg;
  auto bad = [&](const char* m) {
    std::fprintf(stderr, "endosurf_b200: unsupported network config: %s\n", m);
    delete ctx;
    return ES_E_UNSUPPORTED;
  };
  if (c.hidden_dim != HID) return bad("hidden_dim must be 256");
  if (c.n_layers < 3 || 2 * (c.n_layers - 1) + 1 > MAXL || c.n_layers > 16) return bad("n_layers out of range (3..10)");
  if (c.skip_layer == 0 || c.skip_layer >= c.n_layers - 1) return bad("skip_layer must be in [1, n_layers-2] or -1");
  if (c.multires_deform_pos != 6 || c.multires_deform_time != 6 || c.multires_sdf_pos != 6 ||
      c.multires_color_pos != 10 || c.multires_color_dir != 4)
    return bad("multires must be deform 6/6, sdf 6, colour 10/4");
  if (c.precision_terms != 1 && c.precision_terms != 3) return bad("precision_terms must be 1 or 3");
  if (c.skip_layer > 0 && HID - (3 + 6 * 6 + 1 + 2 * 6) <= 0) return bad("deform skip width");

  cudaError_t e = cudaGetDevice(&ctx->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->n_sms, cudaDevAttrMultiProcessorCount, ctx->device);
  int cc_major = 0;
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, ctx->device);
  if (e != cudaSuccess) {
    std::fprintf(stderr, "endosurf_b200: no usable CUDA device: %s\n", cudaGetErrorString(e));
    delete ctx;
    return static_cast<int>(e);
  }
  if (cc_major != 10) {
    std::fprintf(stderr, "endosurf_b200: device is sm_%d0, these kernels are sm_100a only\n", cc_major);
    delete ctx;
    return ES_E_UNSUPPORTED;
  }
  for (int net = 0; net < 3; ++net) build_net_plan(c, net, ctx->plan[net]);
  const NetPlan& D = ctx->plan[ES_NET_DEFORM];
  const NetPlan& S = ctx->plan[ES_NET_SDF];
  const NetPlan& C = ctx->plan[ES_NET_COLOR];
  const int Ld = c.use_deform ? static_cast<int>(D.packs.size()) : 0;
  const int Ls = static_cast<int>(S.packs.size());
  const long long du = c.use_deform ? plan_units(D, Ld) : 0;
  ctx->sdfq_units_n = du + plan_units(S, Ls - 1);
  ctx->geom_units_n = du + plan_units(S, Ls);
  ctx->color_units_n = plan_units(C, static_cast<int>(C.packs.size()));
  // unit offsets
  long long off = 0;
  if (c.use_deform)
    for (auto& k : ctx->plan[ES_NET_DEFORM].packs) { k.unit_off = off; off += 2 * (k.k_total / SUB_K); }
  for (auto& k : ctx->plan[ES_NET_SDF].packs) { k.unit_off = off; off += 2 * (k.k_total / SUB_K); }
  off = 0;
  for (auto& k : ctx->plan[ES_NET_COLOR].packs) { k.unit_off = off; off += 2 * (k.k_total / SUB_K); }

#define CUC(call)                                                                          \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      std::fprintf(stderr, "endosurf_b200: %s: %s\n", #call, cudaGetErrorString(e__));     \
      es_destroy(ctx);                                                                     \
      return static_cast<int>(e__);                                                        \
    }                                                                                      \
  } while (0)
  CUC(cudaMalloc(&ctx->geom_units, static_cast<size_t>(ctx->geom_units_n) * UNIT_BYTES));
  CUC(cudaMalloc(&ctx->color_units, static_cast<size_t>(ctx->color_units_n) * UNIT_BYTES));
  CUC(cudaMalloc(&ctx->geom_bias, static_cast<size_t>(Ld + Ls) * HID * sizeof(float)));
  CUC(cudaMalloc(&ctx->color_bias, C.packs.size() * HID * sizeof(float)));
  CUC(cudaMalloc(&ctx->small, SM_TOTAL_F * sizeof(float)));
  for (int net = 0; net < 3; ++net) {
    const int nl = net == ES_NET_SDF ? c.n_layers - 1 : c.n_layers - 2;
    CUC(cudaMalloc(&ctx->rev_units[net], static_cast<size_t>(nl) * 16 * UNIT_BYTES));
  }
  for (int k = 0; k < 3; ++k) {
    const int n_mma = k == 0 ? 64 : (k == 1 ? 256 : 128);
    CUC(cudaMalloc(&ctx->inadj_units[k], static_cast<size_t>(32) * n_mma * SUB_K * 2));
    CUC(cudaMemset(ctx->inadj_units[k], 0, static_cast<size_t>(32) * n_mma * SUB_K * 2));
    // source columns (reference layout of the layer's input) of the N rows of the operand, layer 0 then skip layer
    const int net = k == 0 ? ES_NET_SDF : ES_NET_COLOR;
    const NetPlan& P = ctx->plan[net];
    const int nx_c = 3 + 6 * c.multires_color_pos, nd_c = 3 + 6 * c.multires_color_dir;
    std::vector<int> cols(2 * n_mma, -1);
    for (int n = 0; n < n_mma; ++n) {
      int rc = -1;
      if (k == 0) rc = ref_column(c, net, chunk_feat(SRC_ENC_SDF, n));
      else if (k == 1) rc = nx_c + 3 + nd_c + n;
      else if (n < 64) rc = ref_column(c, net, chunk_feat(SRC_COLOR_A, n));
      else if (n < 96) rc = ref_column(c, net, chunk_feat(SRC_COLOR_B, n - 64));
      cols[n] = rc;
      cols[n_mma + n] = rc < 0 ? -1 : P.out_dims[0] + rc;  // skip layer input = cat(h[256], network input)
    }
    CUC(cudaMalloc(&ctx->inadj_cols_dev[k], cols.size() * sizeof(int)));
    CUC(cudaMemcpy(ctx->inadj_cols_dev[k], cols.data(), cols.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  for (int net = 0; net < 3; ++net)
    for (int l = 0; l < c.n_layers; ++l)
      CUC(cudaMalloc(&ctx->w_eff[net][l],
                     static_cast<size_t>(ctx->plan[net].out_dims[l]) * ctx->plan[net].in_dims[l] * sizeof(float)));
  CUC(cudaMalloc(&ctx->amax_dev, 4 * sizeof(unsigned int)));
  CUC(cudaMalloc(&ctx->scale_dev, 4 * sizeof(float)));
  CUC(cudaMalloc(&ctx->eik_den_dev, sizeof(float)));
  CUC(cudaMalloc(&ctx->err_dev, sizeof(int)));
  CUC(cudaMemset(ctx->geom_bias, 0, static_cast<size_t>(Ld + Ls) * HID * sizeof(float)));
  CUC(cudaMemset(ctx->color_bias, 0, C.packs.size() * HID * sizeof(float)));
  CUC(cudaMemset(ctx->small, 0, SM_TOTAL_F * sizeof(float)));
  CUC(cudaMemset(ctx->err_dev, 0, sizeof(int)));
  for (int net = 0; net < 3; ++net)
    for (auto& k : ctx->plan[net].packs) {
      CUC(cudaMalloc(&k.colmap_dev, k.colmap.size() * sizeof(int)));
      CUC(cudaMemcpy(k.colmap_dev, k.colmap.data(), k.colmap.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
#undef CUC
  build_chain_programs(ctx);
  if (!c.use_deform) ctx->loaded[ES_NET_DEFORM] = true;
  *out = ctx;
  return 0;
}

void es_destroy(es_ctx* ctx) {
  if (!ctx) return;
  cudaDeviceSynchronize();
  cudaFree(ctx->geom_units);
  cudaFree(ctx->color_units);
  cudaFree(ctx->geom_bias);
  cudaFree(ctx->color_bias);
  cudaFree(ctx->small);
  cudaFree(ctx->err_dev);
  if (ctx->err_host) cudaFreeHost(ctx->err_host);
  for (int net = 0; net < 3; ++net) cudaFree(ctx->rev_units[net]);
  for (int k = 0; k < 3; ++k) {
    cudaFree(ctx->inadj_units[k]);
    cudaFree(ctx->inadj_cols_dev[k]);
  }
  for (int net = 0; net < 3; ++net)
    for (int l = 0; l < 16; ++l) cudaFree(ctx->w_eff[net][l]);
  cudaFree(ctx->amax_dev);
  cudaFree(ctx->scale_dev);
  cudaFree(ctx->eik_den_dev);
  for (auto& wp : ctx->wplans) {
    cudaFree(wp.items_dev);
    cudaFree(wp.colmaps_dev);
  }
  cudaFree(ctx->ws);
  for (int net = 0; net < 3; ++net)
    for (auto& k : ctx->plan[net].packs) cudaFree(k.colmap_dev);
  delete ctx;
}

const char* es_last_error(const es_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int es_num_sms(const es_ctx* ctx) { return ctx ? ctx->n_sms : 0; }
int64_t es_launch_count(const es_ctx* ctx) { return ctx ? ctx->launches : 0; }

int es_sync_check(es_ctx* ctx, void* stream) {
  if (!ctx) return ES_E_BADARG;
  CU(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  int h = 0;
  CU(cudaMemcpy(&h, ctx->err_dev, sizeof(int), cudaMemcpyDeviceToHost));
  if (h != 0) {
    CU(cudaMemset(ctx->err_dev, 0, sizeof(int)));
    return fail(ctx, ES_E_DEVICE, "device-side barrier watchdog tripped, site code " + std::to_string(h));
  }
  return 0;
}

int es_release_workspace(es_ctx* ctx) {
  if (!ctx) return ES_E_BADARG;
  CU(cudaDeviceSynchronize());
  if (ctx->ws) CU(cudaFree(ctx->ws));
  ctx->ws = nullptr;
  ctx->ws_bytes = 0;
  return 0;
}

int es_poll_error(es_ctx* ctx, void* stream) {
  if (!ctx) return ES_E_BADARG;
  if (!ctx->err_host) {
    CU(cudaHostAlloc(reinterpret_cast<void**>(&ctx->err_host), sizeof(int), cudaHostAllocDefault));
    *ctx->err_host = 0;
  }
  // the value an EARLIER poll's copy delivered (never blocks; a tripped watchdog surfaces one call later at most)
  const int h = *reinterpret_cast<volatile int*>(ctx->err_host);
  if (h != 0) {
    *ctx->err_host = 0;
    CU(cudaMemsetAsync(ctx->err_dev, 0, sizeof(int), static_cast<cudaStream_t>(stream)));
    return fail(ctx, ES_E_DEVICE, "device-side barrier watchdog tripped, site code " + std::to_string(h));
  }
  CU(cudaMemcpyAsync(ctx->err_host, ctx->err_dev, sizeof(int), cudaMemcpyDeviceToHost,
                     static_cast<cudaStream_t>(stream)));
  return 0;
}

int es_chunk_colmap(const es_ctx* ctx, int net, int src, int32_t* out64) {
  if (!ctx || !out64 || net < 0 || net > 2) return ES_E_BADARG;
  for (int k = 0; k < 64; ++k) out64[k] = ref_column(ctx->cfg, net, chunk_feat(src, k));
  return 0;
}

int es_load_network(es_ctx* ctx, int net, const float* const* w, const float* const* b, void* stream_) {
  if (!ctx || !w || !b || net < 0 || net > 2) return ES_E_BADARG;
  if (net == ES_NET_DEFORM && !ctx->cfg.use_deform) return fail(ctx, ES_E_BADARG, "context built with use_deform=0");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NetPlan& P = ctx->plan[net];
  const int L = ctx->cfg.n_layers;
  const int Ld = ctx->cfg.use_deform ? static_cast<int>(ctx->plan[ES_NET_DEFORM].packs.size()) : 0;
  uint8_t* units = net == ES_NET_COLOR ? ctx->color_units : ctx->geom_units;
  float* bias = net == ES_NET_COLOR ? ctx->color_bias : ctx->geom_bias + (net == ES_NET_SDF ? Ld * HID : 0);
  for (int l = 0; l < L; ++l)
    if (!w[l] || !b[l]) return fail(ctx, ES_E_BADARG, "null layer pointer");
  PackJobs jobs{};
  constexpr int kMaxJobs = static_cast<int>(sizeof(jobs.j) / sizeof(jobs.j[0]));
  bool overflow = false;
  auto add = [&](const PackJob& j) {
    if (jobs.n < kMaxJobs) jobs.j[jobs.n++] = j;
    else overflow = true;
  };
  auto copy = [&](const float* src, float* dst, int n) {
    PackJob j{};
    j.kind = PACK_COPY;
    j.w = src;
    j.dst = dst;
    j.n_copy = n;
    add(j);
  };
  // forward units + biases of the MMA layers
  for (size_t l = 0; l < P.packs.size(); ++l) {
    const LayerPack& K = P.packs[l];
    PackJob j{};
    j.kind = PACK_FORWARD;
    j.w = w[l] + static_cast<size_t>(K.row_off) * K.n_in;
    j.n_out = K.n_out;
    j.n_in = K.n_in;
    j.cols = K.colmap_dev;
    j.k_total = K.k_total;
    j.scale = K.scale;
    j.units = units + static_cast<size_t>(K.unit_off) * UNIT_BYTES;
    j.pair = ctx->pair_mode;
    add(j);
    if (!(net == ES_NET_SDF && static_cast<int>(l) == L - 1)) copy(b[l], bias + l * HID, K.n_out);
  }
  // output layers kept in fp32 for the epilogue dot products
  const float* wl = w[L - 1];
  const float* bl = b[L - 1];
  if (net == ES_NET_DEFORM) {
    copy(wl, ctx->small + SM_DEFORM_W, 3 * HID);
    copy(bl, ctx->small + SM_DEFORM_B, 3);
  } else if (net == ES_NET_SDF) {
    copy(wl, ctx->small + SM_SDF_W, HID);
    copy(bl, ctx->small + SM_SDF_B, 1);
    copy(bl + 1, ctx->small + SM_FEAT_B, HID);
  } else {
    copy(wl, ctx->small + SM_COLOR_W, 3 * HID);
    copy(bl, ctx->small + SM_COLOR_B, 3);
  }
  // transposed units for the reverse (training) chains
  const float inv_sqrt2 = static_cast<float>(1.0 / std::sqrt(2.0));
  const int skip = ctx->cfg.skip_layer;
  {
    uint8_t* ru = ctx->rev_units[net];
    int r = 0;
    auto transposed = [&](const float* src, int k_valid, int stride, int n_valid, float scale) {
      PackJob j{};
      j.kind = PACK_TRANSPOSED;
      j.w = src;
      j.n_out = k_valid;
      j.n_in = stride;
      j.n_valid = n_valid;
      j.n_mma = HID;
      j.scale = scale;
      j.units = ru + static_cast<size_t>(r++) * 16 * UNIT_BYTES;
      j.pair = ctx->pair_mode;
      add(j);
    };
    if (net == ES_NET_SDF) transposed(w[L - 1] + P.in_dims[L - 1], HID, P.in_dims[L - 1], HID, 1.f);
    for (int m = L - 2; m >= 1; --m)
      transposed(w[m], P.out_dims[m], P.in_dims[m], P.out_dims[m - 1], m == skip ? inv_sqrt2 : 1.f);
  }
  // operands of the input-adjoint launches: the network-input columns of W_0 and of W_skip / sqrt 2
  if (net != ES_NET_DEFORM) {
    for (int k = (net == ES_NET_SDF ? 0 : 1); k <= (net == ES_NET_SDF ? 0 : 2); ++k) {
      const int n_mma = ctx->prog_inadj[k].n_mma;
      const size_t ub = static_cast<size_t>(n_mma) * SUB_K * 2;
      for (int which = 0; which < 2; ++which) {
        if (which == 1 && !(skip > 0 && skip < L - 1)) break;
        const int m = which ? skip : 0;
        PackJob j{};
        j.kind = PACK_TRANSPOSED;
        j.w = w[m];
        j.n_out = P.out_dims[m];
        j.n_in = P.in_dims[m];
        j.cols = ctx->inadj_cols_dev[k] + which * n_mma;
        j.n_mma = n_mma;
        j.scale = which ? inv_sqrt2 : 1.f;
        j.units = ctx->inadj_units[k] + static_cast<size_t>(which) * 16 * ub;
        add(j);
      }
    }
  }
  if (overflow) return fail(ctx, ES_E_UNSUPPORTED, "pack job table too small for this network");
  CU(launch_pack_jobs(jobs, stream));
  ++ctx->launches;
  ctx->loaded[net] = true;
  return 0;
}

int es_load_network_wn(es_ctx* ctx, int net, const float* const* v, const float* const* g, const float* const* b,
                       void* stream_) {
  if (!ctx || !v || !g || !b || net < 0 || net > 2) return ES_E_BADARG;
  if (net == ES_NET_DEFORM && !ctx->cfg.use_deform) return fail(ctx, ES_E_BADARG, "context built with use_deform=0");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const NetPlan& P = ctx->plan[net];
  const int L = ctx->cfg.n_layers;
  WnLayers tab{};
  int max_rows = 0;
  const float* w[16];
  for (int l = 0; l < L; ++l) {
    if (!v[l] || !g[l] || !b[l]) return fail(ctx, ES_E_BADARG, "null layer pointer");
    WnLayer& W = tab.l[l];
    W.v = v[l];
    W.g = g[l];
    W.w_eff = ctx->w_eff[net][l];
    W.n_out = P.out_dims[l];
    W.n_in = P.in_dims[l];
    max_rows = std::max(max_rows, W.n_out);
    w[l] = ctx->w_eff[net][l];
  }
  tab.n = L;
  CU(launch_wn_fold(tab, max_rows, stream));
  ++ctx->launches;
  return es_load_network(ctx, net, w, b, stream_);
}

// profile slots (es_profile)
enum { K_GEOM = 0, K_COLOR = 1, K_SDFQ = 2, K_REV_DEFORM = 3, K_REV_SDF = 4, K_REV_COLOR = 5, K_INADJ = 6, K_WGRAD = 7,
       K_WREDUCE = 8, K_COMPOSITE = 9, K_SMALLM = 10, K_NKINDS = 12 };

// reverse layer (of network `net`'s reverse chain) whose A operand is zbar of the skip layer
static int rev_skip_layer(const es_ctx* ctx, int net) {
  return (net == ES_NET_SDF ? 1 : 0) + (ctx->cfg.n_layers - 2 - ctx->cfg.skip_layer);
}
static bool has_skip(const es_ctx* ctx) {
  return ctx->cfg.skip_layer > 0 && ctx->cfg.skip_layer < ctx->cfg.n_layers - 1;
}

// Which lo halves a training launch keeps / reads (see es_ctx::full_planes).  kind: K_GEOM / K_COLOR forward training
// chains, K_REV_* reverse chains, K_INADJ + k input-adjoint launch k.
static void apply_plane_mode(const es_ctx* ctx, int kind, int inadj_k, ChainProg& p) {
  const bool full = ctx->full_planes != 0;
  if (kind == K_GEOM || kind == K_COLOR) {
    if (full) {
      for (int l = 0; l < p.n_layers; ++l)
        for (int k = 0; k < p.layer[l].n_chunks; ++k) p.layer[l].dump_lo[k] = p.layer[l].dump[k];
      p.pre_dump_lo = p.pre_dump;
      p.post_dump_lo = p.post_dump;
      p.n_dump_lo = p.n_dump;
    } else {
      p.n_dump_lo = 0;
    }
  } else if (kind >= K_REV_DEFORM && kind <= K_REV_COLOR) {
    const int net = kind - K_REV_DEFORM;
    p.gate_use_lo = full ? 1 : 0;
    if (full) {
      for (int l = 0; l < p.n_layers; ++l)
        for (int k = 0; k < p.layer[l].n_chunks; ++k) p.layer[l].dump_lo[k] = p.layer[l].dump[k];
      p.post_dump_lo = p.post_dump;
      p.n_dump_lo = p.n_dump;
    } else if (net != ES_NET_DEFORM) {
      // the input-adjoint launches need zbar_skip and zbar_0 to fp32 accuracy: compact lo record [skip x4 | zbar_0 x4]
      int nlo = 0;
      if (has_skip(ctx)) {
        LayerProg& G = p.layer[rev_skip_layer(ctx, net)];
        for (int k = 0; k < 4; ++k) G.dump_lo[k] = static_cast<uint8_t>(nlo++);
      }
      p.post_dump_lo = static_cast<uint8_t>(nlo);
      nlo += 4;
      p.n_dump_lo = nlo;
    } else {
      p.n_dump_lo = 0;
    }
  } else if (kind == K_INADJ) {
    (void)inadj_k;
    LayerProg& G = p.layer[0];
    if (full) {
      for (int j = 0; j < G.n_chunks; ++j) p.plane_lo[j] = G.arg[j];
      p.n_plane_lo = p.n_plane;
    } else {
      const int skip_n = has_skip(ctx) ? 4 : 0;
      for (int j = 0; j < G.n_chunks; ++j) p.plane_lo[j] = static_cast<uint8_t>(j < 4 ? skip_n + j : j - 4);
      p.n_plane_lo = skip_n + 4;
    }
  }
}

static int timer_begin(es_ctx* ctx, int kind, long long points, cudaStream_t stream, es_ctx::Timed& t) {
  t = es_ctx::Timed{kind, points, nullptr, nullptr};
  if (ctx->profiling) {
    CU(cudaEventCreate(&t.e0));
    CU(cudaEventCreate(&t.e1));
    CU(cudaEventRecord(t.e0, stream));
  }
  return 0;
}
static int timer_end(es_ctx* ctx, cudaStream_t stream, es_ctx::Timed& t) {
  ++ctx->launches;
  if (ctx->profiling) {
    CU(cudaEventRecord(t.e1, stream));
    ctx->timed.push_back(t);
  }
  return 0;
}

static int timed_chain(es_ctx* ctx, int kind, int chain, bool tangent, const ChainProg& prog_in, const ChainIO& io,
                       cudaStream_t stream, bool bwd = false, int inadj_k = 0) {
  ChainProg prog = prog_in;
  if (bwd || io.dump_hi) apply_plane_mode(ctx, kind, inadj_k, prog);
  ChainIO io2 = io;
  io2.trace = (ctx->trace_kind < 0 || ctx->trace_kind == kind) ? ctx->trace_dev : nullptr;
  io2.debug_flags = ctx->debug_flags;
  io2.store_hint = ctx->store_hint;
  es_ctx::Timed t;
  if (int r = timer_begin(ctx, kind, io.n_points, stream, t)) return r;
  const bool pair = ctx->pair_mode && kind != K_INADJ;
  CU(launch_mlp_chain(chain, tangent, pair, prog, io2, ctx->n_sms, stream, bwd));
  return timer_end(ctx, stream, t);
}

static int check_loaded(es_ctx* ctx, bool need_color) {
  if (!ctx->loaded[ES_NET_DEFORM] || !ctx->loaded[ES_NET_SDF] || (need_color && !ctx->loaded[ES_NET_COLOR]))
    return fail(ctx, ES_E_NOWEIGHTS, "es_load_network has not been called for every network");
  return 0;
}

int es_sdf_query(es_ctx* ctx, const float* x, const float* t, int64_t t_div, int64_t t_stride, int64_t n,
                 float* sdf_out, void* stream) {
  if (!ctx || n < 0 || (n > 0 && (!x || !sdf_out)) || t_div <= 0) return ES_E_BADARG;
  if (ctx->cfg.use_deform && !t && n > 0) return fail(ctx, ES_E_BADARG, "time pointer required with use_deform");
  if (int r = check_loaded(ctx, false)) return r;
  if (n == 0) return 0;
  ChainIO io{};
  io.n_points = n;
  io.err = ctx->err_dev;
  io.x = x;
  io.t = t ? t : x;  // never dereferenced meaningfully without deform, but must be a valid pointer
  io.t_div = t ? t_div : 1;
  io.t_stride = t ? t_stride : 0;
  io.out_sdf = sdf_out;
  return timed_chain(ctx, K_SDFQ, CHAIN_SDF, false, ctx->prog_sdfq, io, static_cast<cudaStream_t>(stream));
}

// plane records of one training forward over n points, carved from the caller's stash buffer
struct StashLayout {
  long long tiles_g, tiles_c;
  size_t geom_hi, geom_lo, color_hi, color_lo, total;  // byte offsets (lo = 0-sized unless full planes) and total
};
static StashLayout stash_layout(const es_ctx* ctx, int64_t n) {
  StashLayout s{};
  s.tiles_g = (n + 31) / 32;
  s.tiles_c = (n + TILE_ROWS - 1) / TILE_ROWS;
  const size_t g = static_cast<size_t>(s.tiles_g) * ctx->lay_geom.n_dump * SLOT_HALF_BYTES;
  const size_t c = static_cast<size_t>(s.tiles_c) * ctx->lay_color.n_dump * SLOT_HALF_BYTES;
  size_t off = 0;
  s.geom_hi = off; off += g;
  s.geom_lo = off; off += ctx->full_planes ? g : 0;
  s.color_hi = off; off += c;
  s.color_lo = off; off += ctx->full_planes ? c : 0;
  s.total = off + 256;
  return s;
}

int es_sdf_grid(es_ctx* ctx, const float* bound_min3, const float* bound_max3, int32_t resolution, const float* t,
                float* sdf_out, void* stream_) {
  if (!ctx || !bound_min3 || !bound_max3 || !sdf_out || resolution < 2 || resolution > 2048) return ES_E_BADARG;
  if (ctx->cfg.use_deform && !t) return fail(ctx, ES_E_BADARG, "time pointer required with use_deform");
  if (int r = check_loaded(ctx, false)) return r;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // slabs of about 2 M points: the points never leave the device and the query writes straight into the volume
  const long long plane = static_cast<long long>(resolution) * resolution;
  const int slab = static_cast<int>(std::max<long long>(1, (2LL << 20) / plane));
  if (int r = ensure_ws(ctx, static_cast<size_t>(slab) * plane * 3 * sizeof(float) + 256)) return r;
  float* pts = reinterpret_cast<float*>(ctx->ws);
  for (int x0 = 0; x0 < resolution; x0 += slab) {
    const int nx = std::min(slab, resolution - x0);
    CU(launch_grid_points(bound_min3, bound_max3, resolution, x0, nx, pts, stream));
    ++ctx->launches;
    const long long n = nx * plane;
    if (int r = es_sdf_query(ctx, pts, t, n, 1, n, sdf_out + x0 * plane, stream_)) return r;
  }
  return 0;
}

// bytes of the geometry-feature plane records of n points (colour tiles of 128 points, 4 chunks x (hi, lo) x 16 KiB)
static size_t feat_rec_bytes(int64_t n) {
  return static_cast<size_t>((n + TILE_ROWS - 1) / TILE_ROWS) * 4 * 2 * SLOT_HALF_BYTES;
}

// feat: the caller wants the fp32 feature rows (the colour chain then reads those).  Otherwise the feature goes from
// the geometry chain to the colour chain as fp16 hi/lo plane records (ChainIO::out_feat_rec) in feat_rec, or, when
// that is null too, in scratch carved here.
static int point_forward_impl(es_ctx* ctx, const float* x, const float* t, int64_t t_div, int64_t t_stride,
                              const float* dirs, int64_t dir_div, int64_t dir_stride, int64_t n, float* x_c, float* jac,
                              float* sdf, float* g_c, float* feat, float* rgb, uint8_t* stash, void* stream_,
                              uint8_t* feat_rec = nullptr) {
  if (!ctx || n < 0 || t_div <= 0) return ES_E_BADARG;
  if (n == 0) return 0;
  if (!x) return ES_E_BADARG;
  const bool want_color = rgb != nullptr;
  if (want_color && (!dirs || dir_div <= 0)) return fail(ctx, ES_E_BADARG, "dirs required for rgb");
  if (ctx->cfg.use_deform && !t) return fail(ctx, ES_E_BADARG, "time pointer required with use_deform");
  if (int r = check_loaded(ctx, want_color)) return r;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool deform = ctx->cfg.use_deform != 0;
  // scratch for whatever the caller did not ask for but the colour chain needs
  size_t need = 0;
  {
    Carver c(nullptr);
    if (!x_c) c.take<float>(n * 3);
    if (!jac && deform) c.take<float>(n * 9);
    if (!g_c) c.take<float>(n * 3);
    if (!feat && !feat_rec && want_color) c.take<uint8_t>(feat_rec_bytes(n));  // (also large enough for fp32 rows)
    need = c.off + 256;
  }
  if (int r = ensure_ws(ctx, need)) return r;
  Carver c(ctx->ws);
  float* jac_out = jac;
  if (!x_c) x_c = c.take<float>(n * 3);
  if (!jac && deform) jac = c.take<float>(n * 9);
  if (!g_c) g_c = c.take<float>(n * 3);
  if (!feat && !feat_rec && want_color) feat_rec = c.take<uint8_t>(feat_rec_bytes(n));
  if (feat) feat_rec = nullptr;
  if (feat_rec && !ctx->feat_records) {  // A/B switch: the same scratch holds the fp32 rows instead
    feat = reinterpret_cast<float*>(feat_rec);
    feat_rec = nullptr;
  }
  const StashLayout sl = stash_layout(ctx, n);
  if (feat_rec && n % TILE_ROWS != 0)  // rows past n of the last colour tile: finite values for the MMAs and records
    CU(cudaMemsetAsync(feat_rec + feat_rec_bytes(n) - 4 * 2 * SLOT_HALF_BYTES, 0, 4 * 2 * SLOT_HALF_BYTES, stream));

  ChainIO io{};
  io.n_points = n;
  io.err = ctx->err_dev;
  io.x = x;
  io.t = t ? t : x;
  io.t_div = t ? t_div : 1;
  io.t_stride = t ? t_stride : 0;
  io.out_xc = x_c;
  io.out_jac = deform ? jac : nullptr;
  io.out_sdf = sdf;
  io.out_gc = g_c;
  io.out_feat = feat;
  io.out_feat_rec = want_color ? feat_rec : nullptr;
  if (stash) {
    io.dump_hi = stash + sl.geom_hi;
    io.dump_lo = ctx->full_planes ? stash + sl.geom_lo : nullptr;
  }
  if (int r = timed_chain(ctx, K_GEOM, CHAIN_SDF, true, ctx->prog_geom, io, stream)) return r;
  if (!deform) {
    // canonical = observed space, J = I (endosurf.py:576-577, :626-630)
    CU(cudaMemcpyAsync(x_c, x, static_cast<size_t>(n) * 3 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    if (jac_out) {
      CU(launch_fill_identity_jac(jac_out, n, stream));
      ++ctx->launches;
    }
  }
  if (want_color) {
    ChainIO ic{};
    ic.n_points = n;
    ic.err = ctx->err_dev;
    ic.x_c = x_c;
    ic.g_c = g_c;
    ic.jac = deform ? jac : nullptr;
    ic.dirs = dirs;
    ic.dir_div = dir_div;
    ic.dir_stride = dir_stride;
    ic.feat = feat;
    ic.feat_rec = feat_rec;
    ic.out_rgb = rgb;
    if (stash) {
      ic.dump_hi = stash + sl.color_hi;
      ic.dump_lo = ctx->full_planes ? stash + sl.color_lo : nullptr;
    }
    if (int r = timed_chain(ctx, K_COLOR, CHAIN_COLOR, false, ctx->prog_color, ic, stream)) return r;
  }
  return 0;
}

int es_point_forward(es_ctx* ctx, const float* x, const float* t, int64_t t_div, int64_t t_stride, const float* dirs,
                     int64_t dir_div, int64_t dir_stride, int64_t n, float* x_c, float* jac, float* sdf, float* g_c,
                     float* feat, float* rgb, void* stream) {
  return point_forward_impl(ctx, x, t, t_div, t_stride, dirs, dir_div, dir_stride, n, x_c, jac, sdf, g_c, feat, rgb,
                            nullptr, stream);
}

// =================================================================================================================
// training (differentiable) path
// =================================================================================================================
int es_set_plane_mode(es_ctx* ctx, int32_t full_planes) {
  if (!ctx) return ES_E_BADARG;
  ctx->full_planes = full_planes != 0;
  return 0;
}

int es_train_stash_bytes(const es_ctx* ctx, int64_t n, int64_t* out) {
  if (!ctx || !out || n < 0) return ES_E_BADARG;
  *out = static_cast<int64_t>(stash_layout(ctx, n).total);
  return 0;
}

int es_point_train_forward(es_ctx* ctx, const float* x, const float* t, int64_t t_div, int64_t t_stride,
                           const float* dirs, int64_t dir_div, int64_t dir_stride, int64_t n, float* x_c, float* jac,
                           float* sdf, float* g_c, float* rgb, uint8_t* stash, void* stream) {
  if (!x_c || !sdf || !g_c || !rgb || !stash || (ctx && ctx->cfg.use_deform && !jac)) return ES_E_BADARG;
  return point_forward_impl(ctx, x, t, t_div, t_stride, dirs, dir_div, dir_stride, n, x_c, jac, sdf, g_c, nullptr, rgb,
                            stash, stream);
}

namespace {

// zbar record index (first chunk) of forward layer m of network net; m = L-1 selects the sdf feature rows
int zbar_chunk0(const es_ctx* ctx, int net, int m) {
  const int L = ctx->cfg.n_layers;
  const int nr = ctx->rev_layers[net];
  if (m == 0) return 4 * nr;                         // POST_BWD_DUMP result
  if (net == ES_NET_SDF) return m == L - 1 ? 0 : 4 * (1 + (L - 2 - m));
  return 4 * (L - 2 - m);
}

// base-pointer slots of WgradBases
enum { WB_ZBAR_HI = 0 /* + net */, WB_ZBAR_LO = 3 /* + net */, WB_GEOM_HI = 6, WB_GEOM_LO = 7, WB_COLOR_HI = 8,
       WB_COLOR_LO = 9 };

// Build (or fetch) the weight-gradient work list for n points in the current plane mode.
int get_wgrad_plan(es_ctx* ctx, int64_t n, es_ctx::WgradPlan** out) {
  for (auto& wp : ctx->wplans)
    if (wp.n == n && wp.full == ctx->full_planes) {
      *out = &wp;
      return 0;
    }
  const es_net_config& cfg = ctx->cfg;
  const int L = cfg.n_layers;
  const int Ld = cfg.use_deform ? L - 1 : 0;
  const StashLayout sl = stash_layout(ctx, n);
  const bool full = ctx->full_planes != 0;
  struct Proto {
    int net, layer, mhalf, n_b, a_chunk, b_chunk, bias_mode, row_off, n_out, n_in;
    long long tiles;
    float mul;
    std::vector<int> colmap;
  };
  std::vector<Proto> protos;
  for (int net = 0; net < 3; ++net) {
    if (net == ES_NET_DEFORM && !cfg.use_deform) continue;
    const NetPlan& P = ctx->plan[net];
    const es_ctx::FwdLayout& lay = net == ES_NET_COLOR ? ctx->lay_color : ctx->lay_geom;
    const ChainProg& fp = net == ES_NET_COLOR ? ctx->prog_color : ctx->prog_geom;
    const int l0 = net == ES_NET_SDF ? Ld : 0;
    const int n_mma = static_cast<int>(P.packs.size());  // hidden layers (+ sdf feature layer)
    for (int m = 0; m < n_mma; ++m) {
      const LayerPack& K = P.packs[m];
      const LayerProg& G = fp.layer[l0 + m];
      const bool feat_rows = net == ES_NET_SDF && m == L - 1;
      // runs of consecutive record chunks, at most 4 per group
      int k = 0, koff = 0;
      while (k < G.n_chunks) {
        int len = 1;
        while (k + len < G.n_chunks && len < 4 && lay.in_idx[l0 + m][k + len] == lay.in_idx[l0 + m][k] + len) ++len;
        std::vector<int> cm;
        for (int j = 0; j < len; ++j) {
          const int width = G.nsub[k + j] * SUB_K;
          for (int q = 0; q < 64; ++q) cm.push_back(q < width ? K.colmap[koff + q] : -1);
          koff += width;
        }
        for (int h = 0; h < 2; ++h) {
          if (128 * h >= K.n_out) continue;
          Proto pr;
          pr.net = net;
          pr.layer = m;
          pr.mhalf = h;
          pr.n_b = len;
          pr.a_chunk = zbar_chunk0(ctx, net, m) + 2 * h;
          pr.b_chunk = lay.in_idx[l0 + m][k];
          pr.bias_mode = k == 0 ? (net == ES_NET_COLOR ? 1 : 2) : 0;  // bias sums ride with the first group
          pr.row_off = feat_rows ? 1 : 0;
          pr.n_out = K.n_out;
          pr.n_in = P.in_dims[m];
          pr.tiles = net == ES_NET_COLOR ? sl.tiles_c : sl.tiles_g;
          pr.mul = K.scale;
          pr.colmap = cm;
          protos.push_back(pr);
        }
        k += len;
      }
    }
  }
  const int n_terms = full ? 3 : 1;
  long long total_work = 0;
  for (auto& pr : protos) total_work += pr.tiles * n_terms;
  // short split-K slices: the tensor core accumulates in fp32 with truncation, whose bias grows with the length of one
  // accumulation chain; the partial tiles are summed by the reduce kernel in round-to-nearest fp32
  const long long slice_tiles = std::max<long long>(16, (total_work + ctx->n_sms * 8 - 1) / (ctx->n_sms * 8));
  es_ctx::WgradPlan wp;
  wp.n = n;
  wp.full = ctx->full_planes;
  std::vector<WgradItem> items;
  std::vector<int> colmaps;
  if (protos.size() > sizeof(WgradJobs::j) / sizeof(WgradJob)) return fail(ctx, ES_E_UNSUPPORTED, "too many wgrad jobs");
  for (auto& pr : protos) {
    WgradJob jb{};
    jb.slot0 = static_cast<int>(items.size());
    jb.n_cols = 64 * pr.n_b;
    jb.row0 = 128 * pr.mhalf;
    jb.row_off = pr.row_off;
    jb.n_out = pr.n_out;
    jb.n_in = pr.n_in;
    jb.mul = pr.mul;
    jb.colmap = reinterpret_cast<const int*>(static_cast<uintptr_t>(colmaps.size()));  // offset, patched below
    colmaps.insert(colmaps.end(), pr.colmap.begin(), pr.colmap.end());
    const int rec_a = ctx->prog_rev[pr.net].n_dump;
    const bool color = pr.net == ES_NET_COLOR;
    const int rec_b = color ? ctx->lay_color.n_dump : ctx->lay_geom.n_dump;
    int n_bias = 0;
    // terms: (A hi, B hi) [+ (A lo, B hi) + (A hi, B lo) with every lo plane kept]
    for (int term = 0; term < n_terms; ++term) {
      const bool a_lo = term == 1, b_lo = term == 2;
      for (long long t0 = 0; t0 < pr.tiles; t0 += slice_tiles) {
        WgradItem it{};
        it.a_buf = (a_lo ? WB_ZBAR_LO : WB_ZBAR_HI) + pr.net;
        it.b_buf = color ? (b_lo ? WB_COLOR_LO : WB_COLOR_HI) : (b_lo ? WB_GEOM_LO : WB_GEOM_HI);
        it.a_stride = static_cast<long long>(rec_a) * SLOT_HALF_BYTES;
        it.b_stride = static_cast<long long>(rec_b) * SLOT_HALF_BYTES;
        it.a_off = static_cast<long long>(pr.a_chunk) * SLOT_HALF_BYTES;
        it.b_off = static_cast<long long>(pr.b_chunk) * SLOT_HALF_BYTES;
        it.tile0 = static_cast<int>(t0);
        it.tile1 = static_cast<int>(std::min<long long>(pr.tiles, t0 + slice_tiles));
        it.n_b = pr.n_b;
        it.bias_mode = (pr.bias_mode && term < 2) ? pr.bias_mode : 0;
        it.out = static_cast<int>(items.size());
        if (it.bias_mode) ++n_bias;
        items.push_back(it);
      }
    }
    jb.n_slices = static_cast<int>(items.size()) - jb.slot0;
    jb.n_bias_slices = n_bias;
    wp.jobs.push_back(jb);
    wp.job_net.push_back(pr.net);
    wp.job_layer.push_back(pr.layer);
  }
  // launch order: the two M halves of the same (layer, N group, slice) side by side, so that they can run as one
  // cluster that fetches the shared B tile once (items keep their partial-tile slot in `out`)
  wp.paired = protos.size() % 2 == 0;
  for (size_t j = 0; wp.paired && j + 1 < protos.size(); j += 2)
    wp.paired = protos[j].mhalf == 0 && protos[j + 1].mhalf == 1 && protos[j].net == protos[j + 1].net &&
                protos[j].layer == protos[j + 1].layer && protos[j].b_chunk == protos[j + 1].b_chunk &&
                wp.jobs[j].n_slices == wp.jobs[j + 1].n_slices;
  if (wp.paired) {
    std::vector<WgradItem> launch;
    launch.reserve(items.size());
    for (size_t j = 0; j + 1 < protos.size(); j += 2)
      for (int k = 0; k < wp.jobs[j].n_slices; ++k) {
        launch.push_back(items[wp.jobs[j].slot0 + k]);
        launch.push_back(items[wp.jobs[j + 1].slot0 + k]);
      }
    items.swap(launch);
  }
  wp.n_items = static_cast<int>(items.size());
  CU(cudaMalloc(&wp.items_dev, std::max<size_t>(1, items.size()) * sizeof(WgradItem)));
  CU(cudaMalloc(&wp.colmaps_dev, std::max<size_t>(1, colmaps.size()) * sizeof(int)));
  CU(cudaMemcpy(wp.items_dev, items.data(), items.size() * sizeof(WgradItem), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(wp.colmaps_dev, colmaps.data(), colmaps.size() * sizeof(int), cudaMemcpyHostToDevice));
  for (auto& jb : wp.jobs) jb.colmap = wp.colmaps_dev + reinterpret_cast<uintptr_t>(jb.colmap);
  if (ctx->wplans.size() >= 16) {  // bounded cache
    cudaFree(ctx->wplans.front().items_dev);
    cudaFree(ctx->wplans.front().colmaps_dev);
    ctx->wplans.erase(ctx->wplans.begin());
  }
  ctx->wplans.push_back(std::move(wp));
  *out = &ctx->wplans.back();
  return 0;
}

struct BwdIn {
  int64_t n;
  const float* dirs;
  int64_t dir_div, dir_stride;
  const float* x_c;
  const float* jac;
  const float* g_c;
  const uint8_t* stash;
  float* adj_color;   // [n][4]
  float* adj_sdf;     // [4n][4]
  float* adj_deform;  // [4n][4] or null
};

// scratch of one backward call, carved from the workspace after the caller's own buffers
struct BwdScratch {
  size_t bytes = 0;
  float* feat_bar = nullptr;
  uint8_t* zbar_hi[3] = {nullptr, nullptr, nullptr};
  uint8_t* zbar_lo[3] = {nullptr, nullptr, nullptr};
  float* partial = nullptr;
  float* bias_partial = nullptr;
  float* small_part = nullptr;
  float* gw_eff[3][16] = {};
  uint8_t* gw_first = nullptr;
  size_t gw_bytes = 0;
};
void carve_backward(const es_ctx* ctx, int64_t n, int n_items, Carver& c, BwdScratch& s) {
  const StashLayout sl = stash_layout(ctx, n);
  const bool full = ctx->full_planes != 0;
  s.feat_bar = c.take<float>(n * HID);
  for (int net = 0; net < 3; ++net) {
    if (net == ES_NET_DEFORM && !ctx->cfg.use_deform) continue;
    const long long tiles = net == ES_NET_COLOR ? sl.tiles_c : sl.tiles_g;
    const int rec = ctx->prog_rev[net].n_dump;
    s.zbar_hi[net] = c.take<uint8_t>(static_cast<size_t>(tiles) * rec * SLOT_HALF_BYTES);
    const int rec_lo = full ? rec : (net == ES_NET_DEFORM ? 0 : (has_skip(ctx) ? 8 : 4));
    s.zbar_lo[net] = c.take<uint8_t>(static_cast<size_t>(tiles) * rec_lo * SLOT_HALF_BYTES + 16);
  }
  s.partial = c.take<float>(static_cast<size_t>(n_items) * TILE_ROWS * HID);
  s.bias_partial = c.take<float>(static_cast<size_t>(n_items) * TILE_ROWS);
  s.small_part = c.take<float>(static_cast<size_t>(3) * 2 * ctx->n_sms * (4 * HID + 4));
  s.gw_first = nullptr;
  for (int net = 0; net < 3; ++net)
    for (int l = 0; l < ctx->cfg.n_layers; ++l) {
      s.gw_eff[net][l] = c.take<float>(static_cast<size_t>(ctx->plan[net].out_dims[l]) * ctx->plan[net].in_dims[l]);
      if (!s.gw_first) s.gw_first = reinterpret_cast<uint8_t*>(s.gw_eff[net][l]);
    }
  s.gw_bytes = c.base ? static_cast<size_t>(c.base + c.off - s.gw_first) : 0;  // contiguous: cleared by one memset
  s.bytes = c.off + 256;
}

// The reverse pass of the point pipeline from prepared adjoint rows to d loss / d (weight_v, weight_g, bias).
int backward_core(es_ctx* ctx, const BwdIn& a, const BwdScratch& s, es_ctx::WgradPlan& wp, const es_train_params* prm,
                  cudaStream_t stream) {
  const es_net_config& cfg = ctx->cfg;
  const int L = cfg.n_layers;
  const bool deform = cfg.use_deform != 0;
  const int64_t n = a.n;
  const StashLayout sl = stash_layout(ctx, n);
  const bool full = ctx->full_planes != 0;
  es_ctx::Timed t;
  CU(cudaMemsetAsync(ctx->amax_dev, 0, 4 * sizeof(unsigned int), stream));
  CU(cudaMemsetAsync(s.gw_first, 0, s.gw_bytes, stream));  // every effective-weight gradient (carved contiguously)
  auto rev_chain = [&](int net, const float* adj, const float* adj_feat) -> int {
    ChainIO io{};
    io.n_points = n;
    io.err = ctx->err_dev;
    io.adj = adj;
    io.adj_feat = adj_feat;
    io.scale = ctx->scale_dev + net;
    const bool color = net == ES_NET_COLOR;
    io.gate_hi = a.stash + (color ? sl.color_hi : sl.geom_hi);
    io.gate_lo = full ? a.stash + (color ? sl.color_lo : sl.geom_lo) : nullptr;
    io.dump_hi = s.zbar_hi[net];
    io.dump_lo = s.zbar_lo[net];
    io.t_div = 1;
    io.dir_div = 1;
    return timed_chain(ctx, K_REV_DEFORM + net, color ? CHAIN_COLOR : CHAIN_SDF, !color, ctx->prog_rev[net], io, stream,
                       true);
  };
  auto inadj = [&](int k) -> int {
    const int net = k == 0 ? ES_NET_SDF : ES_NET_COLOR;
    ChainIO io{};
    io.n_points = n;
    io.err = ctx->err_dev;
    io.scale = ctx->scale_dev + net;
    io.plane_hi = s.zbar_hi[net];
    io.plane_lo = s.zbar_lo[net];
    io.x_c = a.x_c;
    io.g_c = a.g_c;
    io.jac = deform ? a.jac : nullptr;
    io.dirs = a.dirs;
    io.dir_div = a.dir_div;
    io.dir_stride = a.dir_stride;
    io.feat_bar = s.feat_bar;
    io.adj_sdf = a.adj_sdf;
    io.adj_deform = deform ? a.adj_deform : nullptr;
    io.amax_bits = ctx->amax_dev + ES_NET_SDF;
    io.t_div = 1;
    return timed_chain(ctx, K_INADJ, k == 0 ? CHAIN_SDF : CHAIN_COLOR, k == 0, ctx->prog_inadj[k], io, stream, true, k);
  };
  auto set_scale = [&](int net, const float* adj, long long count) -> int {
    CU(launch_amax(adj, count, ctx->amax_dev + net, stream));
    CU(launch_scale_from_amax(ctx->amax_dev + net, ctx->scale_dev + net, ctx->scale_target, stream));
    ctx->launches += 2;
    return 0;
  };

  // ---- colour network
  if (int r = set_scale(ES_NET_COLOR, a.adj_color, n * 4)) return r;
  if (int r = rev_chain(ES_NET_COLOR, a.adj_color, nullptr)) return r;
  if (int r = inadj(1)) return r;   // feat_bar (+ its amax into the sdf slot)
  if (int r = inadj(2)) return r;   // adj_sdf += d/d g_c ; adj_deform += d/d (x_c, J)
  // ---- sdf network
  if (int r = set_scale(ES_NET_SDF, a.adj_sdf, n * 16)) return r;
  if (int r = rev_chain(ES_NET_SDF, a.adj_sdf, s.feat_bar)) return r;
  // ---- deformation network
  if (deform) {
    if (int r = inadj(0)) return r;  // adj_deform += d/d x_c through enc6(x_c)
    if (int r = set_scale(ES_NET_DEFORM, a.adj_deform, n * 16)) return r;
    if (int r = rev_chain(ES_NET_DEFORM, a.adj_deform, nullptr)) return r;
  }

  // ---- 3-wide output layers (fp32 adjoints x kept layer inputs, CUDA cores)
  {
    if (int r = timer_begin(ctx, K_SMALLM, n, stream, t)) return r;
    const int nb_c = smallm_blocks(sl.tiles_c, ctx->n_sms), nb_g = smallm_blocks(sl.tiles_g, ctx->n_sms);
    const size_t per = static_cast<size_t>(2) * ctx->n_sms * (4 * HID + 4);
    const long long gs = static_cast<long long>(ctx->lay_geom.n_dump) * SLOT_HALF_BYTES;
    const long long cs = static_cast<long long>(ctx->lay_color.n_dump) * SLOT_HALF_BYTES;
    // colour output layer
    CU(launch_smallm_wgrad(a.stash + sl.color_hi + static_cast<size_t>(ctx->lay_color.tail) * SLOT_HALF_BYTES, cs,
                           sl.tiles_c, a.adj_color, 0, n, s.small_part, nb_c, stream));
    CU(launch_smallm_reduce(s.small_part, nb_c, 0, 3, s.gw_eff[ES_NET_COLOR][L - 1], prm->grad_b[ES_NET_COLOR][L - 1],
                            0, HID, stream));
    // sdf row of the sdf output layer: input = the feature layer's input chunks
    const int Ld = deform ? L - 1 : 0;
    CU(launch_smallm_wgrad(
        a.stash + sl.geom_hi + static_cast<size_t>(ctx->lay_geom.in_idx[Ld + L - 1][0]) * SLOT_HALF_BYTES, gs,
        sl.tiles_g, a.adj_sdf, 1, n, s.small_part + per, nb_g, stream));
    CU(launch_smallm_reduce(s.small_part + per, nb_g, 3, 1, s.gw_eff[ES_NET_SDF][L - 1], prm->grad_b[ES_NET_SDF][L - 1],
                            0, HID, stream));
    ctx->launches += 3;
    if (deform) {
      CU(launch_smallm_wgrad(a.stash + sl.geom_hi + static_cast<size_t>(ctx->lay_geom.tail) * SLOT_HALF_BYTES, gs,
                             sl.tiles_g, a.adj_deform, 1, n, s.small_part + 2 * per, nb_g, stream));
      CU(launch_smallm_reduce(s.small_part + 2 * per, nb_g, 0, 3, s.gw_eff[ES_NET_DEFORM][L - 1],
                              prm->grad_b[ES_NET_DEFORM][L - 1], 0, HID, stream));
      ctx->launches += 2;
    }
    if (int r = timer_end(ctx, stream, t)) return r;
  }

  // ---- 256-wide layers: split-K tcgen05 weight gradients over the plane records, then reduce + scatter
  {
    WgradBases bases{};
    for (int net = 0; net < 3; ++net) {
      bases.p[WB_ZBAR_HI + net] = s.zbar_hi[net];
      bases.p[WB_ZBAR_LO + net] = s.zbar_lo[net];
    }
    bases.p[WB_GEOM_HI] = a.stash + sl.geom_hi;
    bases.p[WB_GEOM_LO] = a.stash + sl.geom_lo;
    bases.p[WB_COLOR_HI] = a.stash + sl.color_hi;
    bases.p[WB_COLOR_LO] = a.stash + sl.color_lo;
    if (int r = timer_begin(ctx, K_WGRAD, n, stream, t)) return r;
    CU(launch_wgrad(wp.items_dev, wp.n_items, bases, s.partial, s.bias_partial, ctx->wgrad_lbo, ctx->wgrad_sbo,
                    ctx->err_dev, stream, wp.paired && ctx->wgrad_pairs));
    if (int r = timer_end(ctx, stream, t)) return r;
    WgradJobs jobs{};
    jobs.n = static_cast<int>(wp.jobs.size());
    for (int j = 0; j < jobs.n; ++j) {
      WgradJob jb = wp.jobs[j];
      const int net = wp.job_net[j], layer = wp.job_layer[j];
      jb.gw = s.gw_eff[net][layer];
      jb.gb = jb.n_bias_slices > 0 ? prm->grad_b[net][layer] : nullptr;
      jb.scale = ctx->scale_dev + net;
      jobs.j[j] = jb;
    }
    if (int r = timer_begin(ctx, K_WREDUCE, n, stream, t)) return r;
    CU(launch_wgrad_reduce(jobs, s.partial, s.bias_partial, stream));
    // weight-norm backward: d W -> d (weight_g, weight_v)
    WnLayers tab{};
    int max_rows = 0;
    for (int net = 0; net < 3; ++net) {
      if (net == ES_NET_DEFORM && !deform) continue;
      for (int l = 0; l < L; ++l) {
        WnLayer& W = tab.l[tab.n++];
        W.v = prm->v[net][l];
        W.g = prm->g[net][l];
        W.gw = s.gw_eff[net][l];
        W.gv = prm->grad_v[net][l];
        W.gg = prm->grad_g[net][l];
        W.n_out = ctx->plan[net].out_dims[l];
        W.n_in = ctx->plan[net].in_dims[l];
        max_rows = std::max(max_rows, W.n_out);
      }
    }
    CU(launch_wn_backward(tab, max_rows, stream));
    ++ctx->launches;
    if (int r = timer_end(ctx, stream, t)) return r;
  }
  return 0;
}

int check_train_params(es_ctx* ctx, const es_train_params* prm) {
  if (!prm) return ES_E_BADARG;
  for (int net = 0; net < 3; ++net) {
    if (net == ES_NET_DEFORM && !ctx->cfg.use_deform) continue;
    if (!prm->v[net] || !prm->g[net] || !prm->grad_v[net] || !prm->grad_g[net] || !prm->grad_b[net])
      return fail(ctx, ES_E_BADARG, "es_train_params: null table");
    for (int l = 0; l < ctx->cfg.n_layers; ++l)
      if (!prm->v[net][l] || !prm->g[net][l] || !prm->grad_v[net][l] || !prm->grad_g[net][l] || !prm->grad_b[net][l])
        return fail(ctx, ES_E_BADARG, "es_train_params: null layer pointer");
  }
  return 0;
}

}  // namespace

int es_point_train_backward(es_ctx* ctx, int64_t n, const float* dirs, int64_t dir_div, int64_t dir_stride,
                            const float* x_c, const float* jac, const float* g_c, const float* rgb,
                            const uint8_t* stash, const float* sdf_bar, const float* gc_bar, const float* jac_bar,
                            const float* rgb_bar, const es_train_params* prm, void* stream_) {
  if (!ctx || n < 0 || !x_c || !g_c || !rgb || !stash || !dirs || dir_div <= 0) return ES_E_BADARG;
  if (ctx->cfg.use_deform && !jac) return ES_E_BADARG;
  if (int r = check_train_params(ctx, prm)) return r;
  if (int r = check_loaded(ctx, true)) return r;
  if (n == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool deform = ctx->cfg.use_deform != 0;
  es_ctx::WgradPlan* wp = nullptr;
  if (int r = get_wgrad_plan(ctx, n, &wp)) return r;
  BwdScratch s;
  float *adj_c, *adj_s, *adj_d;
  for (int pass = 0; pass < 2; ++pass) {
    Carver c(pass ? ctx->ws : nullptr);
    adj_c = c.take<float>(n * 4);
    adj_s = c.take<float>(n * 16);
    adj_d = c.take<float>(deform ? n * 16 : 0);
    carve_backward(ctx, n, wp->n_items, c, s);
    if (!pass)
      if (int r = ensure_ws(ctx, s.bytes)) return r;
  }
  CU(launch_point_adjoints(n, rgb, sdf_bar, gc_bar, deform ? jac_bar : nullptr, rgb_bar, adj_c, adj_s,
                           deform ? adj_d : nullptr, stream));
  ++ctx->launches;
  BwdIn a{n, dirs, dir_div, dir_stride, x_c, jac, g_c, stash, adj_c, adj_s, deform ? adj_d : nullptr};
  return backward_core(ctx, a, s, *wp, prm, stream);
}

int es_render_train_forward(es_ctx* ctx, const float* rays, int64_t n_rays, const float* z_vals, int32_t m,
                            int32_t n_samples, float cos_anneal_ratio, const float* variance, float* x_c, float* jac,
                            float* sdf, float* g_c, float* rgb, uint8_t* stash, const es_render_out* out,
                            float* eik_den, void* stream_) {
  if (!ctx || !rays || !z_vals || !variance || !x_c || !sdf || !g_c || !rgb || !stash || !out || !eik_den ||
      n_rays < 0 || m < 1 || m > 256 || n_samples < 2)
    return ES_E_BADARG;
  if (!out->color_map || !out->depth_map || !out->gradients_o || !out->gradient_o_error || !out->weights ||
      !out->weight_max || !out->cdf || !out->s_val)
    return fail(ctx, ES_E_BADARG, "null output pointer");
  if (ctx->cfg.use_deform && !jac) return ES_E_BADARG;
  if (int r = check_loaded(ctx, true)) return r;
  if (n_rays == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool deform = ctx->cfg.use_deform != 0;
  const int64_t n = n_rays * m;
  const float sample_dist = 2.0f / static_cast<float>(n_samples);
  // mid-points and per-ray partials live past the scratch point_forward_impl carves for itself (feat)
  const size_t fwd_need = feat_rec_bytes(n) + 1024;
  size_t off_pts, off_eik, need;
  {
    Carver c(nullptr);
    c.off = fwd_need;
    c.take<float>(n * 3);
    off_pts = c.off - static_cast<size_t>(n) * 3 * sizeof(float);
    c.take<float>(n_rays * 2);
    off_eik = c.off - static_cast<size_t>(n_rays) * 2 * sizeof(float);
    need = c.off + 256;
  }
  if (int r = ensure_ws(ctx, need)) return r;
  float* pts = reinterpret_cast<float*>(ctx->ws + off_pts);
  float* eik = reinterpret_cast<float*>(ctx->ws + off_eik);
  RayGeom rg{rays, n_rays};
  CU(launch_points_from_z(rg, z_vals, m, 1, sample_dist, pts, stream));
  ++ctx->launches;
  if (int r = point_forward_impl(ctx, pts, rays + 8, m, 9, rays + 3, m, 9, n, x_c, deform ? jac : nullptr, sdf, g_c,
                                 nullptr, rgb, stash, stream_))
    return r;
  es_ctx::Timed t;
  if (int r = timer_begin(ctx, K_COMPOSITE, n, stream, t)) return r;
  CompositeOut co;
  co.color_map = out->color_map;
  co.depth_map = out->depth_map;
  co.gradients_o = out->gradients_o;
  co.weights = out->weights;
  co.cdf = out->cdf;
  co.weight_max = out->weight_max;
  co.eik_partial = eik;
  co.s_val = out->s_val;
  CU(launch_composite(rg, z_vals, m, sample_dist, sdf, g_c, deform ? jac : nullptr, rgb, variance, cos_anneal_ratio,
                      co, stream));
  CU(launch_eikonal_reduce(eik, n_rays, out->gradient_o_error, eik_den, stream));
  ++ctx->launches;
  return timer_end(ctx, stream, t);
}

int es_render_train_backward(es_ctx* ctx, const float* rays, int64_t n_rays, const float* z_vals, int32_t m,
                             int32_t n_samples, float cos_anneal_ratio, const float* variance, const float* x_c,
                             const float* jac, const float* sdf, const float* g_c, const float* rgb,
                             const uint8_t* stash, const float* eik_den, const es_render_grads* bar,
                             const es_train_params* prm, float* variance_grad, void* stream_) {
  if (!ctx || !rays || !z_vals || !variance || !x_c || !sdf || !g_c || !rgb || !stash || !eik_den || !bar ||
      !variance_grad || n_rays < 0 || m < 1 || m > 256 || n_samples < 2)
    return ES_E_BADARG;
  if (ctx->cfg.use_deform && !jac) return ES_E_BADARG;
  if (int r = check_train_params(ctx, prm)) return r;
  if (int r = check_loaded(ctx, true)) return r;
  if (n_rays == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool deform = ctx->cfg.use_deform != 0;
  const int64_t n = n_rays * m;
  const float sample_dist = 2.0f / static_cast<float>(n_samples);
  es_ctx::WgradPlan* wp = nullptr;
  if (int r = get_wgrad_plan(ctx, n, &wp)) return r;
  BwdScratch s;
  float *adj_c, *adj_s, *adj_d, *invs;
  for (int pass = 0; pass < 2; ++pass) {
    Carver c(pass ? ctx->ws : nullptr);
    adj_c = c.take<float>(n * 4);
    adj_s = c.take<float>(n * 16);
    adj_d = c.take<float>(deform ? n * 16 : 0);
    invs = c.take<float>(n_rays);
    carve_backward(ctx, n, wp->n_items, c, s);
    if (!pass)
      if (int r = ensure_ws(ctx, s.bytes)) return r;
  }
  es_ctx::Timed t;
  if (int r = timer_begin(ctx, K_COMPOSITE, n, stream, t)) return r;
  CompositeBwd cb{};
  cb.color_bar = bar->color_map;
  cb.depth_bar = bar->depth_map;
  cb.go_bar = bar->gradients_o;
  cb.weights_bar = bar->weights;
  cb.cdf_bar = bar->cdf;
  cb.sdf_bar = bar->sdf;
  cb.rgb_bar = bar->sampled_color;
  cb.eik_bar = bar->gradient_o_error;
  cb.eik_den = eik_den;
  cb.adj_color = adj_c;
  cb.adj_sdf = adj_s;
  cb.adj_deform = deform ? adj_d : nullptr;
  cb.invs_partial = invs;
  RayGeom rg{rays, n_rays};
  CU(launch_composite_bwd(rg, z_vals, m, sample_dist, sdf, g_c, deform ? jac : nullptr, rgb, variance,
                          cos_anneal_ratio, cb, stream));
  CU(launch_sum_reduce(invs, n_rays, variance_grad, 0, stream));
  ++ctx->launches;
  if (int r = timer_end(ctx, stream, t)) return r;
  BwdIn a{n, rays + 3, m, 9, x_c, jac, g_c, stash, adj_c, adj_s, deform ? adj_d : nullptr};
  return backward_core(ctx, a, s, *wp, prm, stream);
}

int es_wgrad_probe(es_ctx* ctx, const uint8_t* zbar_rec, const uint8_t* in_rec, int64_t n_tiles, int32_t n_b,
                   int32_t bias_mode, float* out, float* bias_out, void* stream_) {
  if (!ctx || !zbar_rec || !in_rec || !out || n_tiles < 1 || n_b < 1 || n_b > 4 || bias_mode < 0 || bias_mode > 2)
    return ES_E_BADARG;
  if (bias_mode && !bias_out) return ES_E_BADARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  std::vector<WgradItem> items;
  WgradJobs jobs{};
  const long long slice = std::max<long long>(1, (n_tiles + 2) / 3);  // three slices: exercises the split-K reduce
  for (int h = 0; h < 2; ++h) {
    WgradJob jb{};
    jb.slot0 = static_cast<int>(items.size());
    for (long long t0 = 0; t0 < n_tiles; t0 += slice) {
      WgradItem it{};
      it.a_buf = 0;
      it.b_buf = 1;
      it.a_stride = 4LL * SLOT_HALF_BYTES;
      it.b_stride = static_cast<long long>(n_b) * SLOT_HALF_BYTES;
      it.a_off = 2LL * h * SLOT_HALF_BYTES;
      it.b_off = 0;
      it.tile0 = static_cast<int>(t0);
      it.tile1 = static_cast<int>(std::min<long long>(n_tiles, t0 + slice));
      it.n_b = n_b;
      it.bias_mode = bias_mode;
      it.out = static_cast<int>(items.size());
      items.push_back(it);
    }
    jb.n_slices = static_cast<int>(items.size()) - jb.slot0;
    jb.n_bias_slices = bias_mode ? jb.n_slices : 0;
    jb.n_cols = 64 * n_b;
    jb.row0 = 128 * h;
    jb.n_out = 256;
    jb.n_in = 64 * n_b;
    jb.gw = out;
    jb.gb = bias_mode ? bias_out : nullptr;
    jb.mul = 1.f;
    jobs.j[jobs.n++] = jb;
  }
  WgradItem* items_dev = nullptr;
  float* partial = nullptr;
  float* bias_partial = nullptr;
  CU(cudaMalloc(&items_dev, items.size() * sizeof(WgradItem)));
  CU(cudaMalloc(&partial, items.size() * TILE_ROWS * HID * sizeof(float)));
  CU(cudaMalloc(&bias_partial, items.size() * TILE_ROWS * sizeof(float)));
  CU(cudaMemcpy(items_dev, items.data(), items.size() * sizeof(WgradItem), cudaMemcpyHostToDevice));
  WgradBases bases{};
  bases.p[0] = zbar_rec;
  bases.p[1] = in_rec;
  bool paired = ctx->wgrad_pairs != 0;
  if (paired) {  // launch order (half 0, half 1) per slice
    std::vector<WgradItem> launch;
    const size_t ns = items.size() / 2;
    for (size_t k = 0; k < ns; ++k) {
      launch.push_back(items[k]);
      launch.push_back(items[ns + k]);
    }
    CU(cudaMemcpy(items_dev, launch.data(), launch.size() * sizeof(WgradItem), cudaMemcpyHostToDevice));
  }
  CU(launch_wgrad(items_dev, static_cast<int>(items.size()), bases, partial, bias_partial, ctx->wgrad_lbo,
                  ctx->wgrad_sbo, ctx->err_dev, stream, paired));
  CU(launch_wgrad_reduce(jobs, partial, bias_partial, stream));
  ctx->launches += 2;
  CU(cudaStreamSynchronize(stream));
  CU(cudaFree(items_dev));
  CU(cudaFree(partial));
  CU(cudaFree(bias_partial));
  return 0;
}

int es_debug_set(es_ctx* ctx, int32_t key, int32_t value) {
  if (!ctx) return ES_E_BADARG;
  switch (key) {
    case 0: ctx->debug_flags = value; return 0;   // ES_ABLATE builds only
    case 1: ctx->wgrad_lbo = value; return 0;     // MN-major descriptor strides of the weight-gradient kernel
    case 2: ctx->wgrad_sbo = value; return 0;
    case 3: ctx->scale_target = std::ldexp(1.f, value); return 0;  // adjoint scale target 2^value
    case 5: ctx->wgrad_pairs = value != 0; return 0;
    case 8: ctx->store_hint = value; return 0;
    case 7: ctx->feat_records = value != 0; return 0;  // A/B switch of the feature plane records
    case 6: ctx->trace_kind = value; return 0;    // ES_TRACE builds: only launches of this kind write the trace (-1: all)
    case 4:  // CTA pairs on/off; the packed weights change layout: the networks must be loaded again
      ctx->pair_mode = value != 0;
      ctx->loaded[ES_NET_SDF] = ctx->loaded[ES_NET_COLOR] = false;
      ctx->loaded[ES_NET_DEFORM] = !ctx->cfg.use_deform;
      return 0;
    default: return ES_E_BADARG;
  }
}

int es_up_sample(es_ctx* ctx, const float* rays, int64_t n_rays, const float* z, const float* sdf, int32_t n,
                 int32_t n_imp, const float* u_vals, float inv_s, float* new_z, void* stream) {
  if (!ctx || !rays || !z || !sdf || !u_vals || !new_z || n_rays < 0 || n < 2 || n_imp < 1) return ES_E_BADARG;
  RayGeom rg{rays, n_rays};
  CU(launch_upsample(rg, z, sdf, n, n_imp, u_vals, inv_s, new_z, static_cast<cudaStream_t>(stream)));
  ++ctx->launches;
  return 0;
}

int es_render_rays(es_ctx* ctx, const float* rays, int64_t n_rays, const es_render_params* p,
                   const es_render_out* out, void* stream_) {
  if (!ctx || !p || !out || n_rays < 0) return ES_E_BADARG;
  if (n_rays == 0) return 0;
  const bool sample_only = out->color_map == nullptr;  // hierarchical sampling only (training path): z_vals out
  if (sample_only) {
    if (!rays || !out->z_vals || p->z_override) return fail(ctx, ES_E_BADARG, "sampling-only call needs rays and z_vals");
  } else if (!rays || !p->variance || !out->depth_map || !out->gradients_o || !out->gradient_o_error ||
             !out->weights || !out->weight_max || !out->cdf || !out->s_val) {
    return fail(ctx, ES_E_BADARG, "null input/output pointer");
  }
  if (int r = check_loaded(ctx, true)) return r;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int ns = p->n_samples;
  const bool up = p->do_upsample && p->n_importance > 0 && !p->z_override;
  const int steps = p->up_sample_steps;
  if (ns < 2) return fail(ctx, ES_E_BADARG, "n_samples < 2");
  if (up && (steps < 1 || p->n_importance % steps != 0 || !p->u_vals))
    return fail(ctx, ES_E_BADARG, "n_importance must be a multiple of up_sample_steps; u_vals required");
  if (!p->z_override && !p->t_vals) return fail(ctx, ES_E_BADARG, "t_vals required");
  const int n_imp = up ? p->n_importance / steps : 0;
  const int M = p->z_override ? ns + (p->do_upsample ? p->n_importance : 0) : ns + (up ? p->n_importance : 0);
  if (M > 256) return fail(ctx, ES_E_UNSUPPORTED, "more than 256 samples per ray");
  const float sample_dist = 2.0f / static_cast<float>(ns);
  const bool deform = ctx->cfg.use_deform != 0;

  const int64_t RC = std::min<int64_t>(n_rays, 8192);  // rays per pass: bounds the scratch (feat is 1 KiB/point)
  const int64_t PC = RC * M;
  size_t need;
  {
    Carver c(nullptr);
    c.take<float>(RC * M);      // zA
    c.take<float>(RC * M);      // zB
    c.take<float>(RC * M);      // sdfA
    c.take<float>(RC * M);      // sdfB
    c.take<float>(RC * 64);     // new_z (n_imp <= 64 enforced below)
    c.take<float>(RC * 64);     // new_sdf
    c.take<float>(PC * 3);      // pts
    c.take<float>(PC * 3);      // x_c
    c.take<float>(PC * 9);      // jac
    c.take<float>(PC * 3);      // g_c
    c.take<float>(PC);          // sdf
    c.take<float>(PC * 3);      // rgb
    c.take<uint8_t>(feat_rec_bytes(PC));  // geometry-feature plane records
    c.take<float>(n_rays * 2);  // eikonal partials (all rays)
    need = c.off + 256;
  }
  if (n_imp > 64) return fail(ctx, ES_E_UNSUPPORTED, "n_importance / up_sample_steps > 64");
  if (int r = ensure_ws(ctx, need)) return r;
  Carver c(ctx->ws);
  float* zA = c.take<float>(RC * M);
  float* zB = c.take<float>(RC * M);
  float* sA = c.take<float>(RC * M);
  float* sB = c.take<float>(RC * M);
  float* new_z = c.take<float>(RC * 64);
  float* new_sdf = c.take<float>(RC * 64);
  float* pts = c.take<float>(PC * 3);
  float* x_c = c.take<float>(PC * 3);
  float* jac = c.take<float>(PC * 9);
  float* g_c = c.take<float>(PC * 3);
  float* sdf = c.take<float>(PC);
  float* rgb = c.take<float>(PC * 3);
  uint8_t* feat_rec = c.take<uint8_t>(feat_rec_bytes(PC));
  float* eik = c.take<float>(n_rays * 2);

  for (int64_t r0 = 0; r0 < n_rays; r0 += RC) {
    const int64_t R = std::min<int64_t>(RC, n_rays - r0);
    RayGeom rg{rays + r0 * 9, R};
    const float* tptr = rays + r0 * 9 + 8;  // time = rays[:, 8]
    float* z = zA;
    float* zalt = zB;
    float* s = sA;
    float* salt = sB;
    int n = ns;
    if (p->z_override) {
      z = const_cast<float*>(p->z_override) + r0 * M;
      n = M;
    } else {
      CU(launch_coarse_z(rg, ns, p->t_vals, p->t_rand ? p->t_rand + r0 : nullptr, sample_dist, z, stream));
      ++ctx->launches;
      if (up) {
        // no-grad hierarchical sampling (endosurf.py:85-110)
        CU(launch_points_from_z(rg, z, n, 0, sample_dist, pts, stream));
        ++ctx->launches;
        if (int rr = es_sdf_query(ctx, pts, tptr, n, 9, R * n, s, stream_)) return rr;
        for (int i = 0; i < steps; ++i) {
          const bool last = (i + 1 == steps);
          CU(launch_upsample(rg, z, s, n, n_imp, p->u_vals, 64.f * static_cast<float>(1 << i), new_z, stream));
          ++ctx->launches;
          if (!last) {
            CU(launch_points_from_z(RayGeom{rg.rays, R}, new_z, n_imp, 0, sample_dist, pts, stream));
            ++ctx->launches;
            if (int rr = es_sdf_query(ctx, pts, tptr, n_imp, 9, R * n_imp, new_sdf, stream_)) return rr;
          }
          CU(launch_merge_z(R, z, s, n, new_z, last ? nullptr : new_sdf, n_imp, zalt, salt, stream));
          ++ctx->launches;
          std::swap(z, zalt);
          std::swap(s, salt);
          n += n_imp;
        }
      }
    }
    if (sample_only) {
      CU(cudaMemcpyAsync(out->z_vals + r0 * M, z, R * M * sizeof(float), cudaMemcpyDeviceToDevice, stream));
      continue;
    }
    // render_core (endosurf.py:134-213)
    CU(launch_points_from_z(rg, z, n, 1, sample_dist, pts, stream));
    ++ctx->launches;
    float* sdf_dst = out->sdf ? out->sdf + r0 * M : sdf;
    float* rgb_dst = out->sampled_color ? out->sampled_color + r0 * M * 3 : rgb;
    if (int rr = point_forward_impl(ctx, pts, tptr, n, 9, rays + r0 * 9 + 3, n, 9, R * n, x_c, deform ? jac : nullptr,
                                    sdf_dst, g_c, nullptr, rgb_dst, nullptr, stream_, feat_rec))
      return rr;
    CompositeOut co;
    co.color_map = out->color_map + r0 * 3;
    co.depth_map = out->depth_map + r0;
    co.gradients_o = out->gradients_o + r0 * M * 3;
    co.weights = out->weights + r0 * M;
    co.cdf = out->cdf + r0 * M;
    co.weight_max = out->weight_max + r0;
    co.eik_partial = eik + r0 * 2;
    co.s_val = out->s_val + r0;
    CU(launch_composite(rg, z, n, sample_dist, sdf_dst, g_c, deform ? jac : nullptr, rgb_dst, p->variance,
                        p->cos_anneal_ratio, co, stream));
    ++ctx->launches;
    if (out->z_vals && z != out->z_vals + r0 * M)
      CU(cudaMemcpyAsync(out->z_vals + r0 * M, z, R * M * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  }
  if (sample_only) return 0;
  CU(launch_eikonal_reduce(eik, n_rays, out->gradient_o_error, nullptr, stream));
  ++ctx->launches;
  return 0;
}

int es_profile_enable(es_ctx* ctx, int32_t on) {
  if (!ctx) return ES_E_BADARG;
  ctx->profiling = on != 0;
  return 0;
}

int es_profile_read(es_ctx* ctx, es_profile* out, void* stream) {
  if (!ctx || !out) return ES_E_BADARG;
  CU(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  std::memset(out, 0, sizeof(*out));
  for (auto& t : ctx->timed) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, t.e0, t.e1));
    if (t.kind < 0 || t.kind >= K_NKINDS) continue;
    out->ms[t.kind] += ms;
    out->launches[t.kind] += 1;
    out->points[t.kind] += t.points;
    cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  ctx->timed.clear();
  return 0;
}

int es_debug_trace(es_ctx* ctx, int64_t* host_out, int64_t capacity_pairs) {
  // host_out == NULL: arm the trace (next chain launches record into it); else copy out [count, (clock, code)...]
  if (!ctx) return ES_E_BADARG;
  const size_t bytes = (2 + 2 * 8000) * sizeof(long long);
  if (!host_out) {
    if (!ctx->trace_dev) CU(cudaMalloc(&ctx->trace_dev, bytes));
    CU(cudaMemset(ctx->trace_dev, 0, bytes));
    return 0;
  }
  if (!ctx->trace_dev) return ES_E_BADARG;
  CU(cudaDeviceSynchronize());
  size_t n = std::min<size_t>(bytes, (2 + 2 * static_cast<size_t>(capacity_pairs)) * sizeof(long long));
  CU(cudaMemcpy(host_out, ctx->trace_dev, n, cudaMemcpyDeviceToHost));
  CU(cudaFree(ctx->trace_dev));
  ctx->trace_dev = nullptr;
  return 0;
}

int es_mma_bench(es_ctx* ctx, const int32_t* cfg15, int32_t grid, int64_t* cycles_host) {
  if (!ctx || !cfg15 || !cycles_host || grid < 1 || grid > 1024) return ES_E_BADARG;
  MmaBenchCfg c;
  std::memcpy(&c, cfg15, sizeof(c));
  long long* dev = nullptr;
  CU(cudaMalloc(&dev, grid * sizeof(long long)));
  CU(launch_mma_bench(c, grid, dev, nullptr));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(cycles_host, dev, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  CU(cudaFree(dev));
  return 0;
}

int es_umma_probe(es_ctx* ctx, const uint16_t* a, const uint16_t* b, float* d, int32_t a_lbo, int32_t a_sbo,
                  int32_t b_lbo, int32_t b_sbo, void* stream) {
  if (!ctx || !a || !b || !d) return ES_E_BADARG;
  CU(launch_umma_probe(a, b, d, a_lbo > 0 ? a_lbo : A_LBO, a_sbo > 0 ? a_sbo : A_SBO, b_lbo > 0 ? b_lbo : B_LBO,
                       b_sbo > 0 ? b_sbo : B_SBO, ctx->err_dev, static_cast<cudaStream_t>(stream)));
  ctx->launches += 2;
  return 0;
}

}  // extern "C"
