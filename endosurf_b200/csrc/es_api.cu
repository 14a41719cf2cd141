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
  // grow-only scratch
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  ChainProg prog_sdfq{}, prog_geom{}, prog_color{};
  // reverse (training) chains: transposed weight units + programs, index = ES_NET_*
  uint8_t* rev_units[3] = {nullptr, nullptr, nullptr};
  int rev_layers[3] = {0, 0, 0};
  ChainProg prog_rev[3]{};
  // training planes: 0 (default) = write / read only the fp16 lo planes the 1-term weight-gradient path needs
  // (softplus gating, input-layer adjoints); 1 = every lo plane (3-term weight gradients)
  int full_planes = 0;
  // optional per-kernel timing (es_profile_*)
  bool profiling = false;
  long long* trace_dev = nullptr;  // debug pipeline trace buffer (es_debug_trace)
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
  ctx->prog_geom = g;
  ChainProg q = g;
  q.n_layers = n - 1;
  q.units_per_tile = static_cast<int>(ctx->sdfq_units_n);
  q.post_op = POST_SDF_TAIL;
  ctx->prog_sdfq = q;
  ChainProg c{};
  common(c);
  c.n_layers = static_cast<int>(C.progs.size());
  for (int l = 0; l < c.n_layers; ++l) c.layer[l] = C.progs[l];
  c.w_units = ctx->color_units;
  c.bias = ctx->color_bias;
  c.units_per_tile = static_cast<int>(ctx->color_units_n);
  c.post_op = POST_COLOR_TAIL;
  ctx->prog_color = c;

  // ---- reverse chains (es_point_backward): standard 4-chunk 256x256 layers on transposed weights
  const int L = cfg.n_layers;
  const int LdS = cfg.use_deform ? L - 1 : 0;  // geometry-chain slot offset of the sdf layers
  auto rev_layer = [&](LayerProg& G, uint8_t src, int stash_slot, int zbar_slot, uint8_t act, bool rank1) {
    std::memset(&G, 0, sizeof(G));
    G.n_chunks = 4;
    for (int k = 0; k < 4; ++k) {
      G.src[k] = src;
      G.arg[k] = static_cast<uint8_t>(k);
      G.nsub[k] = 2;
    }
    G.stash_slot = static_cast<uint8_t>(stash_slot);
    G.zbar_slot = static_cast<uint8_t>(zbar_slot);
    G.bwd_act = act;
    G.rank1 = rank1 ? 1 : 0;
    finish_layer(G);
  };
  for (int net = 0; net < 3; ++net) {
    ChainProg r{};
    common(r);
    r.bias = ctx->geom_bias;
    r.w_units = ctx->rev_units[net];
    r.post_op = POST_BWD_DUMP;
    r.post_zbar_slot = 0;
    int n = 0;
    if (net == ES_NET_SDF) {
      rev_layer(r.layer[n++], SRC_ADJ_FEAT, 0, 0, ACT_SOFTPLUS100, false);   // S_{L-1}^T (feature rows)
      for (int m = L - 2; m >= 1; --m)                                        // S_m^T
        rev_layer(r.layer[n++], SRC_BWD_PREV, LdS + m + 1, m, ACT_SOFTPLUS100, m == L - 2);
      r.post_stash_slot = LdS + 1;
      r.post_bwd_act = ACT_SOFTPLUS100;
    } else {
      const int tail_slot = net == ES_NET_DEFORM ? L - 1 : static_cast<int>(C.progs.size());
      for (int m = L - 2; m >= 1; --m)
        rev_layer(r.layer[n++], m == L - 2 ? SRC_BWD_OUTER3 : SRC_BWD_PREV, m == L - 2 ? tail_slot : m + 1, m,
                  ACT_RELU, false);
      r.post_stash_slot = 1;
      r.post_bwd_act = ACT_RELU;
      r.outer3_w = net == ES_NET_DEFORM ? ctx->small + SM_DEFORM_W : ctx->small + SM_COLOR_W;
    }
    r.n_layers = n;
    r.units_per_tile = n * 16;
    ctx->rev_layers[net] = n;
    ctx->prog_rev[net] = r;
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
  if (c.n_layers < 3 || 2 * (c.n_layers - 1) + 1 > MAXL) return bad("n_layers out of range (3..10)");
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
  for (int net = 0; net < 3; ++net) cudaFree(ctx->rev_units[net]);
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
  for (size_t l = 0; l < P.packs.size(); ++l) {
    const LayerPack& K = P.packs[l];
    if (!w[l] || !b[l]) return fail(ctx, ES_E_BADARG, "null layer pointer");
    CU(launch_pack_layer(w[l] + static_cast<size_t>(K.row_off) * K.n_in, K.n_out, K.n_in, K.colmap_dev, K.k_total,
                         K.scale, units + static_cast<size_t>(K.unit_off) * UNIT_BYTES, stream));
    ++ctx->launches;
    if (!(net == ES_NET_SDF && static_cast<int>(l) == L - 1))
      CU(cudaMemcpyAsync(bias + l * HID, b[l], K.n_out * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  }
  // output layers kept in fp32 for the epilogue dot products
  const float* wl = w[L - 1];
  const float* bl = b[L - 1];
  if (!wl || !bl) return fail(ctx, ES_E_BADARG, "null output layer pointer");
  if (net == ES_NET_DEFORM) {
    CU(cudaMemcpyAsync(ctx->small + SM_DEFORM_W, wl, 3 * HID * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    CU(cudaMemcpyAsync(ctx->small + SM_DEFORM_B, bl, 3 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  } else if (net == ES_NET_SDF) {
    CU(cudaMemcpyAsync(ctx->small + SM_SDF_W, wl, HID * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    CU(cudaMemcpyAsync(ctx->small + SM_SDF_B, bl, sizeof(float), cudaMemcpyDeviceToDevice, stream));
    CU(cudaMemcpyAsync(ctx->small + SM_FEAT_B, bl + 1, HID * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  } else {
    CU(cudaMemcpyAsync(ctx->small + SM_COLOR_W, wl, 3 * HID * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    CU(cudaMemcpyAsync(ctx->small + SM_COLOR_B, bl, 3 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  }
  // transposed units for the reverse (training) chains
  {
    const float inv_sqrt2 = static_cast<float>(1.0 / std::sqrt(2.0));
    const int skip = ctx->cfg.skip_layer;
    uint8_t* ru = ctx->rev_units[net];
    int r = 0;
    if (net == ES_NET_SDF) {
      CU(launch_pack_layer_T(w[L - 1] + P.in_dims[L - 1], HID, P.in_dims[L - 1], HID, 1.f,
                             ru + static_cast<size_t>(r++) * 16 * UNIT_BYTES, stream));
      ++ctx->launches;
    }
    for (int m = L - 2; m >= 1; --m) {
      CU(launch_pack_layer_T(w[m], P.out_dims[m], P.in_dims[m], P.out_dims[m - 1], m == skip ? inv_sqrt2 : 1.f,
                             ru + static_cast<size_t>(r++) * 16 * UNIT_BYTES, stream));
      ++ctx->launches;
    }
  }
  ctx->loaded[net] = true;
  return 0;
}

// Which lo planes a training launch touches (see es_ctx::full_planes).  kind: 0 geometry forward, 1 colour forward,
// 3 + net reverse chains.
static void apply_plane_mode(const es_ctx* ctx, int kind, ChainProg& p) {
  const bool full = ctx->full_planes != 0;
  const int skip = ctx->cfg.skip_layer;
  const int Ld = ctx->cfg.use_deform ? ctx->cfg.n_layers - 1 : 0;
  if (kind == 0) {
    for (int l = 0; l < p.n_layers; ++l) p.layer[l].stash_lo = (full || l >= Ld) ? 1 : 0;  // sdf slots gate softplus
    p.tail_stash_lo = full;
  } else if (kind == 1) {
    for (int l = 0; l < p.n_layers; ++l) p.layer[l].stash_lo = full;
    p.tail_stash_lo = full;
  } else if (kind >= 3) {
    const int net = kind - 3;
    const bool softplus = net == ES_NET_SDF;
    const bool input_adj = net != ES_NET_DEFORM;  // the adjoint of the network input is needed (zbar of layers 0, skip)
    for (int l = 0; l < p.n_layers; ++l) {
      p.layer[l].gate_lo = (full || softplus) ? 1 : 0;
      p.layer[l].zbar_lo = (full || (input_adj && p.layer[l].zbar_slot == skip)) ? 1 : 0;
    }
    p.post_gate_lo = (full || softplus) ? 1 : 0;
    p.post_zbar_lo = (full || input_adj) ? 1 : 0;
  }
}

static int timed_chain(es_ctx* ctx, int kind, int chain, bool tangent, const ChainProg& prog_in, const ChainIO& io,
                       cudaStream_t stream, bool bwd = false) {
  es_ctx::Timed t{kind, io.n_points, nullptr, nullptr};
  ChainProg prog = prog_in;
  if (io.stash_hi) apply_plane_mode(ctx, kind, prog);
  ChainIO io2 = io;
  io2.trace = ctx->trace_dev;
  {
    const char* f = getenv("ES_DEBUG_FLAGS");
    io2.debug_flags = f ? atoi(f) : 0;
  }
  if (ctx->profiling) {
    CU(cudaEventCreate(&t.e0));
    CU(cudaEventCreate(&t.e1));
    CU(cudaEventRecord(t.e0, stream));
  }
  CU(launch_mlp_chain(chain, tangent, ctx->cfg.use_deform != 0, prog, io2, ctx->n_sms, stream, bwd));
  ++ctx->launches;
  if (ctx->profiling) {
    CU(cudaEventRecord(t.e1, stream));
    ctx->timed.push_back(t);
  }
  return 0;
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
  return timed_chain(ctx, 2, CHAIN_SDF, false, ctx->prog_sdfq, io, static_cast<cudaStream_t>(stream));
}

static int point_forward_impl(es_ctx* ctx, const float* x, const float* t, int64_t t_div, int64_t t_stride,
                              const float* dirs, int64_t dir_div, int64_t dir_stride, int64_t n, float* x_c, float* jac,
                              float* sdf, float* g_c, float* feat, float* rgb, uint16_t* gs_hi, uint16_t* gs_lo,
                              uint16_t* cs_hi, uint16_t* cs_lo, void* stream_) {
  if (!ctx || n < 0 || t_div <= 0) return ES_E_BADARG;
  if (n == 0) return 0;
  if (!x) return ES_E_BADARG;
  const bool want_color = rgb != nullptr;
  if (want_color && (!dirs || dir_div <= 0)) return fail(ctx, ES_E_BADARG, "dirs required for rgb");
  if (ctx->cfg.use_deform && !t) return fail(ctx, ES_E_BADARG, "time pointer required with use_deform");
  if (int r = check_loaded(ctx, want_color)) return r;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // scratch for whatever the caller did not ask for but the colour chain needs
  size_t need = 0;
  {
    Carver c(nullptr);
    if (!x_c) c.take<float>(n * 3);
    if (!jac && ctx->cfg.use_deform) c.take<float>(n * 9);
    if (!g_c) c.take<float>(n * 3);
    if (!feat) c.take<float>(n * HID);
    need = c.off + 256;
  }
  if (int r = ensure_ws(ctx, need)) return r;
  Carver c(ctx->ws);
  if (!x_c) x_c = c.take<float>(n * 3);
  if (!jac && ctx->cfg.use_deform) jac = c.take<float>(n * 9);
  if (!g_c) g_c = c.take<float>(n * 3);
  if (!feat) feat = c.take<float>(n * HID);

  ChainIO io{};
  io.n_points = n;
  io.err = ctx->err_dev;
  io.x = x;
  io.t = t ? t : x;
  io.t_div = t ? t_div : 1;
  io.t_stride = t ? t_stride : 0;
  io.out_xc = x_c;
  io.out_jac = ctx->cfg.use_deform ? jac : nullptr;
  io.out_sdf = sdf;
  io.out_gc = g_c;
  io.out_feat = feat;
  io.stash_hi = gs_hi;
  io.stash_lo = gs_lo;
  io.stash_rows = ((n + 31) / 32) * TILE_ROWS;
  if (int r = timed_chain(ctx, 0, CHAIN_SDF, true, ctx->prog_geom, io, stream)) return r;
  if (!ctx->cfg.use_deform)  // canonical = observed space (endosurf.py:576-577)
    CU(cudaMemcpyAsync(x_c, x, static_cast<size_t>(n) * 3 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  if (!ctx->cfg.use_deform && jac) {
    // J = I without a deformation network (endosurf.py:626-630)
    static const float eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    std::vector<float> h(static_cast<size_t>(n) * 9);
    for (int64_t i = 0; i < n; ++i) std::memcpy(&h[i * 9], eye, sizeof(eye));
    CU(cudaMemcpyAsync(jac, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
    CU(cudaStreamSynchronize(stream));
  }
  if (want_color) {
    ChainIO ic{};
    ic.n_points = n;
    ic.err = ctx->err_dev;
    ic.x_c = x_c;
    ic.g_c = g_c;
    ic.jac = ctx->cfg.use_deform ? jac : nullptr;
    ic.dirs = dirs;
    ic.dir_div = dir_div;
    ic.dir_stride = dir_stride;
    ic.feat = feat;
    ic.out_rgb = rgb;
    ic.stash_hi = cs_hi;
    ic.stash_lo = cs_lo;
    ic.stash_rows = ((n + TILE_ROWS - 1) / TILE_ROWS) * TILE_ROWS;
    if (int r = timed_chain(ctx, 1, CHAIN_COLOR, false, ctx->prog_color, ic, stream)) return r;
  }
  return 0;
}

int es_point_forward(es_ctx* ctx, const float* x, const float* t, int64_t t_div, int64_t t_stride, const float* dirs,
                     int64_t dir_div, int64_t dir_stride, int64_t n, float* x_c, float* jac, float* sdf, float* g_c,
                     float* feat, float* rgb, void* stream) {
  return point_forward_impl(ctx, x, t, t_div, t_stride, dirs, dir_div, dir_stride, n, x_c, jac, sdf, g_c, feat, rgb,
                            nullptr, nullptr, nullptr, nullptr, stream);
}

int es_train_layout(const es_ctx* ctx, int64_t n, int64_t* out6) {
  if (!ctx || !out6 || n < 0) return ES_E_BADARG;
  out6[0] = ((n + 31) / 32) * TILE_ROWS;                       // geometry stash rows (4 rows per point)
  out6[1] = ctx->prog_geom.n_layers;                           // geometry stash slots
  out6[2] = ((n + TILE_ROWS - 1) / TILE_ROWS) * TILE_ROWS;     // colour stash rows
  out6[3] = ctx->prog_color.n_layers + 1;                      // colour stash slots (+ output-layer input)
  out6[4] = ctx->cfg.n_layers - 1;                             // zbar slots of each reverse chain (forward layer m)
  out6[5] = ctx->cfg.use_deform ? ctx->cfg.n_layers - 1 : 0;   // geometry slot offset of the sdf layers
  return 0;
}

int es_point_forward_train(es_ctx* ctx, const float* x, const float* t, int64_t t_div, int64_t t_stride,
                           const float* dirs, int64_t dir_div, int64_t dir_stride, int64_t n, float* x_c, float* jac,
                           float* sdf, float* g_c, float* feat, float* rgb, uint16_t* geom_stash_hi,
                           uint16_t* geom_stash_lo, uint16_t* color_stash_hi, uint16_t* color_stash_lo, void* stream) {
  if (!x_c || !sdf || !g_c || !feat || !geom_stash_hi || !geom_stash_lo) return ES_E_BADARG;
  if (rgb && (!color_stash_hi || !color_stash_lo)) return ES_E_BADARG;
  return point_forward_impl(ctx, x, t, t_div, t_stride, dirs, dir_div, dir_stride, n, x_c, jac, sdf, g_c, feat, rgb,
                            geom_stash_hi, geom_stash_lo, color_stash_hi, color_stash_lo, stream);
}

int es_point_backward(es_ctx* ctx, int net, int64_t n, const uint16_t* stash_hi, const uint16_t* stash_lo,
                      const float* adj, const float* adj_feat, uint16_t* zbar_hi, uint16_t* zbar_lo, void* stream) {
  if (!ctx || net < 0 || net > 2 || n < 0 || !stash_hi || !stash_lo || !adj || !zbar_hi || !zbar_lo)
    return ES_E_BADARG;
  if (net == ES_NET_SDF && !adj_feat) return fail(ctx, ES_E_BADARG, "adj_feat required for the sdf chain");
  if (net == ES_NET_DEFORM && !ctx->cfg.use_deform) return fail(ctx, ES_E_BADARG, "no deformation network");
  if (int r = check_loaded(ctx, true)) return r;
  if (n == 0) return 0;
  ChainIO io{};
  io.n_points = n;
  io.err = ctx->err_dev;
  io.stash_hi = const_cast<uint16_t*>(stash_hi);
  io.stash_lo = const_cast<uint16_t*>(stash_lo);
  const bool tangent = net != ES_NET_COLOR;
  io.stash_rows = tangent ? ((n + 31) / 32) * TILE_ROWS : ((n + TILE_ROWS - 1) / TILE_ROWS) * TILE_ROWS;
  io.zbar_hi = zbar_hi;
  io.zbar_lo = zbar_lo;
  io.adj = adj;
  io.adj_feat = adj_feat;
  io.t_div = 1;
  io.dir_div = 1;
  return timed_chain(ctx, 3 + net, tangent ? CHAIN_SDF : CHAIN_COLOR, tangent, ctx->prog_rev[net], io,
                     static_cast<cudaStream_t>(stream), true);
}

int es_set_plane_mode(es_ctx* ctx, int32_t full_planes) {
  if (!ctx) return ES_E_BADARG;
  ctx->full_planes = full_planes != 0;
  return 0;
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
    c.take<float>(PC * HID);    // feat
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
  float* feat = c.take<float>(PC * HID);
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
    if (int rr = es_point_forward(ctx, pts, tptr, n, 9, rays + r0 * 9 + 3, n, 9, R * n, x_c, deform ? jac : nullptr,
                                  sdf_dst, g_c, feat, rgb_dst, stream_))
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
  CU(launch_eikonal_reduce(eik, n_rays, out->gradient_o_error, stream));
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
