// Per-ray kernels of the EndoSurf renderer: sphere intersection + coarse sampling, hierarchical up-sampling
// (NeuS weights -> inverse CDF), sorted merge, and the NeuS alpha compositing.  One warp per ray, warp-shuffle scans
// and reductions.  These are tiny next to the MLP chains; they exist so render_rays runs without host round trips.
// Compiled with -fmad=false so every multiply/add rounds like the reference's separate PyTorch elementwise ops.
//
// Reference (relative to its repo root):
//   src/renderer/utils.py:194-210   get_sphere_intersection     -> sphere_nf()
//   src/renderer/endosurf.py:63-82  coarse z_vals               -> coarse_z_kernel
//   src/renderer/endosurf.py:221-266 up_sample + utils.py:160-191 sample_pdf(det=True) -> upsample_kernel
//   src/renderer/endosurf.py:268-287 cat_z_vals                 -> merge_z_kernel
//   src/renderer/endosurf.py:144-203 render_core (after the network) -> composite_kernel
#include "es_common.cuh"
#include "es_kernels.h"

namespace es {

constexpr int MAX_S = 256;          // max samples per ray handled by the per-ray kernels
constexpr int RAYS_PER_BLOCK = 4;   // one warp per ray

struct Ray {
  float o[3], d[3], dz[3];
};
__device__ __forceinline__ Ray load_ray(const float* rays, long long r) {
  Ray R;
  const float* p = rays + r * 9;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    R.o[i] = __ldg(p + i);
    R.d[i] = __ldg(p + 3 + i);
  }
  const float den = R.d[2] + 1e-6f;  // rays_d / (rays_d[..., 2:] + 1e-6)
#pragma unroll
  for (int i = 0; i < 3; ++i) R.dz[i] = R.d[i] / den;
  return R;
}
__device__ __forceinline__ void sphere_nf(const Ray& R, float& near, float& far) {
  const float dd = R.d[0] * R.d[0] + R.d[1] * R.d[1] + R.d[2] * R.d[2];
  const float d_o = R.d[0] * R.o[0] + R.d[1] * R.o[1] + R.d[2] * R.o[2];
  const float d1 = -d_o / dd;
  float pp = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float p = R.o[i] + d1 * R.d[i];
    pp += p * p;
  }
  const float tmp = 1.f - pp;
  const float d2 = sqrtf(fmaxf(tmp, 0.f)) / sqrtf(dd);
  near = fmaxf(d1 - d2, 0.f);
  far = d1 + d2;
}
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// exclusive prefix over lanes (sum or product)
template <bool PROD>
__device__ __forceinline__ float warp_excl_scan(float v, int lane) {
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = PROD ? inc * n : inc + n;
  }
  float ex = __shfl_up_sync(0xffffffffu, inc, 1);
  return lane == 0 ? (PROD ? 1.f : 0.f) : ex;
}

// ------------------------------------------------------------------------------------------------ coarse z
__global__ void coarse_z_kernel(const float* __restrict__ rays, long long n_rays, int n, const float* __restrict__ tv,
                                const float* __restrict__ t_rand, float sample_dist, float* __restrict__ z) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_rays * n) return;
  const long long r = idx / n;
  const int i = static_cast<int>(idx % n);
  const Ray R = load_ray(rays, r);
  float near, far;
  sphere_nf(R, near, far);
  float v = near + (far - near) * __ldg(tv + i);
  if (t_rand) v = v + __ldg(t_rand + r) * sample_dist;
  z[idx] = v;
}

// pts = o + d_z * z   (mid != 0: z is first moved to the section mid-point, endosurf.py:148-153)
__global__ void points_kernel(const float* __restrict__ rays, long long n_rays, const float* __restrict__ z, int n,
                              int mid, float sample_dist, float* __restrict__ pts) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_rays * n) return;
  const long long r = idx / n;
  const int i = static_cast<int>(idx % n);
  const Ray R = load_ray(rays, r);
  float zz = z[idx];
  if (mid) {
    const float dist = (i + 1 < n) ? (z[idx + 1] - zz) : sample_dist;
    zz = zz + dist * 0.5f;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) pts[idx * 3 + c] = R.o[c] + R.dz[c] * zz;
}

// ------------------------------------------------------------------------------------------------ up-sampling
__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK)
upsample_kernel(const float* __restrict__ rays, long long n_rays, const float* __restrict__ z_in,
                const float* __restrict__ sdf_in, int n, int n_imp, const float* __restrict__ u_vals, float inv_s,
                float* __restrict__ new_z) {
  __shared__ float s_z[RAYS_PER_BLOCK][MAX_S];
  __shared__ float s_a[RAYS_PER_BLOCK][MAX_S];    // alpha, then weights / pdf
  __shared__ float s_cdf[RAYS_PER_BLOCK][MAX_S];  // cdf (n entries)
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * RAYS_PER_BLOCK + w;
  if (r >= n_rays) return;
  const Ray R = load_ray(rays, r);
  float* z = s_z[w];
  float* a = s_a[w];
  float* cdf = s_cdf[w];
  const float* sdf = sdf_in + r * n;
  for (int i = lane; i < n; i += 32) z[i] = z_in[r * n + i];
  __syncwarp();
  // alpha per section i in [0, n-1)
  for (int i = lane; i < n - 1; i += 32) {
    const float z0 = z[i], z1 = z[i + 1];
    const float s0 = __ldg(sdf + i), s1 = __ldg(sdf + i + 1);
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float p0 = R.o[c] + R.dz[c] * z0, p1 = R.o[c] + R.dz[c] * z1;
      r0 += p0 * p0;
      r1 += p1 * p1;
    }
    const bool inside = (sqrtf(r0) < 1.f) || (sqrtf(r1) < 1.f);
    const float cosv = (s1 - s0) / (z1 - z0 + 1e-6f);
    float prev_cos = 0.f;
    if (i > 0) prev_cos = (s0 - __ldg(sdf + i - 1)) / (z0 - z[i - 1] + 1e-6f);
    float cv = fminf(prev_cos, cosv);
    cv = fminf(fmaxf(cv, -1e3f), 0.f) * (inside ? 1.f : 0.f);
    const float dist = z1 - z0;
    const float mid = (s0 + s1) * 0.5f;
    const float pe = mid - cv * dist * 0.5f;
    const float ne = mid + cv * dist * 0.5f;
    const float pc = sigmoidf_acc(pe * inv_s), nc = sigmoidf_acc(ne * inv_s);
    a[i] = (pc - nc + 1e-6f) / (pc + 1e-6f);
  }
  __syncwarp();
  // weights = alpha * exclusive_cumprod(1 - alpha + 1e-7): each lane owns a contiguous segment
  const int m = n - 1;
  const int seg = (m + 31) / 32;
  const int b0 = min(lane * seg, m), b1 = min(b0 + seg, m);
  float prod = 1.f;
  for (int i = b0; i < b1; ++i) prod *= (1.f - a[i] + 1e-7f);
  float T = warp_excl_scan<true>(prod, lane);
  float wsum = 0.f;
  for (int i = b0; i < b1; ++i) {
    const float al = a[i];
    const float wgt = al * T + 1e-5f;  // sample_pdf: weights + 1e-5
    T *= (1.f - al + 1e-7f);
    a[i] = wgt;
    wsum += wgt;
  }
  const float total = warp_sum(wsum);
  // cdf = [0, cumsum(w / total)]
  float part = 0.f;
  for (int i = b0; i < b1; ++i) part += a[i] / total;
  float run = warp_excl_scan<false>(part, lane);
  for (int i = b0; i < b1; ++i) {
    run += a[i] / total;
    cdf[i + 1] = run;
  }
  if (lane == 0) cdf[0] = 0.f;
  __syncwarp();
  // inverse CDF at u_k: inds = searchsorted(cdf, u, right=True) = #entries <= u
  for (int k = lane; k < n_imp; k += 32) {
    const float u = __ldg(u_vals + k);
    int lo = 0, hi = n;  // first index with cdf[idx] > u
    while (lo < hi) {
      const int mid_i = (lo + hi) >> 1;
      if (cdf[mid_i] <= u) lo = mid_i + 1;
      else hi = mid_i;
    }
    const int below = max(lo - 1, 0);
    const int above = min(lo, n - 1);
    float denom = cdf[above] - cdf[below];
    if (denom < 1e-5f) denom = 1.f;
    const float tt = (u - cdf[below]) / denom;
    new_z[r * n_imp + k] = z[below] + tt * (z[above] - z[below]);
  }
}

// ------------------------------------------------------------------------------------------------ merge (stable rank sort)
__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK)
merge_z_kernel(long long n_rays, const float* __restrict__ z, const float* __restrict__ sdf, int n,
               const float* __restrict__ new_z, const float* __restrict__ new_sdf, int m, float* __restrict__ z_out,
               float* __restrict__ sdf_out) {
  __shared__ float s_z[RAYS_PER_BLOCK][MAX_S];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * RAYS_PER_BLOCK + w;
  if (r >= n_rays) return;
  const int tot = n + m;
  float* a = s_z[w];
  for (int i = lane; i < n; i += 32) a[i] = z[r * n + i];
  for (int i = lane; i < m; i += 32) a[n + i] = new_z[r * m + i];
  __syncwarp();
  for (int e = lane; e < tot; e += 32) {
    const float v = a[e];
    int rank = 0;
    for (int j = 0; j < tot; ++j) {
      const float o = a[j];
      rank += (o < v) || (o == v && j < e);
    }
    z_out[r * tot + rank] = v;
    if (new_sdf) sdf_out[r * tot + rank] = (e < n) ? sdf[r * n + e] : new_sdf[r * m + (e - n)];
  }
}

// ------------------------------------------------------------------------------------------------ compositing
__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK)
composite_kernel(const float* __restrict__ rays, long long n_rays, const float* __restrict__ z_in, int m,
                 float sample_dist, const float* __restrict__ sdf, const float* __restrict__ g_c,
                 const float* __restrict__ jac, const float* __restrict__ rgb, const float* __restrict__ variance,
                 float cos_anneal, CompositeOut out) {
  __shared__ float s_a[RAYS_PER_BLOCK][MAX_S];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * RAYS_PER_BLOCK + w;
  if (r >= n_rays) return;
  const Ray R = load_ray(rays, r);
  // SingleVarianceNetwork (endosurf.py:850-852) + clip (endosurf.py:168)
  const float inv_s = fminf(fmaxf(expf(__ldg(variance) * 10.f), 1e-6f), 1e6f);
  float* al = s_a[w];
  const float* z = z_in + r * m;
  float eik_num = 0.f, eik_den = 0.f;
  for (int i = lane; i < m; i += 32) {
    const long long p = r * m + i;
    const float z0 = __ldg(z + i);
    const float dist = (i + 1 < m) ? (__ldg(z + i + 1) - z0) : sample_dist;
    const float mid = z0 + dist * 0.5f;
    float pn = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float q = R.o[c] + R.dz[c] * mid;
      pn += q * q;
    }
    // g_o = J^T g_c   (identity proven against autograd in SURVEY 7.4; J[i][j] = d x_c_i / d x_j)
    float gc[3] = {__ldg(g_c + p * 3), __ldg(g_c + p * 3 + 1), __ldg(g_c + p * 3 + 2)};
    float go[3];
    if (jac) {
      const float* J = jac + p * 9;
#pragma unroll
      for (int j = 0; j < 3; ++j) go[j] = __ldg(J + j) * gc[0] + __ldg(J + 3 + j) * gc[1] + __ldg(J + 6 + j) * gc[2];
    } else {
#pragma unroll
      for (int j = 0; j < 3; ++j) go[j] = gc[j];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) out.gradients_o[p * 3 + j] = go[j];
    const float true_cos = R.d[0] * go[0] + R.d[1] * go[1] + R.d[2] * go[2];
    const float iter_cos =
        -(fmaxf(-true_cos * 0.5f + 0.5f, 0.f) * (1.f - cos_anneal) + fmaxf(-true_cos, 0.f) * cos_anneal);
    const float s = __ldg(sdf + p);
    const float e_next = s + iter_cos * dist * 0.5f;
    const float e_prev = s - iter_cos * dist * 0.5f;
    const float pc = sigmoidf_acc(e_prev * inv_s), nc = sigmoidf_acc(e_next * inv_s);
    const float alpha = fminf(fmaxf((pc - nc + 1e-6f) / (pc + 1e-6f), 0.f), 1.f);
    al[i] = alpha;
    out.cdf[p] = pc;
    const float gn = sqrtf(go[0] * go[0] + go[1] * go[1] + go[2] * go[2]) - 1.f;
    const float relax = sqrtf(pn) < 1.2f ? 1.f : 0.f;
    eik_num += relax * (gn * gn);
    eik_den += relax;
  }
  __syncwarp();
  // weights = alpha * exclusive_cumprod(1 - alpha + 1e-7)
  const int seg = (m + 31) / 32;
  const int b0 = min(lane * seg, m), b1 = min(b0 + seg, m);
  float prod = 1.f;
  for (int i = b0; i < b1; ++i) prod *= (1.f - al[i] + 1e-7f);
  float T = warp_excl_scan<true>(prod, lane);
  float c0 = 0.f, c1 = 0.f, c2 = 0.f, dep = 0.f, wmax = 0.f;
  for (int i = b0; i < b1; ++i) {
    const long long p = r * m + i;
    const float a = al[i];
    const float wgt = a * T;
    T *= (1.f - a + 1e-7f);
    out.weights[p] = wgt;
    const float z0 = __ldg(z + i);
    const float dist = (i + 1 < m) ? (__ldg(z + i + 1) - z0) : sample_dist;
    dep += wgt * (z0 + dist * 0.5f);
    c0 += __ldg(rgb + p * 3) * wgt;
    c1 += __ldg(rgb + p * 3 + 1) * wgt;
    c2 += __ldg(rgb + p * 3 + 2) * wgt;
    wmax = fmaxf(wmax, wgt);
  }
  c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
  dep = warp_sum(dep);
  wmax = warp_max(wmax);
  eik_num = warp_sum(eik_num);
  eik_den = warp_sum(eik_den);
  if (lane == 0) {
    out.color_map[r * 3] = c0;
    out.color_map[r * 3 + 1] = c1;
    out.color_map[r * 3 + 2] = c2;
    out.depth_map[r] = dep;
    out.weight_max[r] = wmax;
    out.eik_partial[r * 2] = eik_num;
    out.eik_partial[r * 2 + 1] = eik_den;
    out.s_val[r] = 1.f / inv_s;
  }
}

// ------------------------------------------------------------------------------------------------ compositing backward
// Reverse of composite_kernel for the training path (the reference lets autograd differentiate endosurf.py:168-203).
// One warp per ray.  Inputs: what the forward saw (z, sdf, g_c, J, rgb, variance) and the adjoints of the outputs;
// outputs: the adjoint rows the three reverse MLP chains start from, and d loss / d inv_s per ray.
//   w_i = alpha_i T_i, T_i = prod_{k<i} x_k, x_k = 1 - alpha_k + 1e-7:
//   alpha_bar_i = wbar_i T_i - (sum_{k>i} wbar_k w_k) / x_i            (closed-form cumprod backward, x_k >= 1e-7)
__global__ void __launch_bounds__(32 * RAYS_PER_BLOCK)
composite_bwd_kernel(const float* __restrict__ rays, long long n_rays, const float* __restrict__ z_in, int m,
                     float sample_dist, const float* __restrict__ sdf, const float* __restrict__ g_c,
                     const float* __restrict__ jac, const float* __restrict__ rgb, const float* __restrict__ variance,
                     float cos_anneal, CompositeBwd b) {
  __shared__ float s_al[RAYS_PER_BLOCK][MAX_S];   // alpha
  __shared__ float s_T[RAYS_PER_BLOCK][MAX_S];    // transmittance before sample i
  __shared__ float s_w[RAYS_PER_BLOCK][MAX_S];    // weights
  __shared__ float s_ab[RAYS_PER_BLOCK][MAX_S];   // wbar, then alpha_bar
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * RAYS_PER_BLOCK + w;
  if (r >= n_rays) return;
  const Ray R = load_ray(rays, r);
  const float raw_s = expf(__ldg(variance) * 10.f);
  const float inv_s = fminf(fmaxf(raw_s, 1e-6f), 1e6f);
  float* al = s_al[w];
  float* sT = s_T[w];
  float* sw = s_w[w];
  float* sab = s_ab[w];
  const float* z = z_in + r * m;
  const float cb[3] = {b.color_bar ? __ldg(b.color_bar + r * 3) : 0.f, b.color_bar ? __ldg(b.color_bar + r * 3 + 1) : 0.f,
                       b.color_bar ? __ldg(b.color_bar + r * 3 + 2) : 0.f};
  const float db = b.depth_bar ? __ldg(b.depth_bar + r) : 0.f;
  const float eb = b.eik_bar ? __ldg(b.eik_bar) / __ldg(b.eik_den) : 0.f;  // d loss / d (sum relax (|g|-1)^2)

  // per-sample forward quantities (same arithmetic as composite_kernel)
  auto sample = [&](int i, float& dist, float& mid, float (&go)[3], float& tc, float& pc, float& nc, float& e_prev,
                    float& e_next, float& ratio, float& relax) {
    const long long p = r * m + i;
    const float z0 = __ldg(z + i);
    dist = (i + 1 < m) ? (__ldg(z + i + 1) - z0) : sample_dist;
    mid = z0 + dist * 0.5f;
    float pn = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float q = R.o[c] + R.dz[c] * mid;
      pn += q * q;
    }
    relax = sqrtf(pn) < 1.2f ? 1.f : 0.f;
    const float gc[3] = {__ldg(g_c + p * 3), __ldg(g_c + p * 3 + 1), __ldg(g_c + p * 3 + 2)};
    if (jac) {
      const float* J = jac + p * 9;
#pragma unroll
      for (int j = 0; j < 3; ++j) go[j] = __ldg(J + j) * gc[0] + __ldg(J + 3 + j) * gc[1] + __ldg(J + 6 + j) * gc[2];
    } else {
#pragma unroll
      for (int j = 0; j < 3; ++j) go[j] = gc[j];
    }
    tc = R.d[0] * go[0] + R.d[1] * go[1] + R.d[2] * go[2];
    const float iter_cos = -(fmaxf(-tc * 0.5f + 0.5f, 0.f) * (1.f - cos_anneal) + fmaxf(-tc, 0.f) * cos_anneal);
    const float sv = __ldg(sdf + p);
    e_next = sv + iter_cos * dist * 0.5f;
    e_prev = sv - iter_cos * dist * 0.5f;
    pc = sigmoidf_acc(e_prev * inv_s);
    nc = sigmoidf_acc(e_next * inv_s);
    ratio = (pc - nc + 1e-6f) / (pc + 1e-6f);
  };

  for (int i = lane; i < m; i += 32) {
    float dist, mid, go[3], tc, pc, nc, ep, en, ratio, relax;
    sample(i, dist, mid, go, tc, pc, nc, ep, en, ratio, relax);
    al[i] = fminf(fmaxf(ratio, 0.f), 1.f);
  }
  __syncwarp();
  const int seg = (m + 31) / 32;
  const int b0 = min(lane * seg, m), b1 = min(b0 + seg, m);
  float prod = 1.f;
  for (int i = b0; i < b1; ++i) prod *= (1.f - al[i] + 1e-7f);
  float T = warp_excl_scan<true>(prod, lane);
  float seg_sum = 0.f;
  for (int i = b0; i < b1; ++i) {
    const long long p = r * m + i;
    const float a = al[i];
    const float wgt = a * T;
    sT[i] = T;
    sw[i] = wgt;
    T *= (1.f - a + 1e-7f);
    const float z0 = __ldg(z + i);
    const float dist = (i + 1 < m) ? (__ldg(z + i + 1) - z0) : sample_dist;
    float wb = b.weights_bar ? __ldg(b.weights_bar + p) : 0.f;
    wb += db * (z0 + dist * 0.5f);
    wb += cb[0] * __ldg(rgb + p * 3) + cb[1] * __ldg(rgb + p * 3 + 1) + cb[2] * __ldg(rgb + p * 3 + 2);
    sab[i] = wb;
    seg_sum += wb * wgt;
  }
  // suffix sums over lanes: S = sum of wbar_k w_k over all samples after this lane's segment
  const float total = warp_sum(seg_sum);
  float S = total - (warp_excl_scan<false>(seg_sum, lane) + seg_sum);
  for (int i = b1 - 1; i >= b0; --i) {
    const float wb = sab[i];
    sab[i] = wb * sT[i] - S / (1.f - al[i] + 1e-7f);
    S += wb * sw[i];
  }
  __syncwarp();

  float invs_bar = 0.f;
  for (int i = lane; i < m; i += 32) {
    const long long p = r * m + i;
    float dist, mid, go[3], tc, pc, nc, ep, en, ratio, relax;
    sample(i, dist, mid, go, tc, pc, nc, ep, en, ratio, relax);
    // colour
    const float wgt = sw[i];
    float4 oc = make_float4(0.f, 0.f, 0.f, 0.f);
    {
      const float c0 = __ldg(rgb + p * 3), c1 = __ldg(rgb + p * 3 + 1), c2 = __ldg(rgb + p * 3 + 2);
      float rb0 = cb[0] * wgt, rb1 = cb[1] * wgt, rb2 = cb[2] * wgt;
      if (b.rgb_bar) {
        rb0 += __ldg(b.rgb_bar + p * 3);
        rb1 += __ldg(b.rgb_bar + p * 3 + 1);
        rb2 += __ldg(b.rgb_bar + p * 3 + 2);
      }
      oc.x = rb0 * c0 * (1.f - c0);  // through the output sigmoid (endosurf.py:841)
      oc.y = rb1 * c1 * (1.f - c1);
      oc.z = rb2 * c2 * (1.f - c2);
    }
    reinterpret_cast<float4*>(b.adj_color)[p] = oc;
    // alpha -> cdfs -> sdf, cos
    const float ratio_bar = (ratio >= 0.f && ratio <= 1.f) ? sab[i] : 0.f;
    const float den = pc + 1e-6f;
    float pc_bar = ratio_bar * nc / (den * den);
    if (b.cdf_bar) pc_bar += __ldg(b.cdf_bar + p);
    const float nc_bar = -ratio_bar / den;
    const float ap = pc_bar * pc * (1.f - pc);  // adjoint of (e_prev * inv_s)
    const float an = nc_bar * nc * (1.f - nc);
    invs_bar += ap * ep + an * en;
    const float ep_bar = ap * inv_s, en_bar = an * inv_s;
    float s_bar = ep_bar + en_bar;
    if (b.sdf_bar) s_bar += __ldg(b.sdf_bar + p);
    const float ic_bar = (en_bar - ep_bar) * dist * 0.5f;
    float tc_bar = 0.f;
    if (-tc * 0.5f + 0.5f > 0.f) tc_bar += 0.5f * (1.f - cos_anneal) * ic_bar;
    if (-tc > 0.f) tc_bar += cos_anneal * ic_bar;
    float gob[3];
    const float gn = sqrtf(go[0] * go[0] + go[1] * go[1] + go[2] * go[2]);
    const float ek = (gn > 0.f) ? eb * relax * 2.f * (gn - 1.f) / gn : 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      gob[j] = tc_bar * R.d[j] + ek * go[j];
      if (b.go_bar) gob[j] += __ldg(b.go_bar + p * 3 + j);
    }
    // g_o = J^T g_c:  gc_bar_i = sum_j J[i][j] gobar_j ;  Jbar[i][j] = g_c[i] gobar[j]
    float4* as = reinterpret_cast<float4*>(b.adj_sdf) + p * 4;
    as[0] = make_float4(0.f, 0.f, 0.f, s_bar);
    const float gc[3] = {__ldg(g_c + p * 3), __ldg(g_c + p * 3 + 1), __ldg(g_c + p * 3 + 2)};
    if (jac) {
      const float* J = jac + p * 9;
#pragma unroll
      for (int i2 = 0; i2 < 3; ++i2)
        as[1 + i2] = make_float4(0.f, 0.f, 0.f, __ldg(J + 3 * i2) * gob[0] + __ldg(J + 3 * i2 + 1) * gob[1] +
                                                    __ldg(J + 3 * i2 + 2) * gob[2]);
      if (b.adj_deform) {
        float4* ad = reinterpret_cast<float4*>(b.adj_deform) + p * 4;
        ad[0] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 3; ++j) ad[1 + j] = make_float4(gc[0] * gob[j], gc[1] * gob[j], gc[2] * gob[j], 0.f);
      }
    } else {
#pragma unroll
      for (int i2 = 0; i2 < 3; ++i2) as[1 + i2] = make_float4(0.f, 0.f, 0.f, gob[i2]);
    }
  }
  invs_bar = warp_sum(invs_bar);
  if (lane == 0) b.invs_partial[r] = (raw_s >= 1e-6f && raw_s <= 1e6f) ? invs_bar * 10.f * inv_s : 0.f;
}

// deterministic single-block reduction of the per-ray eikonal partials -> sum(num) / (sum(den) + 1e-6)
__global__ void eikonal_reduce_kernel(const float* __restrict__ part, long long n_rays, float* __restrict__ out,
                                      float* __restrict__ out_den) {
  __shared__ float s_n[32], s_d[32];
  float n = 0.f, d = 0.f;
  for (long long i = threadIdx.x; i < n_rays; i += blockDim.x) {
    n += part[2 * i];
    d += part[2 * i + 1];
  }
  n = warp_sum(n);
  d = warp_sum(d);
  if ((threadIdx.x & 31) == 0) {
    s_n[threadIdx.x >> 5] = n;
    s_d[threadIdx.x >> 5] = d;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    n = threadIdx.x < (blockDim.x >> 5) ? s_n[threadIdx.x] : 0.f;
    d = threadIdx.x < (blockDim.x >> 5) ? s_d[threadIdx.x] : 0.f;
    n = warp_sum(n);
    d = warp_sum(d);
    if (threadIdx.x == 0) {
      out[0] = n / (d + 1e-6f);
      if (out_den) out_den[0] = d + 1e-6f;
    }
  }
}
// out[0] (+)= sum of part[0..n)   (deterministic, single block)
__global__ void sum_reduce_kernel(const float* __restrict__ part, long long n, float* __restrict__ out, int accumulate) {
  __shared__ float s_n[32];
  float v = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) v += part[i];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_n[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? s_n[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = accumulate ? out[0] + v : v;
  }
}

// point-field training calls without compositing: the adjoint rows of the reverse chains straight from the
// adjoints of (sdf, g_c, J, rgb) at explicit points
__global__ void point_adjoints_kernel(long long n, const float* __restrict__ rgb, const float* __restrict__ sdf_bar,
                                      const float* __restrict__ gc_bar, const float* __restrict__ jac_bar,
                                      const float* __restrict__ rgb_bar, float* __restrict__ adj_color,
                                      float* __restrict__ adj_sdf, float* __restrict__ adj_deform) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float4 oc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rgb_bar) {
    const float c0 = rgb[p * 3], c1 = rgb[p * 3 + 1], c2 = rgb[p * 3 + 2];
    oc.x = rgb_bar[p * 3] * c0 * (1.f - c0);
    oc.y = rgb_bar[p * 3 + 1] * c1 * (1.f - c1);
    oc.z = rgb_bar[p * 3 + 2] * c2 * (1.f - c2);
  }
  reinterpret_cast<float4*>(adj_color)[p] = oc;
  float4* as = reinterpret_cast<float4*>(adj_sdf) + p * 4;
  as[0] = make_float4(0.f, 0.f, 0.f, sdf_bar ? sdf_bar[p] : 0.f);
  for (int j = 0; j < 3; ++j) as[1 + j] = make_float4(0.f, 0.f, 0.f, gc_bar ? gc_bar[p * 3 + j] : 0.f);
  if (adj_deform) {
    float4* ad = reinterpret_cast<float4*>(adj_deform) + p * 4;
    ad[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < 3; ++j)
      ad[1 + j] = jac_bar ? make_float4(jac_bar[p * 9 + j], jac_bar[p * 9 + 3 + j], jac_bar[p * 9 + 6 + j], 0.f)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// points of a regular grid slab [x0, x0 + nx) x res x res in the order of torch.meshgrid(indexing="ij") flattened;
// coordinates = torch.linspace(lo, hi, res) (symmetric evaluation: start + i step below the middle, end - (res-1-i) step
// above it), reference utils.py:139-157
__global__ void grid_points_kernel(float3 lo, float3 hi, int res, int x0, int nx, float* __restrict__ pts) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(nx) * res * res;
  if (idx >= total) return;
  const int iz = static_cast<int>(idx % res);
  const int iy = static_cast<int>((idx / res) % res);
  const int ix = x0 + static_cast<int>(idx / (static_cast<long long>(res) * res));
  auto lin = [res](float a, float b, int i) {
    const float step = (b - a) / static_cast<float>(res - 1);
    return i < res / 2 ? a + step * static_cast<float>(i) : b - step * static_cast<float>(res - 1 - i);
  };
  pts[idx * 3] = lin(lo.x, hi.x, ix);
  pts[idx * 3 + 1] = lin(lo.y, hi.y, iy);
  pts[idx * 3 + 2] = lin(lo.z, hi.z, iz);
}

__global__ void fill_identity_jac_kernel(float* __restrict__ jac, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n * 9) return;
  const int e = static_cast<int>(i % 9);
  jac[i] = (e == 0 || e == 4 || e == 8) ? 1.f : 0.f;
}

// ------------------------------------------------------------------------------------------------ launchers
static inline int blocks_for(long long n, int per) { return static_cast<int>((n + per - 1) / per); }

cudaError_t launch_coarse_z(const RayGeom& rg, int n, const float* t_vals, const float* t_rand, float sample_dist,
                            float* z, cudaStream_t stream) {
  if (rg.n_rays == 0) return cudaSuccess;
  coarse_z_kernel<<<blocks_for(rg.n_rays * n, 256), 256, 0, stream>>>(rg.rays, rg.n_rays, n, t_vals, t_rand,
                                                                      sample_dist, z);
  return cudaGetLastError();
}
cudaError_t launch_points_from_z(const RayGeom& rg, const float* z, int n, int mid, float sample_dist, float* pts,
                                 cudaStream_t stream) {
  if (rg.n_rays == 0) return cudaSuccess;
  points_kernel<<<blocks_for(rg.n_rays * n, 256), 256, 0, stream>>>(rg.rays, rg.n_rays, z, n, mid, sample_dist, pts);
  return cudaGetLastError();
}
cudaError_t launch_upsample(const RayGeom& rg, const float* z, const float* sdf, int n, int n_imp,
                            const float* u_vals, float inv_s, float* new_z, cudaStream_t stream) {
  if (rg.n_rays == 0) return cudaSuccess;
  if (n > MAX_S || n < 2) return cudaErrorInvalidValue;
  upsample_kernel<<<blocks_for(rg.n_rays, RAYS_PER_BLOCK), 32 * RAYS_PER_BLOCK, 0, stream>>>(
      rg.rays, rg.n_rays, z, sdf, n, n_imp, u_vals, inv_s, new_z);
  return cudaGetLastError();
}
cudaError_t launch_merge_z(long long n_rays, const float* z, const float* sdf, int n, const float* new_z,
                           const float* new_sdf, int m, float* z_out, float* sdf_out, cudaStream_t stream) {
  if (n_rays == 0) return cudaSuccess;
  if (n + m > MAX_S) return cudaErrorInvalidValue;
  merge_z_kernel<<<blocks_for(n_rays, RAYS_PER_BLOCK), 32 * RAYS_PER_BLOCK, 0, stream>>>(n_rays, z, sdf, n, new_z,
                                                                                         new_sdf, m, z_out, sdf_out);
  return cudaGetLastError();
}
cudaError_t launch_composite(const RayGeom& rg, const float* z, int m, float sample_dist, const float* sdf,
                             const float* g_c, const float* jac, const float* rgb, const float* variance,
                             float cos_anneal, const CompositeOut& out, cudaStream_t stream) {
  if (rg.n_rays == 0) return cudaSuccess;
  if (m > MAX_S) return cudaErrorInvalidValue;
  composite_kernel<<<blocks_for(rg.n_rays, RAYS_PER_BLOCK), 32 * RAYS_PER_BLOCK, 0, stream>>>(
      rg.rays, rg.n_rays, z, m, sample_dist, sdf, g_c, jac, rgb, variance, cos_anneal, out);
  return cudaGetLastError();
}
cudaError_t launch_eikonal_reduce(const float* eik_partial, long long n_rays, float* out_scalar, float* out_den,
                                  cudaStream_t stream) {
  eikonal_reduce_kernel<<<1, 1024, 0, stream>>>(eik_partial, n_rays, out_scalar, out_den);
  return cudaGetLastError();
}
cudaError_t launch_composite_bwd(const RayGeom& rg, const float* z, int m, float sample_dist, const float* sdf,
                                 const float* g_c, const float* jac, const float* rgb, const float* variance,
                                 float cos_anneal, const CompositeBwd& b, cudaStream_t stream) {
  if (rg.n_rays == 0) return cudaSuccess;
  if (m > MAX_S) return cudaErrorInvalidValue;
  composite_bwd_kernel<<<blocks_for(rg.n_rays, RAYS_PER_BLOCK), 32 * RAYS_PER_BLOCK, 0, stream>>>(
      rg.rays, rg.n_rays, z, m, sample_dist, sdf, g_c, jac, rgb, variance, cos_anneal, b);
  return cudaGetLastError();
}
cudaError_t launch_sum_reduce(const float* part, long long n, float* out, int accumulate, cudaStream_t stream) {
  sum_reduce_kernel<<<1, 1024, 0, stream>>>(part, n, out, accumulate);
  return cudaGetLastError();
}
cudaError_t launch_point_adjoints(long long n, const float* rgb, const float* sdf_bar, const float* gc_bar,
                                  const float* jac_bar, const float* rgb_bar, float* adj_color, float* adj_sdf,
                                  float* adj_deform, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  point_adjoints_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(n, rgb, sdf_bar, gc_bar, jac_bar, rgb_bar, adj_color,
                                                                adj_sdf, adj_deform);
  return cudaGetLastError();
}
cudaError_t launch_grid_points(const float* lo3, const float* hi3, int res, int x0, int nx, float* pts,
                               cudaStream_t stream) {
  const long long total = static_cast<long long>(nx) * res * res;
  if (total == 0) return cudaSuccess;
  grid_points_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(make_float3(lo3[0], lo3[1], lo3[2]),
                                                                 make_float3(hi3[0], hi3[1], hi3[2]), res, x0, nx, pts);
  return cudaGetLastError();
}
cudaError_t launch_fill_identity_jac(float* jac, long long n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  fill_identity_jac_kernel<<<blocks_for(n * 9, 256), 256, 0, stream>>>(jac, n);
  return cudaGetLastError();
}

}  // namespace es
