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

// deterministic single-block reduction of the per-ray eikonal partials -> sum(num) / (sum(den) + 1e-6)
__global__ void eikonal_reduce_kernel(const float* __restrict__ part, long long n_rays, float* __restrict__ out) {
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
    if (threadIdx.x == 0) out[0] = n / (d + 1e-6f);
  }
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
cudaError_t launch_eikonal_reduce(const float* eik_partial, long long n_rays, float* out_scalar,
                                  cudaStream_t stream) {
  eikonal_reduce_kernel<<<1, 1024, 0, stream>>>(eik_partial, n_rays, out_scalar);
  return cudaGetLastError();
}

}  // namespace es
