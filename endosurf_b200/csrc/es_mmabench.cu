// tcgen05 issue-rate microbenchmark (tools/mma_bench.py): how many cycles one M128 x N x K16 fp16 UMMA takes when
// issued back to back from shared-memory operands, as a function of the operand layout described by the
// shared-memory descriptors.  Data values are irrelevant (timing only).
//
// Optional background traffic (ES_MMAB_NOISE="warps,sts,delay,tma"): `warps` extra warps each issue `sts` 128-byte
// shared-memory stores then idle `delay` cycles, and (tma=1) one thread keeps 16 KiB bulk copies L2 -> smem in flight -
// the traffic the epilogue warps and the weight stream of the fused MLP kernel add next to the tensor core's operand reads.
#include "es_common.cuh"
#include "es_kernels.h"
#include <cstdlib>
#include <cstdio>

namespace es {

struct MmaNoise {
  int warps, sts, delay, tma;
  int commit_every;  // ES_MMAB_COMMIT: tcgen05.commit to a scratch mbarrier after every n MMAs (0 = only at the end)
  int ldtm;          // ES_MMAB_LDTM: noise warps also read the idle accumulator buffer (tcgen05.ld) n times per burst
  const uint8_t* gsrc;
};

__global__ void __launch_bounds__(1024, 1)
mma_bench_kernel(MmaBenchCfg cfg, MmaNoise nz, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 196608);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  // fill operands with small finite fp16 values
  for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* done = reinterpret_cast<volatile int*>(tmem_slot + 1);
  if (threadIdx.x == 0) *done = 0;
  __syncthreads();
  if (warp >= 4 && warp < 4 + nz.warps) {
    // store noise into the last 32 KiB of the A/B area's tail (not read by the MMAs: operands use < 160 KiB)
    const uint32_t base = smem_u32(smem) + 163840 + ((warp - 4) & 7) * 4096 + (threadIdx.x & 31) * 4;
    while (!*done) {
      for (int i = 0; i < nz.sts; ++i) sts32<0>(base + (i & 31) * 128, 0x3c003c00u);
      for (int i = 0; i < nz.ldtm; ++i) {
        float v[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 256 + 16 * (i & 15), v);
        tmem_ld_wait();
        if (v[0] == 123.456f) sts32<0>(base, 0u);  // keep the load alive
      }
      const long long t = clock64();
      while (clock64() - t < nz.delay) {}
    }
  } else if (warp == 3 && nz.tma && (threadIdx.x & 31) == 0) {
    // weight-stream noise: keep 2 x 16 KiB bulk copies in flight into smem [131072, 163840)
    uint64_t* nb = bar + 4;
    mbar_init(nb, 1);
    mbar_init(nb + 1, 1);
    mbar_fence_init();
    uint32_t k = 0;
    int err = 0;
    while (!*done) {
      const uint32_t st = k & 1;
      if (k >= 2) mbar_wait(nb + st, ((k >> 1) - 1) & 1, &err, 7);
      mbar_arrive_expect_tx(nb + st, 16384);
      tma_bulk_g2s(smem + 131072 + st * 16384, nz.gsrc + (k % 64) * 16384, 16384, nb + st);
      ++k;
    }
    for (uint32_t j = (k >= 2 ? k - 2 : 0); j < k; ++j) mbar_wait(nb + (j & 1), (j >> 1) & 1, &err, 8);
  }
  if (threadIdx.x == 0) {
    auto desc = [&](uint32_t addr, int layout, int lbo, int sbo) {
      uint64_t d = make_smem_desc(addr, lbo, sbo);
      d |= static_cast<uint64_t>(layout & 7) << 61;
      return d;
    };
    const uint32_t idesc = make_idesc_f16(128, cfg.n);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 65536;
    // lean issue loop: descriptors precomputed, only 64-bit adds between MMAs (cfg.ksteps <= 4)
    uint64_t ad[4], bd[4];
    for (int ks = 0; ks < 4; ++ks) {
      ad[ks] = desc(a0 + ks * cfg.a_kadv, cfg.a_layout, cfg.a_lbo, cfg.a_sbo);
      bd[ks] = desc(b0 + ks * cfg.b_kadv, cfg.b_layout, cfg.b_lbo, cfg.b_sbo);
    }
    const int n_it = cfg.iters;
    long long t0 = clock64();
    if (nz.commit_every > 0) {
      // as the fused kernel does: 2 k-steps per weight unit, a commit (stage release) after every few MMAs
      uint64_t* scratch = bar + 8;
      mbar_init(scratch, 1);
      mbar_fence_init();
      int since = 0;
#pragma unroll 1
      for (int it = 0; it < n_it * cfg.ksteps; ++it) {
        const uint32_t d = nz.ldtm ? tmem_base : tmem_base + ((it >> 2) & 1) * 256;
        umma_f16_ss(d, ad[it & 1], bd[it & 1], idesc, 1);
        if (++since == nz.commit_every) {
          umma_commit(scratch);
          since = 0;
        }
      }
    } else if (cfg.ksteps == 4) {
#pragma unroll 1
      for (int it = 0; it < n_it; ++it) {
        const uint32_t d = nz.ldtm ? tmem_base : tmem_base + (it & 1) * 256;
        umma_f16_ss(d, ad[0], bd[0], idesc, 1);
        umma_f16_ss(d, ad[1], bd[1], idesc, 1);
        umma_f16_ss(d, ad[2], bd[2], idesc, 1);
        umma_f16_ss(d, ad[3], bd[3], idesc, 1);
      }
    } else {
#pragma unroll 1
      for (int it = 0; it < n_it; ++it) {
        const uint32_t d = nz.ldtm ? tmem_base : tmem_base + (it & 1) * 256;
        umma_f16_ss(d, ad[0], bd[0], idesc, 1);
        umma_f16_ss(d, ad[1], bd[1], idesc, 1);
      }
    }
    umma_commit(bar);
    int err = 0;
    mbar_wait(bar, 0, &err, 1);
    long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
    *done = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// CTA-pair variant (ES_MMAB_PAIR=1): the leader of every 2-CTA cluster issues M256 x N x K16 cta_group::2 MMAs back to
// back; A = 128 rows per CTA, B = N/2 rows per CTA (descriptor strides from cfg, B LBO for the half-row unit).
__global__ void __launch_bounds__(128, 1)
mma_bench_pair_kernel(MmaBenchCfg cfg, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 196608);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc2<512>(tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0 && cluster_ctarank() == 0) {
    const uint32_t idesc = make_idesc_f16(256, cfg.n);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 65536;
    uint32_t ad[4], bd[4];
    for (int ks = 0; ks < 4; ++ks) {
      ad[ks] = smem_desc_lo(a0 + ks * cfg.a_kadv, cfg.a_lbo);
      bd[ks] = smem_desc_lo(b0 + ks * cfg.b_kadv, cfg.b_lbo);
    }
    constexpr uint32_t HI = smem_desc_hi(128);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < cfg.iters; ++it) {
      const uint32_t d = tmem_base + (it & 1) * 256;
      umma2_f16_ss_lo<HI, HI>(d, ad[0], bd[0], idesc, 1);
      umma2_f16_ss_lo<HI, HI>(d, ad[1], bd[1], idesc, 1);
      umma2_f16_ss_lo<HI, HI>(d, ad[2], bd[2], idesc, 1);
      umma2_f16_ss_lo<HI, HI>(d, ad[3], bd[3], idesc, 1);
    }
    umma2_commit_sa(smem_u32(bar));
    int err = 0;
    mbar_wait(bar, 0, &err, 1);
    cycles_out[blockIdx.x] = clock64() - t0;
  } else if (threadIdx.x == 0) {
    int err = 0;
    mbar_wait(bar, 0, &err, 1);  // the multicast commit reaches the peer as well
    cycles_out[blockIdx.x] = 0;
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2<512>(tmem_base);
}

cudaError_t launch_mma_bench(const MmaBenchCfg& cfg, int grid, long long* cycles_out, cudaStream_t stream) {
  const int smem = 196608 + 256;
  if (const char* v = getenv("ES_MMAB_PAIR")) {
    if (atoi(v)) {
      cudaError_t e2 = cudaFuncSetAttribute(mma_bench_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e2 != cudaSuccess) return e2;
      cudaLaunchConfig_t lc{};
      lc.gridDim = dim3(grid & ~1);
      lc.blockDim = dim3(128);
      lc.dynamicSmemBytes = smem;
      lc.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      lc.attrs = attr;
      lc.numAttrs = 1;
      return cudaLaunchKernelEx(&lc, mma_bench_pair_kernel, cfg, cycles_out);
    }
  }
  cudaError_t e = cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  MmaNoise nz{0, 0, 0, 0, 0, 0, nullptr};
  if (const char* v = getenv("ES_MMAB_NOISE")) sscanf(v, "%d,%d,%d,%d", &nz.warps, &nz.sts, &nz.delay, &nz.tma);
  if (const char* v = getenv("ES_MMAB_COMMIT")) nz.commit_every = atoi(v);
  if (const char* v = getenv("ES_MMAB_LDTM")) nz.ldtm = atoi(v);
  static uint8_t* gsrc = nullptr;
  if (nz.tma && !gsrc) {
    e = cudaMalloc(&gsrc, 64 * 16384);
    if (e != cudaSuccess) return e;
    cudaMemset(gsrc, 0, 64 * 16384);
  }
  nz.gsrc = gsrc;
  if (nz.warps > 28) nz.warps = 28;
  mma_bench_kernel<<<grid, 128 + 32 * nz.warps, smem, stream>>>(cfg, nz, cycles_out);
  return cudaGetLastError();
}

}  // namespace es
