// tcgen05 issue-rate microbenchmark (tools/mma_bench.py): how many cycles one M128 x N x K16 fp16 UMMA takes when
// issued back to back from shared-memory operands, as a function of the operand layout described by the
// shared-memory descriptors.  Data values are irrelevant (timing only).
#include "es_common.cuh"
#include "es_kernels.h"

namespace es {

__global__ void __launch_bounds__(128, 1)
mma_bench_kernel(MmaBenchCfg cfg, long long* cycles_out) {
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
    if (cfg.ksteps == 4) {
#pragma unroll 1
      for (int it = 0; it < n_it; ++it) {
        const uint32_t d = tmem_base + (it & 1) * 256;
        umma_f16_ss(d, ad[0], bd[0], idesc, 1);
        umma_f16_ss(d, ad[1], bd[1], idesc, 1);
        umma_f16_ss(d, ad[2], bd[2], idesc, 1);
        umma_f16_ss(d, ad[3], bd[3], idesc, 1);
      }
    } else {
#pragma unroll 1
      for (int it = 0; it < n_it; ++it) {
        const uint32_t d = tmem_base + (it & 1) * 256;
        umma_f16_ss(d, ad[0], bd[0], idesc, 1);
        umma_f16_ss(d, ad[1], bd[1], idesc, 1);
      }
    }
    umma_commit(bar);
    int err = 0;
    mbar_wait(bar, 0, &err, 1);
    long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

cudaError_t launch_mma_bench(const MmaBenchCfg& cfg, int grid, long long* cycles_out, cudaStream_t stream) {
  const int smem = 196608 + 64;
  cudaError_t e = cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  mma_bench_kernel<<<grid, 128, smem, stream>>>(cfg, cycles_out);
  return cudaGetLastError();
}

}  // namespace es
