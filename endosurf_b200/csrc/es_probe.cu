// tcgen05 layout self-test: D[128,256] = A[128,64] * B[256,64]^T with fp16 operands staged in shared memory in the
// same canonical no-swizzle K-major layout, the same shared-memory descriptors, the same bulk-TMA weight copy and
// the same TMEM read-back as the fused MLP kernel.  The descriptor strides are runtime arguments so a test can
// confirm the (LBO, SBO) convention on real hardware.
#include "es_common.cuh"
#include "es_program.h"
#include "es_kernels.h"
#include <cstdlib>

namespace es {

constexpr int PROBE_A_BYTES = 128 * 64 * 2;  // 16 KiB
constexpr int PROBE_B_BYTES = 256 * 64 * 2;  // 32 KiB

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint16_t* __restrict__ a, const uint8_t* __restrict__ b_packed, float* __restrict__ d,
                  int a_lbo, int a_sbo, int b_lbo, int b_sbo, int* err, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sa = smem;
  uint8_t* sb = smem + PROBE_A_BYTES;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + PROBE_A_BYTES + PROBE_B_BYTES);
  uint64_t* bar_d = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_d, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // A: row r = threadIdx.x, written by generic stores exactly like the epilogue does
  {
    const int r = threadIdx.x;
    for (int kg = 0; kg < 8; ++kg) {
      uint4 v = *reinterpret_cast<const uint4*>(a + r * 64 + kg * 8);
      *reinterpret_cast<uint4*>(sa + kg * A_LBO + r * 16) = v;
    }
    fence_proxy_async_smem();
  }
  // B: already packed in global memory as two 16 KiB units [kgroup][n][8]; bulk TMA like the producer warp
  if (threadIdx.x == 0 && (mode & 1)) {
    mbar_arrive_expect_tx(bar_w, PROBE_B_BYTES);
    tma_bulk_g2s(sb, b_packed, UNIT_BYTES, bar_w);
    tma_bulk_g2s(sb + UNIT_BYTES, b_packed + UNIT_BYTES, UNIT_BYTES, bar_w);
  }
  __syncthreads();
  if (threadIdx.x == 0 && (mode & 2)) {
    if (mode & 1) mbar_wait(bar_w, 0, err, 900);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, 256);
    for (int ks = 0; ks < 4; ++ks) {
      // unit u = ks/2 holds k-groups 4u..4u+3 at stride B_LBO inside the unit
      const uint32_t b_addr = smem_u32(sb) + (ks / 2) * UNIT_BYTES + (ks % 2) * 2 * B_LBO;
      const uint32_t a_addr = smem_u32(sa) + ks * 2 * A_LBO;
      umma_f16_ss(tmem_base, make_smem_desc(a_addr, a_lbo, a_sbo), make_smem_desc(b_addr, b_lbo, b_sbo), idesc,
                   ks > 0);
    }
    umma_commit(bar_d);
  }
  if (mode & 2) mbar_wait(bar_d, 0, err, 901);
  tc_fence_after();
  for (int blk = 0; blk < 8 && (mode & 4); ++blk) {
    float v[32];
    tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + blk * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) d[(warp * 32 + lane) * 256 + blk * 32 + i] = v[i];
  }
  // mode bit 8: read back through the 16x256b fragment shape (both 16-lane halves of the warp's quadrant)
  for (int blk = 0; blk < 16 && (mode & 8); ++blk) {
    for (int half = 0; half < 2; ++half) {
      float r[8];
      tmem_ld_16x256b_x2(tmem_base + (static_cast<uint32_t>(warp * 32 + 16 * half) << 16) + blk * 16, r[0], r[1], r[2],
                         r[3], r[4], r[5], r[6], r[7]);
      tmem_ld_wait();
      const int row = warp * 32 + 16 * half + (lane >> 2), col = blk * 16 + 2 * (lane & 3);
      d[row * 256 + col] = r[0];
      d[row * 256 + col + 1] = r[1];
      d[(row + 8) * 256 + col] = r[2];
      d[(row + 8) * 256 + col + 1] = r[3];
      d[row * 256 + col + 8] = r[4];
      d[row * 256 + col + 9] = r[5];
      d[(row + 8) * 256 + col + 8] = r[6];
      d[(row + 8) * 256 + col + 9] = r[7];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

// pack B [256][64] bf16 row-major -> two units [kgroup 0..3][n][8]
__global__ void probe_pack_b(const uint16_t* __restrict__ b, uint8_t* __restrict__ out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (kg, n)
  if (idx >= 8 * 256) return;
  int kg = idx / 256, n = idx % 256;
  uint4 v = *reinterpret_cast<const uint4*>(b + n * 64 + kg * 8);
  *reinterpret_cast<uint4*>(out + (kg / 4) * UNIT_BYTES + (kg % 4) * B_LBO + n * 16) = v;
}

cudaError_t launch_umma_probe(const uint16_t* a, const uint16_t* b, float* d, int a_lbo, int a_sbo, int b_lbo,
                              int b_sbo, int* err, cudaStream_t stream) {
  const char* m = getenv("ES_PROBE_MODE");
  const int mode = m ? atoi(m) : 7;
  uint8_t* packed = nullptr;
  cudaError_t e = cudaMallocAsync(&packed, PROBE_B_BYTES, stream);
  if (e != cudaSuccess) return e;
  probe_pack_b<<<8, 256, 0, stream>>>(b, packed);
  const int smem = PROBE_A_BYTES + PROBE_B_BYTES + 64;
  e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  umma_probe_kernel<<<1, 128, smem, stream>>>(a, packed, d, a_lbo, a_sbo, b_lbo, b_sbo, err, mode);
  e = cudaGetLastError();
  cudaFreeAsync(packed, stream);
  return e;
}

}  // namespace es
