// Weight packing: effective fp32 weights (weight-norm already folded, reference utils.py:57-58) -> fp16 hi/lo UMMA
// operand units in the exact order the TMA producer streams them.  Runs once per parameter update (1.65 M elements).
#include "es_common.cuh"
#include "es_program.h"
#include "es_kernels.h"

namespace es {

__global__ void pack_layer_kernel(const float* __restrict__ w, int n_out, int n_in, const int* __restrict__ colmap,
                                  int k_total, float scale, uint8_t* __restrict__ units) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (sb, kg, n)
  const int n_sub = k_total / SUB_K;
  if (idx >= n_sub * 4 * HID) return;
  const int n = idx % HID;
  const int kg = (idx / HID) % 4;
  const int sb = idx / (4 * HID);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = colmap[sb * SUB_K + kg * 8 + j];
    v[j] = (col >= 0 && n < n_out) ? w[static_cast<size_t>(n) * n_in + col] * scale : 0.f;
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  uint8_t* p = units + static_cast<size_t>(2 * sb) * UNIT_BYTES + kg * B_LBO + n * 16;
  *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(p + UNIT_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void pack_layer_T_kernel(const float* __restrict__ w, int k_valid, int n_in_stride, int n_valid,
                                    float scale, uint8_t* __restrict__ units) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (sb, kg, n) over K = 256
  if (idx >= (HID / SUB_K) * 4 * HID) return;
  const int n = idx % HID;
  const int kg = (idx / HID) % 4;
  const int sb = idx / (4 * HID);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = sb * SUB_K + kg * 8 + j;
    v[j] = (k < k_valid && n < n_valid) ? w[static_cast<size_t>(k) * n_in_stride + n] * scale : 0.f;
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  uint8_t* p = units + static_cast<size_t>(2 * sb) * UNIT_BYTES + kg * B_LBO + n * 16;
  *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(p + UNIT_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// Input-adjoint operand: B[n][k] = w[k][cols[n]] * scale (k = forward out-feature < k_valid, n < n_mma; cols[n] < 0:
// zero).  K = 256 (8 sub-blocks of 32); units of n_mma * 64 bytes, [hi(sb0), lo(sb0), hi(sb1), ...], each
// [k-group 0..3][n][8 fp16] (the K-major layout of the chains' weight units with N = n_mma).
__global__ void pack_inadj_kernel(const float* __restrict__ w, int k_valid, int n_in_stride,
                                  const int* __restrict__ cols, int n_mma, float scale, uint8_t* __restrict__ units) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (sb, kg, n)
  if (idx >= (HID / SUB_K) * 4 * n_mma) return;
  const int n = idx % n_mma;
  const int kg = (idx / n_mma) % 4;
  const int sb = idx / (4 * n_mma);
  const int col = cols[n];
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = sb * SUB_K + kg * 8 + j;
    v[j] = (k < k_valid && col >= 0) ? w[static_cast<size_t>(k) * n_in_stride + col] * scale : 0.f;
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  const size_t ub = static_cast<size_t>(n_mma) * SUB_K * 2;
  uint8_t* p = units + static_cast<size_t>(2 * sb) * ub + static_cast<size_t>(kg) * n_mma * 16 + n * 16;
  *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(p + ub) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

cudaError_t launch_pack_inadj(const float* w, int k_valid, int n_in_stride, const int* cols_dev, int n_mma, float scale,
                              uint8_t* units_out, cudaStream_t stream) {
  const int total = (HID / SUB_K) * 4 * n_mma;
  pack_inadj_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w, k_valid, n_in_stride, cols_dev, n_mma, scale, units_out);
  return cudaGetLastError();
}

cudaError_t launch_pack_layer_T(const float* w, int k_valid, int n_in_stride, int n_valid, float scale,
                                uint8_t* units_out, cudaStream_t stream) {
  const int total = (HID / SUB_K) * 4 * HID;
  pack_layer_T_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w, k_valid, n_in_stride, n_valid, scale, units_out);
  return cudaGetLastError();
}

cudaError_t launch_pack_layer(const float* w, int n_out, int n_in, const int* colmap_dev, int k_total, float scale,
                              uint8_t* units_out, cudaStream_t stream) {
  const int total = (k_total / SUB_K) * 4 * HID;
  pack_layer_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w, n_out, n_in, colmap_dev, k_total, scale, units_out);
  return cudaGetLastError();
}

}  // namespace es
