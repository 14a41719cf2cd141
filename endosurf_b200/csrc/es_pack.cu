// Weight packing: effective fp32 weights (weight-norm already folded, reference utils.py:57-58) -> fp16 hi/lo UMMA
// operand units in the exact order the TMA producers stream them.  Runs once per parameter update (1.65 M elements):
// ONE launch per network walks a job table (forward units, transposed units for the reverse chains, the input-adjoint
// operands, and the bias / output-layer copies).
#include "es_common.cuh"
#include "es_program.h"
#include "es_kernels.h"

namespace es {

// byte offset of (k-group kg, weight row n) inside a unit of n_rows rows.  Single CTA: [kg][n_rows][8 fp16].  CTA pair:
// the two halves of the rows are contiguous blocks, [half][kg][n_rows / 2][8 fp16], because each CTA of the pair streams
// and holds only its half (the B operand of a cta_group::2 MMA is split over the pair along N).
__device__ __forceinline__ size_t unit_offset(int pair, int n_rows, int kg, int n) {
  if (!pair) return static_cast<size_t>(kg) * n_rows * 16 + n * 16;
  const int h = n_rows / 2;
  return static_cast<size_t>(n / h) * (h * 64) + static_cast<size_t>(kg) * h * 16 + (n % h) * 16;
}

// forward layer: gather columns of w [n_out, n_in] into the kernel's K order (colmap[k] = source column or -1), scale,
// zero-pad rows to 256 and K to 32 n_sub; units [hi(sub0), lo(sub0), hi(sub1), ...], each [k-group 0..3][256][8 fp16]
__device__ __forceinline__ void pack_forward(const PackJob& J, int idx) {
  const int n = idx % HID;
  const int kg = (idx / HID) % 4;
  const int sb = idx / (4 * HID);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = J.cols[sb * SUB_K + kg * 8 + j];
    v[j] = (col >= 0 && n < J.n_out) ? J.w[static_cast<size_t>(n) * J.n_in + col] * J.scale : 0.f;
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  uint8_t* p = J.units + static_cast<size_t>(2 * sb) * UNIT_BYTES + unit_offset(J.pair, HID, kg, n);
  *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(p + UNIT_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// transposed / input-adjoint operand: B[n][k] = w[k][col(n)] * scale for k < n_out (forward out-features), K = 256 (8
// sub-blocks); col(n) = cols[n] if a column table is given, else n for n < n_valid; units of n_mma * 64 bytes
__device__ __forceinline__ void pack_transposed(const PackJob& J, int idx) {
  const int n_mma = J.n_mma;
  const int n = idx % n_mma;
  const int kg = (idx / n_mma) % 4;
  const int sb = idx / (4 * n_mma);
  const int col = J.cols ? J.cols[n] : (n < J.n_valid ? n : -1);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = sb * SUB_K + kg * 8 + j;
    v[j] = (k < J.n_out && col >= 0) ? J.w[static_cast<size_t>(k) * J.n_in + col] * J.scale : 0.f;
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  const size_t ub = static_cast<size_t>(n_mma) * SUB_K * 2;
  uint8_t* p = J.units + static_cast<size_t>(2 * sb) * ub + unit_offset(J.pair, n_mma, kg, n);
  *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(p + ub) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void pack_jobs_kernel(const __grid_constant__ PackJobs jobs) {
  int j = 0;
  while (j + 1 < jobs.n && static_cast<int>(blockIdx.x) >= jobs.j[j + 1].block0) ++j;
  const PackJob& J = jobs.j[j];
  const int idx = (blockIdx.x - J.block0) * blockDim.x + threadIdx.x;
  if (idx >= J.work) return;
  if (J.kind == PACK_FORWARD) pack_forward(J, idx);
  else if (J.kind == PACK_TRANSPOSED) pack_transposed(J, idx);
  else J.dst[idx] = J.w[idx];  // PACK_COPY
}

int pack_job_work(const PackJob& J) {
  if (J.kind == PACK_FORWARD) return (J.k_total / SUB_K) * 4 * HID;
  if (J.kind == PACK_TRANSPOSED) return (HID / SUB_K) * 4 * J.n_mma;
  return J.n_copy;
}

cudaError_t launch_pack_jobs(PackJobs& jobs, cudaStream_t stream) {
  if (jobs.n <= 0) return cudaSuccess;
  int blocks = 0;
  for (int i = 0; i < jobs.n; ++i) {
    jobs.j[i].work = pack_job_work(jobs.j[i]);
    jobs.j[i].block0 = blocks;
    blocks += (jobs.j[i].work + 255) / 256;
  }
  pack_jobs_kernel<<<blocks, 256, 0, stream>>>(jobs);
  return cudaGetLastError();
}

}  // namespace es
