// Center loss (loss.py:29-45), sharded by class like W.
//   centers_batch = gather(centers, labels)
//   loss          = mean(square(features - centers_batch))      over B*D, PRE-update centers
//   centers       = scatter_sub(centers, labels, (1 - alpha) * (centers_batch - features))
//                   (duplicate labels accumulate, each diff taken against the pre-update row)
// and the gradient the 'losses' collection entry weight*loss sends back to the features,
//   dX_i += weight * 2 (x_i - c_{y_i}) / (B*D).
// One block per batch row; a shard only touches rows whose label it owns.  Duplicates are
// summed in row order by the block of the FIRST occurrence, so the result is deterministic;
// the loss is summed in a fixed order by the last block to finish.
#include <stdint.h>

#include "../../include/asoftmax_b200.h"
#include "asm_common.cuh"

namespace asmh {

__device__ __forceinline__ long long label_at(const void* labels, int label_bytes, int i) {
  return label_bytes == 8 ? reinterpret_cast<const long long*>(labels)[i]
                          : (long long)reinterpret_cast<const int*>(labels)[i];
}

__global__ void __launch_bounds__(256) center_loss_kernel(
    const float* __restrict__ X, int B, int D, const void* __restrict__ labels, int label_bytes,
    float* centers, int C_local, int class_offset, float alpha, float weight, float* row_loss,
    unsigned int* counter, float* loss_out, float* dX_accum) {
  __shared__ float red[256];
  __shared__ bool is_last;
  const int i = blockIdx.x;
  const long long yl = label_at(labels, label_bytes, i) - class_offset;
  const bool owned = yl >= 0 && yl < C_local;
  float acc = 0.f;
  if (owned) {
    float* crow = centers + (size_t)yl * D;
    const float gscale = weight * 2.0f / ((float)B * (float)D);
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      const float c = crow[d];
      const float diff = X[(size_t)i * D + d] - c;
      acc = fmaf(diff, diff, acc);
      if (dX_accum) dX_accum[(size_t)i * D + d] += gscale * diff;
    }
    // this kernel only READS `centers`; the update runs in the stream-ordered second kernel
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) row_loss[i] = red[0];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float tot = 0.f;
  for (int r = threadIdx.x; r < B; r += 256) tot += __ldcg(row_loss + r);
  red[threadIdx.x] = tot;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *loss_out = red[0] / ((float)B * (float)D);
    *counter = 0u;
  }
}

// second pass: the block of the first occurrence of each label applies the summed update
__global__ void __launch_bounds__(256) center_update_kernel(
    const float* __restrict__ X, int B, int D, const void* __restrict__ labels, int label_bytes,
    float* centers, int C_local, int class_offset, float alpha) {
  const int i = blockIdx.x;
  const long long yl = label_at(labels, label_bytes, i) - class_offset;
  if (yl < 0 || yl >= C_local) return;
  for (int j = 0; j < i; ++j)
    if (label_at(labels, label_bytes, j) - class_offset == yl) return;   // not the first occurrence
  float* crow = centers + (size_t)yl * D;
  const float k = 1.0f - alpha;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float c = crow[d];
    float upd = 0.f;
    for (int j = i; j < B; ++j)                                          // row order: deterministic
      if (label_at(labels, label_bytes, j) - class_offset == yl) upd += k * (c - X[(size_t)j * D + d]);
    crow[d] = c - upd;
  }
}

}  // namespace asmh

extern "C" int asm_center_loss(const float* X, int32_t B, int32_t D, const void* labels,
                               int32_t label_bytes, float* centers, int32_t C_local,
                               int32_t class_offset, float alpha, float weight, float* loss_out,
                               float* dX_accum_or_null, float* scratch, void* cuda_stream) {
  if (!X || !labels || !centers || !loss_out || !scratch || B <= 0 || D <= 0 || C_local <= 0 ||
      (label_bytes != 4 && label_bytes != 8))
    return ASM_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  // scratch: [B] row losses followed by one uint32 ticket (must be zero on first use)
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + B);
  asmh::center_loss_kernel<<<B, 256, 0, st>>>(X, B, D, labels, label_bytes, centers, C_local,
                                              class_offset, alpha, weight, scratch, counter,
                                              loss_out, dX_accum_or_null);
  asmh::center_update_kernel<<<B, 256, 0, st>>>(X, B, D, labels, label_bytes, centers, C_local,
                                                class_offset, alpha);
  return cudaGetLastError() == cudaSuccess ? ASM_OK : ASM_ERR_CUDA;
}
