// Center loss (loss.py:29-45), sharded by class like W.
//   centers_batch = gather(centers, labels)
//   loss          = mean(square(features - centers_batch))      over B*D, PRE-update centers
//   centers       = scatter_sub(centers, labels, (1 - alpha) * (centers_batch - features))
//                   (duplicate labels accumulate, each diff taken against the pre-update row)
// and the gradient the 'losses' collection entry weight*loss sends back to the features,
//   dX_i += weight * 2 (x_i - c_{y_i}) / (B*D).
//
// Two launches, HBM-bound, deterministic (no float atomics):
//   center_sort_kernel   one block sorts the (local label, row) pairs of the batch in shared memory
//                        (bitonic): rows with the same label become one contiguous segment, in row
//                        order; rows whose label another shard owns sort to the end
//   center_apply_kernel  one block per segment: the center row is read ONCE, every member row is
//                        read once (loss term, gradient, update term summed in row order), the
//                        center row is written once
// Algorithmic bytes: 3 * B * D * 4 (x in, dX read-modify-write) + 2 * D * 4 per touched center.
#include <stdint.h>

#include "../../include/asoftmax_b200.h"
#include "asm_common.cuh"

namespace asmh {

__device__ __forceinline__ long long label_at(const void* labels, int label_bytes, int i) {
  return label_bytes == 8 ? reinterpret_cast<const long long*>(labels)[i]
                          : (long long)reinterpret_cast<const int*>(labels)[i];
}

constexpr unsigned kNotOwned = 0xFFFFFFFFu;

// order[p]  = batch row at sorted position p (-1 from the first row this shard does not own)
// seglen[p] = number of rows of the segment that STARTS at p, 0 elsewhere
template <int CAP>
__global__ void __launch_bounds__(1024) center_sort_kernel(const void* __restrict__ labels, int label_bytes,
                                                           int B, int C_local, int class_offset,
                                                           int* __restrict__ order, int* __restrict__ seglen) {
  __shared__ unsigned long long key[CAP];
  for (int i = threadIdx.x; i < CAP; i += blockDim.x) {
    unsigned lab = kNotOwned;
    if (i < B) {
      const long long yl = label_at(labels, label_bytes, i) - class_offset;
      if (yl >= 0 && yl < C_local) lab = (unsigned)yl;
    }
    key[i] = ((unsigned long long)lab << 32) | (unsigned)i;
  }
  __syncthreads();
  for (int k = 2; k <= CAP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < CAP; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = key[i], b = key[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { key[i] = b; key[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int p = threadIdx.x; p < B; p += blockDim.x) {
    const unsigned lab = (unsigned)(key[p] >> 32);
    const bool owned = lab != kNotOwned;
    order[p] = owned ? (int)(key[p] & 0xFFFFFFFFu) : -1;
    int len = 0;
    if (owned && (p == 0 || (unsigned)(key[p - 1] >> 32) != lab)) {
      len = 1;
      while (p + len < B && (unsigned)(key[p + len] >> 32) == lab) ++len;
    }
    seglen[p] = len;
  }
}

// block = one segment (all rows of the batch with one local label); D / 4 float4 lanes, strided
__global__ void __launch_bounds__(128) center_apply_kernel(
    const float* __restrict__ X, int B, int D, const void* __restrict__ labels, int label_bytes,
    float* centers, int class_offset, float alpha, float weight, const int* __restrict__ order,
    const int* __restrict__ seglen, float* row_loss, unsigned int* counter, float* loss_out,
    float* dX_accum) {
  __shared__ float red[128];
  __shared__ bool is_last;
  const int p = blockIdx.x;
  const int len = seglen[p];
  if (threadIdx.x == 0 && len == 0) row_loss[p] = 0.f;           // later members of a segment, other shards' rows
  if (len > 0) {
    const int first = order[p];
    const long long yl = label_at(labels, label_bytes, first) - class_offset;
    float4* crow = reinterpret_cast<float4*>(centers + (size_t)yl * D);
    const float k = 1.0f - alpha;
    const float gscale = weight * 2.0f / ((float)B * (float)D);
    float lacc = 0.f;                                            // this thread's share of the segment's loss
    for (int d4 = threadIdx.x; d4 < D / 4; d4 += blockDim.x) {
      const float4 c = crow[d4];                                  // PRE-update center: every diff uses it
      float4 upd = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int m = 0; m < len; ++m) {                             // row order: deterministic sums
        const int row = order[p + m];
        const float4 x = __ldg(reinterpret_cast<const float4*>(X + (size_t)row * D) + d4);
        const float4 df = make_float4(x.x - c.x, x.y - c.y, x.z - c.z, x.w - c.w);
        lacc += df.x * df.x + df.y * df.y + df.z * df.z + df.w * df.w;
        if (dX_accum) {
          float4* g = reinterpret_cast<float4*>(dX_accum + (size_t)row * D) + d4;
          float4 gv = *g;
          gv.x = fmaf(gscale, df.x, gv.x); gv.y = fmaf(gscale, df.y, gv.y);
          gv.z = fmaf(gscale, df.z, gv.z); gv.w = fmaf(gscale, df.w, gv.w);
          *g = gv;
        }
        upd.x -= k * df.x; upd.y -= k * df.y; upd.z -= k * df.z; upd.w -= k * df.w;   // (1-alpha)(c - x)
      }
      crow[d4] = make_float4(c.x - upd.x, c.y - upd.y, c.z - upd.z, c.w - upd.w);
    }
    red[threadIdx.x] = lacc;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) row_loss[p] = red[0];
  }
  // fixed-order total by the last block to finish
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float tot = 0.f;
  for (int r = threadIdx.x; r < B; r += 128) tot += __ldcg(row_loss + r);
  red[threadIdx.x] = tot;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *loss_out = red[0] / ((float)B * (float)D);
    *counter = 0u;
  }
}

}  // namespace asmh

extern "C" size_t asm_center_scratch_bytes(int32_t B) { return B > 0 ? ((size_t)3 * B + 4) * 4 : 0; }

extern "C" int asm_center_loss(const float* X, int32_t B, int32_t D, const void* labels,
                               int32_t label_bytes, float* centers, int32_t C_local,
                               int32_t class_offset, float alpha, float weight, float* loss_out,
                               float* dX_accum_or_null, float* scratch, void* cuda_stream) {
  if (!X || !labels || !centers || !loss_out || !scratch || B <= 0 || D <= 0 || D % 4 != 0 || C_local <= 0 ||
      B > 4096 || (label_bytes != 4 && label_bytes != 8))
    return ASM_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  // scratch: [B] row losses | [B] sorted row order | [B] segment lengths | one uint32 ticket (zero on first use)
  float* row_loss = scratch;
  int* order = reinterpret_cast<int*>(scratch + B);
  int* seglen = reinterpret_cast<int*>(scratch + 2 * (size_t)B);
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + 3 * (size_t)B);
  if (B <= 1024) asmh::center_sort_kernel<1024><<<1, 1024, 0, st>>>(labels, label_bytes, B, C_local, class_offset, order, seglen);
  else if (B <= 2048) asmh::center_sort_kernel<2048><<<1, 1024, 0, st>>>(labels, label_bytes, B, C_local, class_offset, order, seglen);
  else asmh::center_sort_kernel<4096><<<1, 1024, 0, st>>>(labels, label_bytes, B, C_local, class_offset, order, seglen);
  asmh::center_apply_kernel<<<B, 128, 0, st>>>(X, B, D, labels, label_bytes, centers, class_offset, alpha, weight,
                                               order, seglen, row_loss, counter, loss_out, dX_accum_or_null);
  return cudaGetLastError() == cudaSuccess ? ASM_OK : ASM_ERR_CUDA;
}
