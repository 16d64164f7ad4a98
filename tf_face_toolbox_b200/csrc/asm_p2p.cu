// Class-sharded step over NVLink peer memory: the three collectives of the sharded head
// (all-gather X / labels, all-gather of the [3,B] softmax statistics, reduce-scatter of dX)
// done by the head's own kernels with P2P loads from the peers' symmetric buffers and
// release/acquire flags -- no NCCL launch anywhere in the step.
//
// Every rank owns one symmetric block (peer-mapped by the host, e.g. torch symmetric memory):
//   x   [2][b_max, D] fp32   this rank's embeddings, double-buffered by step parity
//   y   [2][b_max]    int32  this rank's labels
//   st  [2][3, B_max] fp32   this shard's (max, sum-exp, target logit) statistics
//   dx  [2][B_max, D] fp32   this shard's dX contribution for ALL rows
//   fl  [3][16]       u32    flags[phase][src rank] = last step that rank published
// Producers write only their OWN block and then push their step number into every peer's
// flag word (st.release.sys); consumers spin on their local flag words (ld.acquire.sys) and
// read the peers' data with L1-bypassing loads.  Parity double-buffering is sufficient
// because a rank can run at most one step ahead of its slowest peer (it needs that peer's
// step-s statistics and dX before it can finish step s).
#include <stdio.h>

#include "asm_common.cuh"
#include "asm_kernels.cuh"

namespace asmh {

namespace {
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// spin until peer `src` has published `want` for `phase`; traps after ~2 s instead of hanging
__device__ __forceinline__ void p2p_wait(const P2P& p, int phase, int src, unsigned want) {
  const unsigned* f = p.flags_local() + phase * kFlagStride + src;
  if ((int)(ld_acquire_sys(f) - want) >= 0) return;
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(f) - want) < 0) {
    if (clock64() - t0 > 4000000000ll) {
      printf("asoftmax_b200: p2p wait timeout rank %d phase %d src %d want %u have %u\n", p.rank,
             phase, src, want, ld_acquire_sys(f));
      __trap();
    }
  }
}
}  // namespace

// ---- phase 0 producer: publish this rank's embeddings / labels -------------------------
__global__ void __launch_bounds__(256) p2p_pack_kernel(P2P p, const float* X, const void* labels,
                                                       int label_bytes, int D) {
  const unsigned step = *p.step_dev + 1;          // the step being started (bumped by signal 0)
  float* xs = p.x(p.rank, step & 1);
  int* ys = p.y(p.rank, step & 1);
  const size_t n4 = (size_t)p.b_local * D / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(xs)[i] = __ldg(reinterpret_cast<const float4*>(X) + i);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.b_local; i += gridDim.x * blockDim.x)
    ys[i] = label_bytes == 8 ? (int)reinterpret_cast<const long long*>(labels)[i]
                             : reinterpret_cast<const int*>(labels)[i];
}

// ---- signal: everything this rank's earlier kernels wrote is published for `phase` ------
__global__ void p2p_signal_kernel(P2P p, int phase, int bump) {
  unsigned step = *p.step_dev;
  if (bump) {
    step += 1;
    if (threadIdx.x == 0) *p.step_dev = step;
  }
  __threadfence_system();
  if ((int)threadIdx.x < p.world)
    st_release_sys(p.flags_of(threadIdx.x) + phase * kFlagStride + p.rank, step);
}

// ---- phase 0 consumer: gather all rows of X / labels from their owners -----------------
__global__ void __launch_bounds__(256) p2p_gather_x_kernel(P2P p, float* Xg, int* yg, int D) {
  const unsigned step = *p.step_dev;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int B = p.b_local * p.world;
  if (row >= B) return;
  const int src = row / p.b_local, lr = row - src * p.b_local;
  if (lane == 0) p2p_wait(p, 0, src, step);
  __syncwarp();
  const float4* xs = reinterpret_cast<const float4*>(p.x(src, step & 1) + (size_t)lr * D);
  float4* dst = reinterpret_cast<float4*>(Xg + (size_t)row * D);
  for (int i = lane; i < D / 4; i += 32) dst[i] = __ldcv(xs + i);
  if (lane == 0) yg[row] = __ldcv(p.y(src, step & 1) + lr);
}

// ---- phase 1 consumer: gather the [3,B] statistics of every shard ----------------------
__global__ void __launch_bounds__(256) p2p_gather_stats_kernel(P2P p, float* stats_all, int B) {
  const unsigned step = *p.step_dev;
  const int g = blockIdx.y;
  if (threadIdx.x == 0) p2p_wait(p, 1, g, step);
  __syncthreads();
  const float* src = p.st(g, step & 1);
  float* dst = stats_all + (size_t)g * 3 * B;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * B; i += gridDim.x * blockDim.x)
    dst[i] = __ldcv(src + i);
}

// ---- phase 2 consumer: dX rows of this rank = sum over shards (fixed order) -------------
__global__ void __launch_bounds__(256) p2p_reduce_dx_kernel(P2P p, float* dX_local, int D) {
  const unsigned step = *p.step_dev;
  __shared__ int dummy;
  if (threadIdx.x == 0) {
    for (int g = 0; g < p.world; ++g) p2p_wait(p, 2, g, step);
    dummy = 0;
  }
  __syncthreads();
  const size_t n4 = (size_t)p.b_local * D / 4;
  const size_t row0 = (size_t)p.rank * p.b_local * D / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int g = 0; g < p.world; ++g) {
      const float4 v = __ldcv(reinterpret_cast<const float4*>(p.dx(g, step & 1)) + row0 + i);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    reinterpret_cast<float4*>(dX_local)[i] = a;
  }
}

void launch_p2p_pack(const P2P& p, const float* X, const void* labels, int label_bytes, int D,
                     cudaStream_t st) {
  int blocks = (int)(((size_t)p.b_local * D / 4 + 255) / 256);
  if (blocks > 64) blocks = 64;
  p2p_pack_kernel<<<blocks, 256, 0, st>>>(p, X, labels, label_bytes, D);
}
void launch_p2p_signal(const P2P& p, int phase, int bump, cudaStream_t st) {
  p2p_signal_kernel<<<1, 32, 0, st>>>(p, phase, bump);
}
void launch_p2p_gather_x(const P2P& p, float* Xg, int* yg, int D, cudaStream_t st) {
  const int B = p.b_local * p.world;
  p2p_gather_x_kernel<<<(B + 7) / 8, 256, 0, st>>>(p, Xg, yg, D);
}
void launch_p2p_gather_stats(const P2P& p, float* stats_all, int B, cudaStream_t st) {
  dim3 grd((3 * B + 255) / 256 > 8 ? 8 : (3 * B + 255) / 256, p.world);
  p2p_gather_stats_kernel<<<grd, 256, 0, st>>>(p, stats_all, B);
}
void launch_p2p_reduce_dx(const P2P& p, float* dX_local, int D, cudaStream_t st) {
  int blocks = (int)(((size_t)p.b_local * D / 4 + 255) / 256);
  if (blocks > 148) blocks = 148;
  p2p_reduce_dx_kernel<<<blocks, 256, 0, st>>>(p, dX_local, D);
}

}  // namespace asmh
