// Device-side primitives of the NVLink peer-memory transport (see asm_p2p.cu for the protocol).
#pragma once
#include <stdio.h>

#include "asm_kernels.cuh"

namespace asmh {

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// The step this rank is executing.  The counter holds the number of COMPLETED steps; the last
// kernel of a step advances it once every kernel of the step has read it, so all kernels of
// one step -- eager or replayed from a CUDA graph -- agree on the value.
__device__ __forceinline__ unsigned p2p_current_step(const P2P& p) { return *p.step_dev + 1u; }

// Spin until peer `src` has published step `want` for `phase`.  A peer that never arrives does
// not hang the GPU and does not poison the context: after p.timeout_ns (0 = wait for ever) the
// wait gives up, records (phase, src) in the handle's error word -- asm_p2p_status() reports
// it -- and the kernel runs on with whatever the peer's buffers hold.
__device__ __forceinline__ void p2p_wait(const P2P& p, int phase, int src, unsigned want) {
  const unsigned* f = p.flags_local() + phase * kFlagStride + src;
  if ((int)(ld_acquire_sys(f) - want) >= 0) return;
  const unsigned long long t0 = global_timer_ns();
  while ((int)(ld_acquire_sys(f) - want) < 0) {
    if (p.timeout_ns != 0 && global_timer_ns() - t0 > p.timeout_ns) {
      atomicCAS(p.err_dev, 0u, 0x80000000u | ((unsigned)phase << 8) | (unsigned)src);
      return;
    }
    __nanosleep(64);
  }
}

// Everything this rank wrote for `phase` of step `step` is visible to every peer afterwards.
// Called by ONE thread of the last block to finish producing (the caller has already ordered
// the other blocks' writes before its own with the usual threadfence + ticket pattern).
__device__ __forceinline__ void p2p_publish(const P2P& p, int phase, unsigned step) {
  __threadfence_system();
  for (int r = 0; r < p.world; ++r) st_release_sys(p.flags_of(r) + phase * kFlagStride + p.rank, step);
}

// "last block done" ticket: returns this block's arrival index (0 .. nblocks-1) to every thread
// of the block; the word resets itself for the next launch.
__device__ __forceinline__ unsigned block_ticket(unsigned* word, unsigned nblocks) {
  __shared__ unsigned tk;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    tk = atomicAdd(word, 1u);
    if (tk == nblocks - 1) *word = 0u;
  }
  __syncthreads();
  return tk;
}

}  // namespace asmh
