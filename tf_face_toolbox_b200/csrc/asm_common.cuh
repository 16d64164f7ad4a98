// Shared device helpers for the A-softmax head kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace asmh {

constexpr int kRowTile = 128;      // batch-row tile used for the q_part column-sum partials

// ---------------------------------------------------------------------------------------
// psi(theta) = (-1)^k cos(m theta) - 2k on the target column (SURVEY.md 8a, row a1').
// k is found by comparing t = cos(theta) against cos(l*pi/m), l = 1..m-1 (no acosf).
// Returns psi and d(psi)/dt; k / sign are piecewise constants.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void psi_eval(float t, int m, float& psi, float& dpsi) {
  t = fminf(1.0f, fmaxf(-1.0f, t));
  float T, dT;
  int k = 0;
  const float t2 = t * t;
  if (m == 4) {
    T = fmaf(8.0f * t2, t2 - 1.0f, 1.0f);            // 8t^4 - 8t^2 + 1
    dT = t * fmaf(32.0f, t2, -16.0f);                // 32t^3 - 16t
    k = (t <= 0.70710678118654752f) + (t <= 0.0f) + (t <= -0.70710678118654752f);
  } else if (m == 3) {
    T = t * fmaf(4.0f, t2, -3.0f);
    dT = fmaf(12.0f, t2, -3.0f);
    k = (t <= 0.5f) + (t <= -0.5f);
  } else if (m == 2) {
    T = fmaf(2.0f, t2, -1.0f);
    dT = 4.0f * t;
    k = (t <= 0.0f);
  } else {
    T = t;
    dT = 1.0f;
  }
  const float sgn = (k & 1) ? -1.0f : 1.0f;
  psi = fmaf(sgn, T, -2.0f * (float)k);
  dpsi = sgn * dT;
}

// f_{i,y} = (lambda * s + n * psi(s/n)) / (1 + lambda)
__device__ __forceinline__ float target_logit(float s, float n, float inv_n, int m, float lambda) {
  float psi, dpsi;
  psi_eval(s * inv_n, m, psi, dpsi);
  return (lambda * s + n * psi) / (1.0f + lambda);
}

// lambda of this step: a by-value kernel argument, or (for CUDA-graph replay, where kernel
// arguments are frozen) a device scalar the host updates between replays
__device__ __forceinline__ float step_lambda(float by_value, const float* dev) {
  return dev ? __ldg(dev) : by_value;
}

// Programmatic dependent launch (PDL).  pdl_trigger(): the next kernel in the stream may
// start launching (its CTAs become resident as SMs free up and run their prologue).
// pdl_wait(): blocks until the previous kernel has completed and its writes are visible --
// required before touching anything a predecessor produced.  Both are no-ops for a kernel
// that was launched without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Online-softmax pair combine: (m, z) <- (m, z) (+) (m2, z2)
__device__ __forceinline__ void ms_combine(float& m, float& z, float m2, float z2) {
  const float mn = fmaxf(m, m2);
  if (mn == -INFINITY) { m = mn; z = 0.f; return; }
  z = z * __expf(m - mn) + z2 * __expf(m2 - mn);
  m = mn;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace asmh
