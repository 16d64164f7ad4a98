// Per-row pieces shared by the single-shard kernels (asm_prep.cu) and the kernels that fold the
// NVLink exchange into them (asm_p2p.cu).
#pragma once
#include "asm_common.cuh"
#include "asm_kernels.cuh"

namespace asmh {

// One lane per row, after the row's global (max M, sum-exp Z, target logit f_y) are known:
//   lse = M + log Z, loss_i = lse - f_y, the exp2 offset of the backward, and the target-column
//   gradient coefficients
//   g_y = (e^{f_y - lse} - 1)/B,  G'_y = g_y (lambda + psi')/(1 + lambda),
//   r   = g_y (psi - t psi') / ((1 + lambda) n)            
// A row whose label lies outside [0, C_total) (ylocal == -2) has no defined loss: it gets NaN,
// as the reference's sparse_softmax_cross_entropy does on the GPU (nets/sphere.py:109).
__device__ __forceinline__ void row_epilogue(const Step& s, int row, float m, float z, float fy,
                                             int yl_row, float tgt_f_row, float tgt_s_row,
                                             float inv_n_row) {
  const bool owned = yl_row >= 0;
  const float lse = m + logf(z);
  s.lse[row] = lse;
  const float gB = s.invB * s.gscale;            // every gradient carries 1/B and the caller's scale
  s.negoff[row] = -lse * 1.4426950408889634f + log2f(gB);
  s.rowloss[row] = yl_row == -2 ? __int_as_float(0x7fc00000) : lse - fy;
  float gt = 0.f, r = 0.f;
  if (owned) {
    float psi, dpsi;
    const float t = fminf(1.f, fmaxf(-1.f, tgt_s_row * inv_n_row));
    psi_eval(t, s.m, psi, dpsi);
    const float gy = (expf(tgt_f_row - lse) - 1.0f) * gB;
    const float lam = step_lambda(s.lambda, s.lambda_dev);
    const float il = 1.0f / (1.0f + lam);
    gt = gy * (lam + dpsi) * il;
    r = gy * (psi - t * dpsi) * il * inv_n_row;
  }
  s.gtarget[row] = gt;
  s.rcoef[row] = r;
}

// dX = sum_z dx_part[z] + r_i * x_i for the float4 elements  vb*nt + t, stepping by nb*nt
// (float4 along D; D % 4 == 0).  KS_T > 0: all KS_T partial loads are issued before the first
// add (one memory round trip); KS_T == 0: generic loop in batches of four.  Summation order
// z = 0..KS-1 either way.
template <int KS_T>
__device__ __forceinline__ void dx_finish_rows(const Step& s, float* dxo, unsigned vb, unsigned nb) {
  const size_t total4 = (size_t)s.B * s.D / 4;
  const size_t stride4 = total4;
  for (size_t i = vb * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)nb * blockDim.x) {
    const int row = (int)((i * 4) / s.D);
    const float4* p = reinterpret_cast<const float4*>(s.dx_part) + i;
    const float r = s.rcoef[row];
    float4 x;
    if (s.x_bf16) {                                   // embeddings handed over as bf16
      const uint2 h = __ldg(reinterpret_cast<const uint2*>(s.X) + i);
      const __nv_bfloat162 a2 = *reinterpret_cast<const __nv_bfloat162*>(&h.x);
      const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&h.y);
      x = make_float4(__low2float(a2), __high2float(a2), __low2float(b2), __high2float(b2));
    } else {
      x = __ldg(reinterpret_cast<const float4*>(s.X) + i);
    }
    float4 a = make_float4(r * x.x, r * x.y, r * x.z, r * x.w);
    if (KS_T > 0) {
      float4 v[KS_T > 0 ? KS_T : 1];
#pragma unroll
      for (int z = 0; z < KS_T; ++z) v[z] = __ldcg(p + (size_t)z * stride4);
#pragma unroll
      for (int z = 0; z < KS_T; ++z) { a.x += v[z].x; a.y += v[z].y; a.z += v[z].z; a.w += v[z].w; }
    } else {
      int z = 0;
      for (; z + 4 <= s.KS; z += 4) {
        const float4 v0 = __ldcg(p + (size_t)(z + 0) * stride4);
        const float4 v1 = __ldcg(p + (size_t)(z + 1) * stride4);
        const float4 v2 = __ldcg(p + (size_t)(z + 2) * stride4);
        const float4 v3 = __ldcg(p + (size_t)(z + 3) * stride4);
        a.x += v0.x; a.y += v0.y; a.z += v0.z; a.w += v0.w;
        a.x += v1.x; a.y += v1.y; a.z += v1.z; a.w += v1.w;
        a.x += v2.x; a.y += v2.y; a.z += v2.z; a.w += v2.w;
        a.x += v3.x; a.y += v3.y; a.z += v3.z; a.w += v3.w;
      }
      for (; z < s.KS; ++z) {
        const float4 v = __ldcg(p + (size_t)z * stride4);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
    }
    reinterpret_cast<float4*>(dxo)[i] = a;
  }
}

// Fixed-order mean of the per-row losses by the last block of a grid to get here (deterministic).
__device__ __forceinline__ void last_block_loss(const Step& s, float* red /* [256] shared */, bool* is_last) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *is_last = (atomicAdd(s.counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!*is_last) return;
  __threadfence();
  float acc = 0.f;
  for (int i = threadIdx.x; i < s.B; i += 256) acc += __ldcg(s.rowloss + i);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (s.loss) *s.loss = red[0] * s.invB;
    *s.counter = 0u;
  }
}

}  // namespace asmh
