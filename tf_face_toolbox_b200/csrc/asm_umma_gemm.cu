// bf16 mode: the four contractions of the A-softmax head as tcgen05 / TMEM tensor-core
// GEMMs fed by TMA, with the head's epilogues fused (hand-written for sm_100a).
//
// One persistent, warp-specialised kernel template (CTA tile 128 x 256 x 64, 4-stage
// TMA->smem ring, 2 accumulator stages of 256 TMEM columns so the epilogue of tile t
// overlaps the MMAs of tile t+1):
//   warp 0  : TMA producer (one elected lane)      warp 1 : tcgen05.mma issuer (one lane)
//   warp 2  : TMEM allocator                       warps 4-7 : epilogue, one TMEM lane each
//
//   KIND        D[MxN]                 A (M x K)                 B (N x K)
//   FWD / BWDG  S  [B x C]    K = D    Xb [B,D]   K-major        Wb [D,Cp]  MN-major
//   DW          dW [D x C]    K = B    Xb [B,D]   MN-major       G''[B,Cp]  MN-major
//   DX          dX [B x D]    K = C    G''[B,Cp]  K-major        Wb [D,Cp]  K-major (split-K)
// so the bf16 copy of W keeps the reference's [D, C] orientation (no transpose anywhere),
// every operand is read by TMA with 128-byte swizzle, and nothing but [B]-sized statistics
// leaves the forward kernel.
#include "asm_common.cuh"
#include "asm_kernels.cuh"
#include "asm_umma.cuh"

namespace asmh {

namespace {
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;           // 16 KB
constexpr int B_BYTES = BN * BK * 2;           // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES; // 48 KB
constexpr int CHUNK_BYTES = 64 * BK * 2;       // one 64-wide MN-major chunk: 8 KB
constexpr int AUX_BYTES = 256 + 4 * BN * 4;    // barriers + tmem ptr, column-sum scratch
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + AUX_BYTES + 1024;
constexpr float LOG2E = 1.4426950408889634f;

enum { U_FWD = 0, U_BWDG = 1, U_DW = 2, U_DX = 3 };

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
}  // namespace

template <int KIND>
__global__ void __launch_bounds__(256, 1)
umma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
            Step s, UmmaArgs g) {
  constexpr bool A_MN = (KIND == U_DW);
  constexpr bool B_MN = (KIND != U_DX);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* colsum = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);  // [4][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&mapA);
    ptx::prefetch_tmap(&mapB);
  }
  if (warp == 1 && ptx::elect_one()) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull[a], 1);
      ptx::mbar_init(&tempty[a], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int tiles_mn = g.mt * g.nt;
  const int total = tiles_mn * g.ks;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      uint32_t it = 0;
      for (int u = blockIdx.x; u < total; u += gridDim.x) {
        const int z = u / tiles_mn, t = u - z * tiles_mn;
        const int n_idx = t / g.mt, m_idx = t - n_idx * g.mt;
        const int m0 = m_idx * BM, n0 = n_idx * BN;
        const int kb0 = z * g.kb_per, kb1 = min(g.kb_total, kb0 + g.kb_per);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int st = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(&empty[st], ph ^ 1);
          ptx::mbar_expect_tx(&full[st], STAGE_BYTES);
          uint8_t* sA = smem + st * STAGE_BYTES;
          uint8_t* sB = sA + A_BYTES;
          const int k0 = kb * BK;
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)
              ptx::tma_load_2d(sA + c * CHUNK_BYTES, &mapA, &full[st], m0 + c * 64, k0);
          } else {
            ptx::tma_load_2d(sA, &mapA, &full[st], k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              ptx::tma_load_2d(sB + c * CHUNK_BYTES, &mapB, &full[st], n0 + c * 64, k0);
          } else {
            ptx::tma_load_2d(sB, &mapB, &full[st], k0, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      const uint64_t hiA = A_MN ? g.desc_hi_mn : g.desc_hi_k;
      const uint64_t hiB = B_MN ? g.desc_hi_mn : g.desc_hi_k;
      const uint32_t stepA = A_MN ? g.kstep_mn : 32u;
      const uint32_t stepB = B_MN ? g.kstep_mn : 32u;
      uint32_t it = 0, lt = 0;
      for (int u = blockIdx.x; u < total; u += gridDim.x, ++lt) {
        const int z = u / tiles_mn;
        const int kb0 = z * g.kb_per, kb1 = min(g.kb_total, kb0 + g.kb_per);
        const uint32_t a = lt & 1, aph = (lt >> 1) & 1;
        ptx::mbar_wait(&tempty[a], aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int st = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(&full[st], ph);
          ptx::tc_fence_after();
          const uint32_t aA = ptx::smem_u32(smem + st * STAGE_BYTES);
          const uint32_t aB = aA + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            ptx::umma_bf16(d_tmem, ptx::smem_desc(hiA, aA + kk * stepA),
                           ptx::smem_desc(hiB, aB + kk * stepB), idesc,
                           (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty[st]);     // frees the smem slot when these MMAs retire
        }
        ptx::umma_commit(&tfull[a]);        // accumulator ready for the epilogue
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (128 threads)
    const int q4 = warp & 3;
    const int row_in_tile = q4 * 32 + lane;
    uint32_t lt = 0;
    for (int u = blockIdx.x; u < total; u += gridDim.x, ++lt) {
      const int z = u / tiles_mn, t = u - z * tiles_mn;
      const int n_idx = t / g.mt, m_idx = t - n_idx * g.mt;
      const int m0 = m_idx * BM, n0 = n_idx * BN;
      const uint32_t a = lt & 1, aph = (lt >> 1) & 1;
      ptx::mbar_wait(&tfull[a], aph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + a * BN;
      const int row = m0 + row_in_tile;
      uint32_t r[32];

      if (KIND == U_FWD) {
        const bool rv = row < s.B;
        const int yl = rv ? s.ylocal[row] : -1;
        float run_m = -INFINITY, run_z = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int jb = n0 + c * 32;
          ptx::tmem_ld32(taddr + c * 32, r);
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int b4 = 0; b4 < 8; ++b4) {
            const float4 ic = __ldg(reinterpret_cast<const float4*>(s.inv_c + jb) + b4);
            v[b4 * 4 + 0] = __uint_as_float(r[b4 * 4 + 0]) * ic.x;
            v[b4 * 4 + 1] = __uint_as_float(r[b4 * 4 + 1]) * ic.y;
            v[b4 * 4 + 2] = __uint_as_float(r[b4 * 4 + 2]) * ic.z;
            v[b4 * 4 + 3] = __uint_as_float(r[b4 * 4 + 3]) * ic.w;
          }
          if (jb + 32 > s.C) {
#pragma unroll
            for (int b = 0; b < 32; ++b)
              if (jb + b >= s.C) v[b] = -INFINITY;
          }
          if (yl >= jb && yl < jb + 32) {
#pragma unroll
            for (int b = 0; b < 32; ++b)
              if (jb + b == yl) {
                s.tgt_s[row] = v[b];
                v[b] = target_logit(v[b], s.n[row], s.inv_n[row], s.m, s.lambda);
                s.tgt_f[row] = v[b];
              }
          }
          if (s.logits && rv) {
#pragma unroll
            for (int b = 0; b < 32; ++b)
              if (jb + b < s.C) s.logits[(size_t)row * s.C + jb + b] = v[b];
          }
          float cm = v[0];
#pragma unroll
          for (int b = 1; b < 32; ++b) cm = fmaxf(cm, v[b]);
          if (cm > -INFINITY) {
            const float nm = fmaxf(run_m, cm);
            const float nml = nm * LOG2E;
            float zs = 0.f;
#pragma unroll
            for (int b = 0; b < 32; ++b) zs += exp2f(fmaf(v[b], LOG2E, -nml));
            run_z = run_z * exp2f((run_m - nm) * LOG2E) + zs;
            run_m = nm;
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tempty[a]);
        if (rv) s.part[(size_t)row * s.NT + n_idx] = make_float2(run_m, run_z);
      } else if (KIND == U_BWDG) {
        const bool rv = row < s.B;
        const int yl = rv ? s.ylocal[row] : -1;
        const float lsel = rv ? s.lse[row] * LOG2E : INFINITY;
        const float gt = rv ? s.gtarget[row] : 0.f;
        __nv_bfloat16* Gw = reinterpret_cast<__nv_bfloat16*>(s.G);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int jb = n0 + c * 32;
          ptx::tmem_ld32(taddr + c * 32, r);
          ptx::tmem_ld_wait();
          float gq[32];   // G'' = G' / c_j
          float pr[32];   // G' * s   (column-sum terms of q_j)
#pragma unroll
          for (int b4 = 0; b4 < 8; ++b4) {
            const float4 ic4 = __ldg(reinterpret_cast<const float4*>(s.inv_c + jb) + b4);
            const float icv[4] = {ic4.x, ic4.y, ic4.z, ic4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int b = b4 * 4 + e;
              const float sv = __uint_as_float(r[b]) * icv[e];
              const float gp = exp2f(fmaf(sv, LOG2E, -lsel)) * s.invB;
              pr[b] = gp * sv;
              gq[b] = gp * icv[e];
            }
          }
          if (yl >= jb && yl < jb + 32) {
#pragma unroll
            for (int b = 0; b < 32; ++b)
              if (jb + b == yl) {
                const float ic = s.inv_c[yl];
                const float sv = __uint_as_float(r[b]) * ic;
                pr[b] = gt * sv;
                gq[b] = gt * ic;
              }
          }
          if (rv) {
            uint4* dst = reinterpret_cast<uint4*>(Gw + (size_t)row * s.Cp + jb);
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
              uint4 o;
              o.x = pack_bf16(gq[v4 * 8 + 0], gq[v4 * 8 + 1]);
              o.y = pack_bf16(gq[v4 * 8 + 2], gq[v4 * 8 + 3]);
              o.z = pack_bf16(gq[v4 * 8 + 4], gq[v4 * 8 + 5]);
              o.w = pack_bf16(gq[v4 * 8 + 6], gq[v4 * 8 + 7]);
              dst[v4] = o;
            }
          }
          // butterfly transpose-reduce: lane l ends with sum over the warp's 32 rows of
          // column (jb + l)
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              const float send = up ? pr[i] : pr[i + off];
              const float keep = up ? pr[i + off] : pr[i];
              pr[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          colsum[q4 * BN + c * 32 + lane] = pr[0];
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tempty[a]);
        named_bar_sync(1, 128);
        {
          const int e = threadIdx.x - 128;
#pragma unroll
          for (int cc = e; cc < BN; cc += 128) {
            const float q = colsum[cc] + colsum[BN + cc] + colsum[2 * BN + cc] + colsum[3 * BN + cc];
            s.q_part[(size_t)m_idx * s.Cp + n0 + cc] = q;
          }
        }
        named_bar_sync(1, 128);
      } else if (KIND == U_DW) {
        const bool rv = row < s.D;                      // row = d
        const int vecw = (s.C % 4 == 0) ? 4 : ((s.C % 2 == 0) ? 2 : 1);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int jb = n0 + c * 32;
          ptx::tmem_ld32(taddr + c * 32, r);
          ptx::tmem_ld_wait();
          if (!rv || jb >= s.C) continue;
          float o[32];
          const uint4* wsrc = reinterpret_cast<const uint4*>(s.Wb + (size_t)row * s.Cp + jb);
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            const uint4 w = __ldg(wsrc + v4);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            const float4 c0 = __ldg(reinterpret_cast<const float4*>(s.coef + jb) + v4 * 2);
            const float4 c1 = __ldg(reinterpret_cast<const float4*>(s.coef + jb) + v4 * 2 + 1);
            const float cf[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __nv_bfloat162 wp = *reinterpret_cast<const __nv_bfloat162*>(&ww[e]);
              const int b = v4 * 8 + e * 2;
              o[b] = fmaf(-__low2float(wp), cf[e * 2], __uint_as_float(r[b]));
              o[b + 1] = fmaf(-__high2float(wp), cf[e * 2 + 1], __uint_as_float(r[b + 1]));
            }
          }
          float* dst = s.dW + (size_t)row * s.C + jb;
          if (jb + 32 <= s.C && vecw == 4) {
#pragma unroll
            for (int b = 0; b < 32; b += 4)
              *reinterpret_cast<float4*>(dst + b) = make_float4(o[b], o[b + 1], o[b + 2], o[b + 3]);
          } else if (jb + 32 <= s.C && vecw == 2) {
#pragma unroll
            for (int b = 0; b < 32; b += 2)
              *reinterpret_cast<float2*>(dst + b) = make_float2(o[b], o[b + 1]);
          } else {
#pragma unroll
            for (int b = 0; b < 32; ++b)
              if (jb + b < s.C) dst[b] = o[b];
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tempty[a]);
      } else {  // U_DX
        const bool rv = row < s.B;
        float* out = s.dx_part + ((size_t)z * s.B + row) * s.D + n0;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          ptx::tmem_ld32(taddr + c * 32, r);
          ptx::tmem_ld_wait();
          if (rv && n0 + c * 32 < s.D) {
#pragma unroll
            for (int b = 0; b < 32; b += 4)
              *reinterpret_cast<float4*>(out + c * 32 + b) =
                  make_float4(__uint_as_float(r[b]), __uint_as_float(r[b + 1]),
                              __uint_as_float(r[b + 2]), __uint_as_float(r[b + 3]));
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tempty[a]);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
          cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// bf16 row-major [outer, inner] tensor with `pitch` elements per row; box = {box_inner, box_outer}
bool encode_map(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch,
                uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

int umma_forward_tiles(int Cp) { return (Cp + BN - 1) / BN; }

int umma_dx_splits(int B, int D, int Cp, int num_sms) {
  const int tiles = ((B + BM - 1) / BM) * ((D + BN - 1) / BN);
  int ks = num_sms / tiles;
  const int kb_total = (Cp + BK - 1) / BK;
  if (ks > kb_total) ks = kb_total;
  if (ks < 1) ks = 1;
  const int kb_per = (kb_total + ks - 1) / ks;
  return (kb_total + kb_per - 1) / kb_per;       // every split non-empty
}

bool umma_build_maps(UmmaMaps* m, const Step& s) {
  bool ok = true;
  // Xb [B, D]
  ok &= encode_map(&m->xb_k, s.Xb, s.D, s.B, s.D, 64, 128);    // A of FWD/BWDG (K-major)
  ok &= encode_map(&m->xb_mn, s.Xb, s.D, s.B, s.D, 64, 64);    // A of DW (MN-major chunks)
  // Wb [D, Cp]
  ok &= encode_map(&m->wb_mn, s.Wb, s.Cp, s.D, s.Cp, 64, 64);  // B of FWD/BWDG (MN-major)
  ok &= encode_map(&m->wb_k, s.Wb, s.Cp, s.D, s.Cp, 64, 256);  // B of DX (K-major)
  // G'' [B, Cp]
  ok &= encode_map(&m->g_k, s.G, s.Cp, s.B, s.Cp, 64, 128);    // A of DX (K-major)
  ok &= encode_map(&m->g_mn, s.G, s.Cp, s.B, s.Cp, 64, 64);    // B of DW (MN-major)
  return ok;
}

static UmmaArgs base_args(const UmmaTuning& tu) {
  UmmaArgs g{};
  g.desc_hi_k = ptx::make_smem_desc_hi(16, 1024);
  g.desc_hi_mn = ptx::make_smem_desc_hi(tu.mn_lbo, tu.mn_sbo);
  g.kstep_mn = tu.mn_kstep;
  g.ks = 1;
  return g;
}

cudaError_t umma_configure() {
  cudaError_t e;
  e = cudaFuncSetAttribute(umma_kernel<U_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(umma_kernel<U_BWDG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(umma_kernel<U_DW>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(umma_kernel<U_DX>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
}

void launch_umma_forward(const Step& s, const UmmaMaps& m, const UmmaTuning& tu, int num_sms,
                         cudaStream_t st) {
  UmmaArgs g = base_args(tu);
  g.mt = (s.B + BM - 1) / BM;
  g.nt = s.Cp / BN;
  g.kb_total = (s.D + BK - 1) / BK;
  g.kb_per = g.kb_total;
  const int total = g.mt * g.nt;
  umma_kernel<U_FWD><<<min(total, num_sms), 256, SMEM_BYTES, st>>>(m.xb_k, m.wb_mn, s, g);
}

void launch_umma_bwdg(const Step& s, const UmmaMaps& m, const UmmaTuning& tu, int num_sms,
                      cudaStream_t st) {
  UmmaArgs g = base_args(tu);   // recompute S -> G'' (bf16) + q_part
  g.mt = (s.B + BM - 1) / BM;
  g.nt = s.Cp / BN;
  g.kb_total = (s.D + BK - 1) / BK;
  g.kb_per = g.kb_total;
  umma_kernel<U_BWDG><<<min(g.mt * g.nt, num_sms), 256, SMEM_BYTES, st>>>(m.xb_k, m.wb_mn, s, g);
}

void launch_umma_dw(const Step& s, const UmmaMaps& m, const UmmaTuning& tu, int num_sms,
                    cudaStream_t st) {
  UmmaArgs g = base_args(tu);   // dW = Xb^T G'' - Wb * coef
  g.mt = (s.D + BM - 1) / BM;
  g.nt = s.Cp / BN;
  g.kb_total = (s.B + BK - 1) / BK;
  g.kb_per = g.kb_total;
  umma_kernel<U_DW><<<min(g.mt * g.nt, num_sms), 256, SMEM_BYTES, st>>>(m.xb_mn, m.g_mn, s, g);
}

void launch_umma_dx(const Step& s, const UmmaMaps& m, const UmmaTuning& tu, int num_sms,
                    cudaStream_t st) {
  UmmaArgs g = base_args(tu);   // dX partials = G'' Wb^T, split over the classes
  g.mt = (s.B + BM - 1) / BM;
  g.nt = (s.D + BN - 1) / BN;
  g.kb_total = (s.Cp + BK - 1) / BK;
  g.ks = s.KS;
  g.kb_per = (g.kb_total + g.ks - 1) / g.ks;
  umma_kernel<U_DX><<<min(g.mt * g.nt * g.ks, num_sms), 256, SMEM_BYTES, st>>>(m.g_k, m.wb_k, s, g);
}

}  // namespace asmh
