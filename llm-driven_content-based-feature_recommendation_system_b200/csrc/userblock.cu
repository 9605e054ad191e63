// Same-user correction of the distinct-item in-batch softmax (losses.logq_infonce_columns, C2 of SURVEY.md 8a).
//
// The fused tensor-core pass scores every row against the DISTINCT items of the batch with their multiplicities.
// The reference also masks the columns that belong to the row's own user (tower_code/v1_refine_usertower.py:848):
// those are at most L per row and they all live in the row's own user block, so they are handled here as one small
// dense block per user: rows i of user b  x  the target items of the same rows,
//
//     s_ij    = scale * <u_i, c[pos_col_j]> - bias[pos_col_j]          i, j in the user's rows
//     s_pos_i = s_ii                                                   (the label logit)
//     own_lse_i = log sum_{j : pos_col_j != pos_col_i} exp(s_ij)       (-inf when the user has no other item)
//
// which the host combines with the dense pass: Z_i = e^{lse0_i} + e^{s_pos_i} - e^{own_lse_i}.
// One CTA per user; the user's u rows and item rows are staged in shared memory once (fp32, padded stride), the
// l x l logits live in shared memory, and the backward adds ONE gradient row per (user, item) into d_cols instead
// of one per (row, item) pair.
#include "common.cuh"
#include "../../include/rs_twotower.h"

namespace rs {

#define UB_D 128
#define UB_STRIDE 132          // padded row stride (elements), see UbStore
#define UB_THREADS 128

struct UbParams {
  const void* u; const void* cols;
  const int64_t* pos_col;     // [n_rows]
  const int* row_cu;          // [n_users + 1] row offsets of the users
  const float* col_bias;      // [n_cols] or NULL
  int64_t n_users, n_cols;
  int max_len;
  float scale;
};

// The staged rows keep the operands' own element type (16-bit operands: half the shared memory of an fp32 copy, so
// twice as many user blocks are resident per SM -- the kernels are bound by the latency of the row gathers, not by
// arithmetic).  Row stride 132 elements: 528 B (fp32, 128-bit reads) / 264 B (16-bit, 64-bit reads) -- reads of the
// same column of 32 different rows are bank-conflict free in both cases.
template <int DT> struct UbStore { using T = uint16_t; };
template <> struct UbStore<RS_F32> { using T = float; };

template <int DT>
__device__ __forceinline__ float4 ub_lds4(const typename UbStore<DT>::T* p) {         // 4 consecutive elements as fp32
  if constexpr (DT == RS_F32) return *reinterpret_cast<const float4*>(p);
  else {
    const uint2 w = *reinterpret_cast<const uint2*>(p);
    float2 a, b;
    if constexpr (DT == RS_BF16) { a = unpack_bf16(w.x); b = unpack_bf16(w.y); }
    else { a = unpack_f16(w.x); b = unpack_f16(w.y); }
    return make_float4(a.x, a.y, b.x, b.y);
  }
}
template <int DT>
__device__ __forceinline__ float ub_lds1(const typename UbStore<DT>::T* p) {
  if constexpr (DT == RS_F32) return *p;
  else if constexpr (DT == RS_BF16) return __uint_as_float((uint32_t)(*p) << 16);
  else return __half2float(*reinterpret_cast<const __half*>(p));
}
// copy 4 consecutive elements of a global row (element offset `off`) into shared memory, bit for bit
template <int DT>
__device__ __forceinline__ void ub_copy4(typename UbStore<DT>::T* dst, const void* base, int64_t off, bool ok) {
  if constexpr (DT == RS_F32) {
    *reinterpret_cast<float4*>(dst) = ok ? __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    *reinterpret_cast<uint2*>(dst) = ok ? __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + off))
                                        : make_uint2(0u, 0u);
  }
}

// 4 consecutive elements of a global row as stored (bit for bit), zeros when !ok
template <int DT> struct UbVec { using T = uint2; };
template <> struct UbVec<RS_F32> { using T = float4; };
template <int DT>
__device__ __forceinline__ typename UbVec<DT>::T ub_ldg4(const void* base, int64_t off, bool ok) {
  if constexpr (DT == RS_F32)
    return ok ? __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
  else
    return ok ? __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + off)) : make_uint2(0u, 0u);
}

// stage the user's rows: sU[i][:] = u[r0 + i], sC[i][:] = cols[pos_col[r0 + i]]; sCol[i] = pos_col (int), sBias[i]
template <int DT>
__device__ __forceinline__ void ub_stage(const UbParams& p, int64_t r0, int len, typename UbStore<DT>::T* sU,
                                         typename UbStore<DT>::T* sC, int* sCol, float* sBias) {
  for (int i = threadIdx.x; i < len; i += UB_THREADS) {
    const int64_t c = __ldg(p.pos_col + r0 + i);
    const bool ok = c >= 0 && c < p.n_cols;
    sCol[i] = ok ? (int)c : -1;
    sBias[i] = (ok && p.col_bias) ? __ldg(p.col_bias + c) : 0.f;
  }
  __syncthreads();
  // a warp fetches 4 row pairs per round: all 8 global loads are issued before the first shared-memory store (the
  // one-pair-per-iteration loop exposed a full memory latency per row: the kernels are bound by these gathers)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int NW = UB_THREADS / 32, G = 4;
  using VT = typename UbVec<DT>::T;
  for (int i0 = w; i0 < len; i0 += NW * G) {
    VT vu[G], vc[G];
#pragma unroll
    for (int q = 0; q < G; ++q) {
      const int i = i0 + q * NW;
      const bool live = i < len;
      const int c = live ? sCol[i] : -1;
      vu[q] = ub_ldg4<DT>(p.u, (r0 + (live ? i : 0)) * UB_D + 4 * lane, live);
      vc[q] = ub_ldg4<DT>(p.cols, (int64_t)(c >= 0 ? c : 0) * UB_D + 4 * lane, c >= 0);
    }
#pragma unroll
    for (int q = 0; q < G; ++q) {
      const int i = i0 + q * NW;
      if (i < len) {
        *reinterpret_cast<VT*>(sU + i * UB_STRIDE + 4 * lane) = vu[q];
        *reinterpret_cast<VT*>(sC + i * UB_STRIDE + 4 * lane) = vc[q];
      }
    }
  }
  __syncthreads();
}

// sS[i][j] = scale * <u_i, c_j> - bias_j for all (i, j) of the block
template <int DT>
__device__ __forceinline__ void ub_logits(const UbParams& p, int len, const typename UbStore<DT>::T* sU,
                                          const typename UbStore<DT>::T* sC, const float* sBias, float* sS) {
  for (int pr = threadIdx.x; pr < len * len; pr += UB_THREADS) {
    const int i = pr / len, j = pr - i * len;
    const typename UbStore<DT>::T* a = sU + i * UB_STRIDE;
    const typename UbStore<DT>::T* b = sC + j * UB_STRIDE;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
    for (int k = 0; k < UB_D; k += 4) {
      const float4 x = ub_lds4<DT>(a + k), y = ub_lds4<DT>(b + k);
      a0 = fmaf(x.x, y.x, a0); a1 = fmaf(x.y, y.y, a1); a2 = fmaf(x.z, y.z, a2); a3 = fmaf(x.w, y.w, a3);
    }
    sS[i * p.max_len + j] = ((a0 + a1) + (a2 + a3)) * p.scale - sBias[j];
  }
  __syncthreads();
}

template <int DT>
__global__ void __launch_bounds__(UB_THREADS) ub_fwd_kernel(UbParams p, float* __restrict__ s_pos,
                                                            float* __restrict__ own_lse) {
  using ST = typename UbStore<DT>::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ST* sU = reinterpret_cast<ST*>(smem_raw);
  ST* sC = sU + p.max_len * UB_STRIDE;
  float* sS = reinterpret_cast<float*>(sC + p.max_len * UB_STRIDE);
  float* sBias = sS + p.max_len * p.max_len;
  int* sCol = reinterpret_cast<int*>(sBias + p.max_len);
  for (int64_t b = blockIdx.x; b < p.n_users; b += gridDim.x) {
    const int64_t r0 = __ldg(p.row_cu + b);
    const int true_len = (int)(__ldg(p.row_cu + b + 1) - r0);
    const int len = min(true_len, p.max_len);
    // a user with more rows than the block holds cannot be represented: poison the rows that do not fit so that the
    // loss turns NaN (loud) instead of carrying silently wrong values
    for (int i = len + threadIdx.x; i < true_len; i += UB_THREADS) {
      s_pos[r0 + i] = __int_as_float(0x7fc00000);
      own_lse[r0 + i] = __int_as_float(0x7fc00000);
    }
    __syncthreads();
    ub_stage<DT>(p, r0, len, sU, sC, sCol, sBias);
    ub_logits<DT>(p, len, sU, sC, sBias, sS);
    for (int i = threadIdx.x; i < len; i += UB_THREADS) {
      const int ci = sCol[i];
      float m = -INFINITY;
      for (int j = 0; j < len; ++j)
        if (sCol[j] != ci && sCol[j] >= 0) m = fmaxf(m, sS[i * p.max_len + j]);
      float l = 0.f;
      if (m > -INFINITY)
        for (int j = 0; j < len; ++j)
          if (sCol[j] != ci && sCol[j] >= 0) l += __expf(sS[i * p.max_len + j] - m);
      s_pos[r0 + i] = ci >= 0 ? sS[i * p.max_len + i] : -INFINITY;
      own_lse[r0 + i] = m > -INFINITY ? m + __logf(l) : -INFINITY;
    }
  }
}

// d_u[i] = scale * sum_j dS_ij c_j ;  d_cols[pos_col_j] += scale * sum_i dS_ij u_i
//   dS_ij = g_own_i * exp(s_ij - own_lse_i) for pos_col_j != pos_col_i,  dS_ii = g_pos_i
template <int DT>
__global__ void __launch_bounds__(UB_THREADS) ub_bwd_kernel(UbParams p, const float* __restrict__ own_lse,
                                                            const float* __restrict__ g_pos,
                                                            const float* __restrict__ g_own, float* __restrict__ d_u,
                                                            float* __restrict__ d_cols) {
  using ST = typename UbStore<DT>::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ST* sU = reinterpret_cast<ST*>(smem_raw);
  ST* sC = sU + p.max_len * UB_STRIDE;
  float* sS = reinterpret_cast<float*>(sC + p.max_len * UB_STRIDE);
  float* sBias = sS + p.max_len * p.max_len;
  int* sCol = reinterpret_cast<int*>(sBias + p.max_len);
  for (int64_t b = blockIdx.x; b < p.n_users; b += gridDim.x) {
    const int64_t r0 = __ldg(p.row_cu + b);
    const int len = min((int)(__ldg(p.row_cu + b + 1) - r0), p.max_len);
    __syncthreads();
    ub_stage<DT>(p, r0, len, sU, sC, sCol, sBias);
    ub_logits<DT>(p, len, sU, sC, sBias, sS);
    // logits -> coefficients, in place
    for (int pr = threadIdx.x; pr < len * len; pr += UB_THREADS) {
      const int i = pr / len, j = pr - i * len;
      const int ci = sCol[i], cj = sCol[j];
      float c = 0.f;
      if (cj >= 0 && ci >= 0) {
        if (i == j) c = __ldg(g_pos + r0 + i);
        else if (cj != ci) {
          const float ol = __ldg(own_lse + r0 + i);
          c = ol > -INFINITY ? __ldg(g_own + r0 + i) * __expf(sS[pr / len * p.max_len + j] - ol) : 0.f;
        }
      }
      sS[i * p.max_len + j] = c * p.scale;
    }
    __syncthreads();
    // thread k owns feature k of every row; 8 rows (resp. 8 columns) of the coefficient block per pass over the other
    // index, so that one staged element feeds 8 FMAs (the kernel is instruction-issue bound: ncu r02b, 63 % issue)
    const int k = threadIdx.x;
    for (int i0 = 0; i0 < len; i0 += 8) {
      float acc[8];
      int ri[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { acc[e] = 0.f; ri[e] = min(i0 + e, len - 1) * p.max_len; }
      for (int j = 0; j < len; ++j) {
        const float c = ub_lds1<DT>(sC + j * UB_STRIDE + k);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(sS[ri[e] + j], c, acc[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (i0 + e < len) d_u[(r0 + i0 + e) * UB_D + k] = acc[e];
    }
    for (int j0 = 0; j0 < len; j0 += 8) {          // (max_len is a multiple of 4: the second float4 may lie past `len`,
      float acc[8];                                //  inside the row's padding or the next row -- finite, discarded)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      const bool two = j0 + 4 < p.max_len;
      for (int i = 0; i < len; ++i) {
        const float u = ub_lds1<DT>(sU + i * UB_STRIDE + k);
        const float4 s0 = *reinterpret_cast<const float4*>(sS + i * p.max_len + j0);
        const float4 s1 = two ? *reinterpret_cast<const float4*>(sS + i * p.max_len + j0 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[0] = fmaf(s0.x, u, acc[0]); acc[1] = fmaf(s0.y, u, acc[1]); acc[2] = fmaf(s0.z, u, acc[2]); acc[3] = fmaf(s0.w, u, acc[3]);
        acc[4] = fmaf(s1.x, u, acc[4]); acc[5] = fmaf(s1.y, u, acc[5]); acc[6] = fmaf(s1.z, u, acc[6]); acc[7] = fmaf(s1.w, u, acc[7]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int j = j0 + e;
        if (j < len && sCol[j] >= 0) atomicAdd(d_cols + (int64_t)sCol[j] * UB_D + k, acc[e]);
      }
    }
  }
}

// =============================================================================================
// 16-bit operands: the three small products of a user block on the tensor cores (mma.sync m16n8k16)
// =============================================================================================
// A user block is l <= 64 rows: S = U C^T is [l, l] over K = 128, and the two gradient products are [l, l] x [l, 128].
// The SIMT kernels above spend ~1000 issue slots per thread on them (ub_bwd_kernel: 63 % issue in ncu r02b); here a warp
// owns a 16-row tile and the whole block costs a few dozen mma instructions.  Operands stay in shared memory as staged
// (row pitch 136 elements = 272 B: 16-byte aligned rows, conflict-free ldmatrix).  The gradient coefficients are split
// into a 16-bit head and a 16-bit remainder (two mma per product): their rounding error is 2^-17, not 2^-9, so the
// gradients keep the accuracy of the fp32-coefficient kernels.  Logits, softmax statistics and coefficients never
// leave registers (quad shuffles reduce a row).
#define UBM_LD 136
#define UBM_CLD 72             // coefficient planes: [64][72] 16-bit

template <int DT>
__device__ __forceinline__ void ubm_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (DT == RS_BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ubm_ldsm4(uint32_t (&r)[4], const uint16_t* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ubm_ldsm4_trans(uint32_t (&r)[4], const uint16_t* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <int DT> __device__ __forceinline__ uint16_t ubm_to16(float x) {
  if constexpr (DT == RS_BF16) { __nv_bfloat16 h = __float2bfloat16_rn(x); return *reinterpret_cast<uint16_t*>(&h); }
  else { __half h = __float2half_rn(x); return *reinterpret_cast<uint16_t*>(&h); }
}
template <int DT> __device__ __forceinline__ float ubm_from16(uint16_t u) {
  if constexpr (DT == RS_BF16) return __uint_as_float((uint32_t)u << 16);
  else return __half2float(*reinterpret_cast<const __half*>(&u));
}
__device__ __forceinline__ float ubm_quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float ubm_quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ void red_add_f2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// stage rows [0, LT) (LT = len rounded up to 16): real rows from global memory, the padding rows as zeros
__device__ __forceinline__ void ubm_stage(const UbParams& p, int64_t r0, int len, int LT, uint16_t* sU, uint16_t* sC,
                                          int* sCol, float* sBias) {
  for (int i = threadIdx.x; i < LT; i += UB_THREADS) {
    int64_t c = -1;
    if (i < len) c = __ldg(p.pos_col + r0 + i);
    const bool ok = c >= 0 && c < p.n_cols;
    sCol[i] = ok ? (int)c : -1;
    sBias[i] = (ok && p.col_bias) ? __ldg(p.col_bias + c) : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int NW = UB_THREADS / 32, G = 4;
  for (int i0 = w; i0 < LT; i0 += NW * G) {
    uint2 vu[G], vc[G];
#pragma unroll
    for (int q = 0; q < G; ++q) {
      const int i = i0 + q * NW;
      const bool live = i < len;
      const int c = i < LT ? sCol[i] : -1;
      vu[q] = live ? __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p.u) + (r0 + i) * UB_D + 4 * lane))
                   : make_uint2(0u, 0u);
      vc[q] = c >= 0 ? __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p.cols) + (int64_t)c * UB_D + 4 * lane))
                     : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int q = 0; q < G; ++q) {
      const int i = i0 + q * NW;
      if (i < LT) {
        *reinterpret_cast<uint2*>(sU + i * UBM_LD + 4 * lane) = vu[q];
        *reinterpret_cast<uint2*>(sC + i * UBM_LD + 4 * lane) = vc[q];
      }
    }
  }
  __syncthreads();
}

// s[ct][e]: logits of the warp's 16-row tile `rt` against the column tiles ct < LT/8 (8 columns each), scale and bias
// applied.  Fragment layout: e = 0,1 -> row g, columns 8 ct + 2 t (+1); e = 2,3 -> row g + 8.
template <int DT>
__device__ __forceinline__ void ubm_logits(float (&s)[8][4], const UbParams& p, int rt, int LT, const uint16_t* sU,
                                           const uint16_t* sC, const float* sBias, int lane) {
  uint32_t a[8][4];
  {
    const uint16_t* base = sU + (rt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * UBM_LD + (lane >> 4) * 8;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) ubm_ldsm4(a[ks], base + 16 * ks);
  }
  const int t = lane & 3;
#pragma unroll
  for (int ct = 0; ct < 8; ++ct) {
    s[ct][0] = s[ct][1] = s[ct][2] = s[ct][3] = 0.f;
    if (ct * 8 < LT) {
      const uint16_t* bb = sC + (ct * 8 + (lane & 7)) * UBM_LD + (lane >> 3) * 8;
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {                  // two k-steps per ldmatrix.x4
        uint32_t r[4];
        ubm_ldsm4(r, bb + 32 * k2);
        ubm_mma<DT>(s[ct], a[2 * k2], r[0], r[1]);
        ubm_mma<DT>(s[ct], a[2 * k2 + 1], r[2], r[3]);
      }
      const float b0 = sBias[ct * 8 + 2 * t], b1 = sBias[ct * 8 + 2 * t + 1];
      s[ct][0] = s[ct][0] * p.scale - b0; s[ct][1] = s[ct][1] * p.scale - b1;
      s[ct][2] = s[ct][2] * p.scale - b0; s[ct][3] = s[ct][3] * p.scale - b1;
    }
  }
}

// label logit and same-user log-sum-exp of the 16 rows of tile `rt` (one warp)
template <int DT>
__device__ __forceinline__ void ubm_fwd_rows(const UbParams& p, int64_t r0, int len, int LT, int rt, const uint16_t* sU,
                                             const uint16_t* sC, const float* sBias, const int* sCol, int lane,
                                             float* __restrict__ s_pos, float* __restrict__ own_lse) {
  const int g = lane >> 2, t = lane & 3;
    float s[8][4];
    ubm_logits<DT>(s, p, rt, LT, sU, sC, sBias, lane);
    const int i0 = rt * 16 + g, i1 = i0 + 8;
    const int ci[2] = {sCol[i0], sCol[i1]};
    float m[2] = {-INFINITY, -INFINITY}, sp[2] = {0.f, 0.f};
#pragma unroll
    for (int ct = 0; ct < 8; ++ct)
      if (ct * 8 < LT) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = ct * 8 + 2 * t + (e & 1), r = e >> 1;
          const int cj = sCol[j];
          if (j == (r ? i1 : i0)) sp[r] = s[ct][e];
          if (cj >= 0 && cj != ci[r]) m[r] = fmaxf(m[r], s[ct][e]);
        }
      }
    m[0] = ubm_quad_max(m[0]); m[1] = ubm_quad_max(m[1]);
    float l[2] = {0.f, 0.f};
#pragma unroll
    for (int ct = 0; ct < 8; ++ct)
      if (ct * 8 < LT) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = ct * 8 + 2 * t + (e & 1), r = e >> 1;
          const int cj = sCol[j];
          if (cj >= 0 && cj != ci[r] && m[r] > -INFINITY) l[r] += __expf(s[ct][e] - m[r]);
        }
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float lt = ubm_quad_sum(l[r]), spv = ubm_quad_sum(sp[r]);
      const int i = r ? i1 : i0;
      if (t == 0 && i < len) {
        s_pos[r0 + i] = ci[r] >= 0 ? spv : -INFINITY;
        own_lse[r0 + i] = m[r] > -INFINITY ? m[r] + __logf(lt) : -INFINITY;
      }
    }
}

template <int DT>
__global__ void __launch_bounds__(UB_THREADS) ub_fwd_mma_kernel(UbParams p, int ML16, float* __restrict__ s_pos,
                                                                float* __restrict__ own_lse) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint16_t* sU = reinterpret_cast<uint16_t*>(smem_raw);
  uint16_t* sC = sU + ML16 * UBM_LD;
  float* sBias = reinterpret_cast<float*>(sC + ML16 * UBM_LD);
  int* sCol = reinterpret_cast<int*>(sBias + ML16);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t b = blockIdx.x; b < p.n_users; b += gridDim.x) {
    const int64_t r0 = __ldg(p.row_cu + b);
    const int true_len = (int)(__ldg(p.row_cu + b + 1) - r0);
    const int len = min(true_len, p.max_len);
    for (int i = len + threadIdx.x; i < true_len; i += UB_THREADS) {          // (see ub_fwd_kernel)
      s_pos[r0 + i] = __int_as_float(0x7fc00000);
      own_lse[r0 + i] = __int_as_float(0x7fc00000);
    }
    if (len <= 0) continue;
    const int LT = (len + 15) & ~15;
    __syncthreads();
    ubm_stage(p, r0, len, LT, sU, sC, sCol, sBias);
    if (w * 16 < len) ubm_fwd_rows<DT>(p, r0, len, LT, w, sU, sC, sBias, sCol, lane, s_pos, own_lse);
  }
}

// acc[4 n-tiles of 8 features][4] += (A_hi + A_lo)[16 x 16] . T[16 rows x 32 features], T staged with pitch UBM_LD
template <int DT>
__device__ __forceinline__ void ubm_mma_tile(float (&acc)[4][4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                             const uint16_t* tile, int lane) {
  const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int np = 0; np < 2; ++np) {
    uint32_t r[4];
    ubm_ldsm4_trans(r, tile + ((mi & 1) * 8 + rr) * UBM_LD + 8 * (2 * np + (mi >> 1)));
    ubm_mma<DT>(acc[2 * np], ah, r[0], r[1]);
    ubm_mma<DT>(acc[2 * np], al, r[0], r[1]);
    ubm_mma<DT>(acc[2 * np + 1], ah, r[2], r[3]);
    ubm_mma<DT>(acc[2 * np + 1], al, r[2], r[3]);
  }
}

// gradient coefficients of the 16 rows of tile `rt` (one warp) -> sH / sL (16-bit head + remainder)
template <int DT>
__device__ __forceinline__ void ubm_bwd_coef(const UbParams& p, int64_t r0, int len, int LT, int rt, const uint16_t* sU,
                                             const uint16_t* sC, uint16_t* sH, uint16_t* sL, const float* sBias,
                                             const int* sCol, int lane, const float* __restrict__ own_lse,
                                             const float* __restrict__ g_pos, const float* __restrict__ g_own) {
  const int g = lane >> 2, t = lane & 3;
    float s[8][4];
    ubm_logits<DT>(s, p, rt, LT, sU, sC, sBias, lane);
    const int i0 = rt * 16 + g, i1 = i0 + 8;
    const int ci[2] = {sCol[i0], sCol[i1]};
    float gp[2] = {0.f, 0.f}, go[2] = {0.f, 0.f}, ol[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = r ? i1 : i0;
      if (i < len && ci[r] >= 0) { gp[r] = __ldg(g_pos + r0 + i); go[r] = __ldg(g_own + r0 + i); ol[r] = __ldg(own_lse + r0 + i); }
    }
#pragma unroll
    for (int ct = 0; ct < 8; ++ct)
      if (ct * 8 < LT) {
        float c[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = ct * 8 + 2 * t + (e & 1), r = e >> 1;
          const int i = r ? i1 : i0;
          const int cj = sCol[j];
          float v = 0.f;
          if (ci[r] >= 0 && cj >= 0) {
            if (i == j) v = gp[r];
            else if (cj != ci[r] && ol[r] > -INFINITY) v = go[r] * __expf(s[ct][e] - ol[r]);
          }
          c[e] = v * p.scale;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int i = r ? i1 : i0;
          const uint16_t h0 = ubm_to16<DT>(c[2 * r]), h1 = ubm_to16<DT>(c[2 * r + 1]);
          const uint16_t l0 = ubm_to16<DT>(c[2 * r] - ubm_from16<DT>(h0)), l1 = ubm_to16<DT>(c[2 * r + 1] - ubm_from16<DT>(h1));
          *reinterpret_cast<uint32_t*>(sH + i * UBM_CLD + ct * 8 + 2 * t) = (uint32_t)h0 | ((uint32_t)h1 << 16);
          *reinterpret_cast<uint32_t*>(sL + i * UBM_CLD + ct * 8 + 2 * t) = (uint32_t)l0 | ((uint32_t)l1 << 16);
        }
      }
}

// d_u rows of tile `rt` and the d_cols contributions of the items of tile `rt` (one warp; the coefficient planes of ALL
// tiles must be complete)
template <int DT>
__device__ __forceinline__ void ubm_bwd_products(const UbParams& p, int64_t r0, int len, int LT, int rt, const uint16_t* sU,
                                                 const uint16_t* sC, const uint16_t* sH, const uint16_t* sL,
                                                 const int* sCol, int lane, float* __restrict__ d_u,
                                                 float* __restrict__ d_cols) {
  const int g = lane >> 2, t = lane & 3;
    // ---- d_u rows of tile rt:  sum_j coef[i][j] c_j      A = coef[16 i][16 j] (row-major), B = sC rows j (trans)
    const uint16_t* abase = (rt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * UBM_CLD + (lane >> 4) * 8 + sH;
    const int i0 = rt * 16 + g, i1 = i0 + 8;
#pragma unroll 1
    for (int fg = 0; fg < 4; ++fg) {
      float acc[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
      for (int kt = 0; kt * 16 < LT; ++kt) {
        uint32_t ah[4], al[4];
        ubm_ldsm4(ah, abase + kt * 16);
        ubm_ldsm4(al, abase + kt * 16 + (sL - sH));
        ubm_mma_tile<DT>(acc, ah, al, sC + kt * 16 * UBM_LD + fg * 32, lane);
      }
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int f = fg * 32 + 8 * n + 2 * t;
        if (i0 < len) *reinterpret_cast<float2*>(d_u + (r0 + i0) * UB_D + f) = make_float2(acc[n][0], acc[n][1]);
        if (i1 < len) *reinterpret_cast<float2*>(d_u + (r0 + i1) * UB_D + f) = make_float2(acc[n][2], acc[n][3]);
      }
    }
    // ---- d_cols rows of the items j of tile rt:  sum_i coef[i][j] u_i      A = coef^T (ldmatrix.trans), B = sU rows i
    const int j0 = rt * 16 + g, j1 = j0 + 8;
    const int cj0 = sCol[j0], cj1 = sCol[j1];
    const uint16_t* tbase = sH + ((lane & 7) + (lane >> 4) * 8) * UBM_CLD + rt * 16 + ((lane >> 3) & 1) * 8;
#pragma unroll 1
    for (int fg = 0; fg < 4; ++fg) {
      float acc[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
      for (int it = 0; it * 16 < LT; ++it) {
        uint32_t ah[4], al[4];
        ubm_ldsm4_trans(ah, tbase + it * 16 * UBM_CLD);
        ubm_ldsm4_trans(al, tbase + it * 16 * UBM_CLD + (sL - sH));
        ubm_mma_tile<DT>(acc, ah, al, sU + it * 16 * UBM_LD + fg * 32, lane);
      }
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int f = fg * 32 + 8 * n + 2 * t;
        if (cj0 >= 0) red_add_f2(d_cols + (int64_t)cj0 * UB_D + f, acc[n][0], acc[n][1]);
        if (cj1 >= 0) red_add_f2(d_cols + (int64_t)cj1 * UB_D + f, acc[n][2], acc[n][3]);
      }
    }
}

template <int DT>
__global__ void __launch_bounds__(UB_THREADS) ub_bwd_mma_kernel(UbParams p, int ML16, const float* __restrict__ own_lse,
                                                                const float* __restrict__ g_pos,
                                                                const float* __restrict__ g_own, float* __restrict__ d_u,
                                                                float* __restrict__ d_cols) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint16_t* sU = reinterpret_cast<uint16_t*>(smem_raw);
  uint16_t* sC = sU + ML16 * UBM_LD;
  uint16_t* sH = sC + ML16 * UBM_LD;                    // coefficient heads  [ML16][UBM_CLD]
  uint16_t* sL = sH + ML16 * UBM_CLD;                   // coefficient remainders
  float* sBias = reinterpret_cast<float*>(sL + ML16 * UBM_CLD);
  int* sCol = reinterpret_cast<int*>(sBias + ML16);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t b = blockIdx.x; b < p.n_users; b += gridDim.x) {
    const int64_t r0 = __ldg(p.row_cu + b);
    const int true_len = (int)(__ldg(p.row_cu + b + 1) - r0);
    const int len = min(true_len, p.max_len);
    if (len <= 0) continue;
    const int LT = (len + 15) & ~15;
    __syncthreads();
    ubm_stage(p, r0, len, LT, sU, sC, sCol, sBias);
    if (w * 16 < LT) ubm_bwd_coef<DT>(p, r0, len, LT, w, sU, sC, sH, sL, sBias, sCol, lane, own_lse, g_pos, g_own);
    __syncthreads();
    if (w * 16 < LT) ubm_bwd_products<DT>(p, r0, len, LT, w, sU, sC, sH, sL, sCol, lane, d_u, d_cols);
  }
}

// (A warp-per-user variant for users with <= 16 rows -- four independent users per CTA, no block barriers -- was measured
// and dropped: tools/ub_probe.py, backward 156 -> 165 us.  The kernels are bound by the row gathers and the item-gradient
// reductions, not by the barriers.)
}  // namespace rs

using namespace rs;

static size_t ubm_smem(int ml16, bool bwd) {
  return (size_t)2 * ml16 * UBM_LD * 2 + (bwd ? (size_t)2 * ml16 * UBM_CLD * 2 : 0) + (size_t)ml16 * (sizeof(float) + sizeof(int));
}
static size_t ub_smem(int max_len, int dtype) {
  const size_t esz = dtype == RS_F32 ? 4 : 2;
  return (size_t)2 * max_len * UB_STRIDE * esz + ((size_t)max_len * max_len + 2 * (size_t)max_len) * sizeof(float);
}

#define UB_DISPATCH(dt, NAME, ...)                                      \
  switch (dt) {                                                         \
    case RS_F32: { constexpr int NAME = RS_F32; __VA_ARGS__; break; }   \
    case RS_F16: { constexpr int NAME = RS_F16; __VA_ARGS__; break; }   \
    case RS_BF16: { constexpr int NAME = RS_BF16; __VA_ARGS__; break; } \
    default: return RS_ERR_BAD_ARG;                                     \
  }

static int ub_check(const void* u, const void* cols, const int64_t* pos_col, const int32_t* row_cu, int64_t n_users,
                    int64_t n_cols, int64_t dim, int max_len) {
  if (!u || !cols || !pos_col || !row_cu || n_users <= 0 || n_cols <= 0 || max_len <= 0) return RS_ERR_BAD_ARG;
  if (dim != UB_D || max_len > 64) return RS_ERR_UNSUPPORTED;
  return RS_OK;
}

extern "C" int rs_user_block_logits_fwd(const void* u, const void* cols, int dtype, const int64_t* pos_col,
                                        const int32_t* row_cu, int64_t n_users, int64_t n_cols, int64_t dim, int max_len,
                                        float scale, const float* col_bias, float* s_pos, float* own_lse, void* stream) {
  int rc = ub_check(u, cols, pos_col, row_cu, n_users, n_cols, dim, max_len);
  if (rc != RS_OK) return rc;
  if (!s_pos || !own_lse) return RS_ERR_BAD_ARG;
  max_len = (max_len + 3) & ~3;
  UbParams p = {u, cols, pos_col, row_cu, col_bias, n_users, n_cols, max_len, scale};
  const size_t smem = ub_smem(max_len, dtype);
  const int grid = (int)(n_users < (int64_t)RS_NUM_SMS * 12 ? n_users : (int64_t)RS_NUM_SMS * 12);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype != RS_F32) {                       // 16-bit operands: tensor-core tiles
    const int ml16 = (max_len + 15) & ~15;
    const size_t sm = ubm_smem(ml16, false);
#define UBM_FWD(DT)                                                                                                   \
    do {                                                                                                              \
      cudaError_t e = cudaFuncSetAttribute(ub_fwd_mma_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
      if (e != cudaSuccess) return (int)e;                                                                            \
      ub_fwd_mma_kernel<DT><<<grid, UB_THREADS, sm, st>>>(p, ml16, s_pos, own_lse);                                    \
    } while (0)
    if (dtype == RS_BF16) UBM_FWD(RS_BF16); else if (dtype == RS_F16) UBM_FWD(RS_F16); else return RS_ERR_BAD_ARG;
#undef UBM_FWD
    RS_LAUNCH_CHECK();
    return RS_OK;
  }
  UB_DISPATCH(dtype, DT, {
    cudaError_t e = cudaFuncSetAttribute(ub_fwd_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ub_fwd_kernel<DT><<<grid, UB_THREADS, smem, st>>>(p, s_pos, own_lse);
  });
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_user_block_logits_bwd(const void* u, const void* cols, int dtype, const int64_t* pos_col,
                                        const int32_t* row_cu, int64_t n_users, int64_t n_cols, int64_t dim, int max_len,
                                        float scale, const float* col_bias, const float* own_lse, const float* g_pos,
                                        const float* g_own, float* d_u, float* d_cols, void* stream) {
  int rc = ub_check(u, cols, pos_col, row_cu, n_users, n_cols, dim, max_len);
  if (rc != RS_OK) return rc;
  if (!own_lse || !g_pos || !g_own || !d_u || !d_cols) return RS_ERR_BAD_ARG;
  max_len = (max_len + 3) & ~3;
  UbParams p = {u, cols, pos_col, row_cu, col_bias, n_users, n_cols, max_len, scale};
  const size_t smem = ub_smem(max_len, dtype);
  const int grid = (int)(n_users < (int64_t)RS_NUM_SMS * 12 ? n_users : (int64_t)RS_NUM_SMS * 12);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype != RS_F32) {
    const int ml16 = (max_len + 15) & ~15;
    const size_t sm = ubm_smem(ml16, true);
#define UBM_BWD(DT)                                                                                                   \
    do {                                                                                                              \
      cudaError_t e = cudaFuncSetAttribute(ub_bwd_mma_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
      if (e != cudaSuccess) return (int)e;                                                                            \
      ub_bwd_mma_kernel<DT><<<grid, UB_THREADS, sm, st>>>(p, ml16, own_lse, g_pos, g_own, d_u, d_cols);                \
    } while (0)
    if (dtype == RS_BF16) UBM_BWD(RS_BF16); else if (dtype == RS_F16) UBM_BWD(RS_F16); else return RS_ERR_BAD_ARG;
#undef UBM_BWD
    RS_LAUNCH_CHECK();
    return RS_OK;
  }
  UB_DISPATCH(dtype, DT, {
    cudaError_t e = cudaFuncSetAttribute(ub_bwd_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ub_bwd_kernel<DT><<<grid, UB_THREADS, smem, st>>>(p, own_lse, g_pos, g_own, d_u, d_cols);
  });
  RS_LAUNCH_CHECK();
  return RS_OK;
}
