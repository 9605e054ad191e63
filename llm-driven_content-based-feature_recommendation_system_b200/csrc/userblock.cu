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

}  // namespace rs

using namespace rs;

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
  UB_DISPATCH(dtype, DT, {
    cudaError_t e = cudaFuncSetAttribute(ub_bwd_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ub_bwd_kernel<DT><<<grid, UB_THREADS, smem, st>>>(p, own_lse, g_pos, g_own, d_u, d_cols);
  });
  RS_LAUNCH_CHECK();
  return RS_OK;
}
