// F1: FM second-order interaction 0.5*sum_d((sum_f v)^2 - sum_f v^2), fused with the field gather,
// the linear term and the concatenated-embedding output that feeds the deep MLP.  Pure gather:
// HBM-bound, one warp per sample, the (sum, sum of squares) pair lives in registers.
// There is no reference implementation: follows deepctr-torch 0.2.9 `FM`.
#include "common.cuh"
#include "../../include/rs_twotower.h"

namespace rs {

#define FM_MAX_F 64

// k = 4*LPR floats per row; LPR lanes per row; 32/LPR fields per warp step
template <int LPR, int OD>
__global__ void __launch_bounds__(256) fm_fwd_kernel(const int64_t* __restrict__ ids,
                                                     const int64_t* __restrict__ offsets, int64_t B, int F,
                                                     const float* __restrict__ emb, const float* __restrict__ lin,
                                                     int64_t total_rows, float* __restrict__ fm,
                                                     void* __restrict__ concat, int* __restrict__ oob) {
  constexpr int K = 4 * LPR;
  constexpr int FPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int grp = lane / LPR, gl = lane % LPR;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t off0 = lane < F ? __ldg(offsets + lane) : 0;
  const int64_t off1 = lane + 32 < F ? __ldg(offsets + lane + 32) : 0;
  for (int64_t b = warp; b < B; b += nwarps) {
    // coalesced read of this sample's F ids, turned into absolute rows
    int64_t r0 = lane < F ? __ldg(ids + b * F + lane) + off0 : -1;
    int64_t r1 = lane + 32 < F ? __ldg(ids + b * F + lane + 32) + off1 : -1;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    float linacc = 0.f;
    for (int f0 = 0; f0 < F; f0 += FPW) {
      const int f = f0 + grp;
      const int64_t ra = __shfl_sync(0xffffffffu, r0, f & 31);
      const int64_t rb = __shfl_sync(0xffffffffu, r1, f & 31);
      const int64_t row = f < 32 ? ra : rb;
      if (f < F) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row >= 0 && row < total_rows) {
          v = ldg_f4(emb + row * K + 4 * gl);
          if (lin && gl == 0) linacc += __ldg(lin + row);
        } else if (oob && gl == 0) *oob = 1;
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
        if (concat) store4<OD>(concat, (b * F + f) * K + 4 * gl, v);
      }
    }
    // combine the FPW field groups (lanes with equal gl)
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
      s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
      s.z += __shfl_xor_sync(0xffffffffu, s.z, o); s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
      q.x += __shfl_xor_sync(0xffffffffu, q.x, o); q.y += __shfl_xor_sync(0xffffffffu, q.y, o);
      q.z += __shfl_xor_sync(0xffffffffu, q.z, o); q.w += __shfl_xor_sync(0xffffffffu, q.w, o);
    }
    float part = linacc;
    if (grp == 0) part += 0.5f * ((s.x * s.x - q.x) + (s.y * s.y - q.y) + (s.z * s.z - q.z) + (s.w * s.w - q.w));
    part = warp_sum(part);
    if (lane == 0) fm[b] = part;
  }
}

template <int LPR, int GD>
__global__ void __launch_bounds__(256) fm_bwd_kernel(const int64_t* __restrict__ ids,
                                                     const int64_t* __restrict__ offsets, int64_t B, int F,
                                                     const float* __restrict__ emb, const float* __restrict__ d_fm,
                                                     const void* __restrict__ d_concat, int64_t total_rows,
                                                     float* __restrict__ d_emb, float* __restrict__ d_lin) {
  constexpr int K = 4 * LPR;
  constexpr int FPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int grp = lane / LPR, gl = lane % LPR;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t off0 = lane < F ? __ldg(offsets + lane) : 0;
  const int64_t off1 = lane + 32 < F ? __ldg(offsets + lane + 32) : 0;
  for (int64_t b = warp; b < B; b += nwarps) {
    int64_t r0 = lane < F ? __ldg(ids + b * F + lane) + off0 : -1;
    int64_t r1 = lane + 32 < F ? __ldg(ids + b * F + lane + 32) + off1 : -1;
    const float g = d_fm ? __ldg(d_fm + b) : 0.f;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d_fm) {                                    // pass 1: S_b = sum_f v
      for (int f0 = 0; f0 < F; f0 += FPW) {
        const int f = f0 + grp;
        const int64_t ra = __shfl_sync(0xffffffffu, r0, f & 31);
        const int64_t rb = __shfl_sync(0xffffffffu, r1, f & 31);
        const int64_t row = f < 32 ? ra : rb;
        if (f < F && row >= 0 && row < total_rows) {
          const float4 v = ldg_f4(emb + row * K + 4 * gl);
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
      }
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) {
        s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
        s.z += __shfl_xor_sync(0xffffffffu, s.z, o); s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
      }
    }
    for (int f0 = 0; f0 < F; f0 += FPW) {          // pass 2: rows hit L1/L2
      const int f = f0 + grp;
      const int64_t ra = __shfl_sync(0xffffffffu, r0, f & 31);
      const int64_t rb = __shfl_sync(0xffffffffu, r1, f & 31);
      const int64_t row = f < 32 ? ra : rb;
      if (f < F && row >= 0 && row < total_rows) {
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d_fm) {
          const float4 v = ldg_f4(emb + row * K + 4 * gl);
          d = make_float4(g * (s.x - v.x), g * (s.y - v.y), g * (s.z - v.z), g * (s.w - v.w));
        }
        if (d_concat) {
          const float4 c = load4<GD>(d_concat, (b * F + f) * K + 4 * gl);
          d.x += c.x; d.y += c.y; d.z += c.z; d.w += c.w;
        }
        red_add_f4(d_emb + row * K + 4 * gl, d);
        if (d_lin && d_fm && gl == 0) atomicAdd(d_lin + row, g);
      }
    }
  }
}

}  // namespace rs

using namespace rs;

#define DISPATCH_DT(dt, NAME, ...)                                      \
  switch (dt) {                                                         \
    case RS_F32: { constexpr int NAME = RS_F32; __VA_ARGS__; break; }   \
    case RS_F16: { constexpr int NAME = RS_F16; __VA_ARGS__; break; }   \
    case RS_BF16: { constexpr int NAME = RS_BF16; __VA_ARGS__; break; } \
    default: return RS_ERR_BAD_ARG;                                     \
  }
#define DISPATCH_LPR(k, NAME, ...)                                   \
  switch (k) {                                                       \
    case 4: { constexpr int NAME = 1; __VA_ARGS__; break; }          \
    case 8: { constexpr int NAME = 2; __VA_ARGS__; break; }          \
    case 16: { constexpr int NAME = 4; __VA_ARGS__; break; }         \
    case 32: { constexpr int NAME = 8; __VA_ARGS__; break; }         \
    case 64: { constexpr int NAME = 16; __VA_ARGS__; break; }        \
    case 128: { constexpr int NAME = 32; __VA_ARGS__; break; }       \
    default: return RS_ERR_UNSUPPORTED;                              \
  }

extern "C" int rs_fm_fwd(const int64_t* ids, const int64_t* offsets, int64_t B, int64_t F, const float* emb,
                         int64_t k, const float* lin, int64_t total_rows, float* fm, void* concat, int concat_dtype,
                         int* oob_flag, void* stream) {
  if (B == 0) return RS_OK;
  if (!ids || !offsets || !emb || !fm || F <= 0 || F > FM_MAX_F) return RS_ERR_BAD_ARG;
  const int grid = grid_for_warps(B, 8, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (!concat) concat_dtype = RS_F32;
  DISPATCH_LPR(k, LPR, DISPATCH_DT(concat_dtype, OD, fm_fwd_kernel<LPR, OD><<<grid, 256, 0, st>>>(
      ids, offsets, B, (int)F, emb, lin, total_rows, fm, concat, oob_flag)));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_fm_bwd(const int64_t* ids, const int64_t* offsets, int64_t B, int64_t F, const float* emb,
                         int64_t k, const float* d_fm, const void* d_concat, int d_concat_dtype, int64_t total_rows,
                         float* d_emb, float* d_lin, void* stream) {
  if (B == 0) return RS_OK;
  if (!ids || !offsets || !emb || !d_emb || F <= 0 || F > FM_MAX_F || (!d_fm && !d_concat)) return RS_ERR_BAD_ARG;
  const int grid = grid_for_warps(B, 8, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_concat) d_concat_dtype = RS_F32;
  DISPATCH_LPR(k, LPR, DISPATCH_DT(d_concat_dtype, GD, fm_bwd_kernel<LPR, GD><<<grid, 256, 0, st>>>(
      ids, offsets, B, (int)F, emb, d_fm, d_concat, total_rows, d_emb, d_lin)));
  RS_LAUNCH_CHECK();
  return RS_OK;
}
