// Shared device helpers for the two-tower hot-path kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define RS_F32 0
#define RS_F16 1
#define RS_BF16 2

#define RS_OK 0
#define RS_ERR_BAD_ARG 10001
#define RS_ERR_UNSUPPORTED 10002
#define RS_ERR_WORKSPACE 10003

#define RS_MAX_TABLES 8
#define RS_NUM_SMS 148

// statistics only (how many kernels this library has launched in this process); see rs_launch_count()
extern "C" void rs_count_launches(int n);
#define RS_LAUNCH_CHECK_N(n)                   \
  do {                                         \
    cudaError_t e__ = cudaGetLastError();      \
    if (e__ != cudaSuccess) return (int)e__;   \
    rs_count_launches(n);                      \
  } while (0)
#define RS_LAUNCH_CHECK() RS_LAUNCH_CHECK_N(1)

namespace rs {

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming (read-once) 128-bit load: do not allocate in L1
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream_u2(void* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
// 128-bit vector reduction into global memory (sm_90+): one instruction adds four fp32 lanes
__device__ __forceinline__ void red_add_f4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
__device__ __forceinline__ float2 unpack_f16(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }

// Load 4 consecutive elements (element offset `off`, multiple of 4) of a row stored as DT, as fp32.
template <int DT>
__device__ __forceinline__ float4 load4(const void* base, int64_t off) {
  if constexpr (DT == RS_F32) {
    return ld_stream_f4(reinterpret_cast<const float*>(base) + off);
  } else {
    uint2 u = ld_stream_u2(reinterpret_cast<const uint16_t*>(base) + off);
    float2 a, b;
    if constexpr (DT == RS_BF16) { a = unpack_bf16(u.x); b = unpack_bf16(u.y); }
    else { a = unpack_f16(u.x); b = unpack_f16(u.y); }
    return make_float4(a.x, a.y, b.x, b.y);
  }
}
template <int DT>
__device__ __forceinline__ void store4(void* base, int64_t off, float4 v) {
  if constexpr (DT == RS_F32) {
    st_stream_f4(reinterpret_cast<float*>(base) + off, v);
  } else {
    uint2 u;
    if constexpr (DT == RS_BF16) { u.x = pack_bf16(v.x, v.y); u.y = pack_bf16(v.z, v.w); }
    else { u.x = pack_f16(v.x, v.y); u.y = pack_f16(v.z, v.w); }
    st_stream_u2(reinterpret_cast<uint16_t*>(base) + off, u);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// unfused multiply-then-add, each rounded to fp32: matches `acc += row * gate` of the reference bit for bit
__device__ __forceinline__ float4 mul_add_rn(float4 acc, float4 row, float g) {
  acc.x = __fadd_rn(acc.x, __fmul_rn(row.x, g));
  acc.y = __fadd_rn(acc.y, __fmul_rn(row.y, g));
  acc.z = __fadd_rn(acc.z, __fmul_rn(row.z, g));
  acc.w = __fadd_rn(acc.w, __fmul_rn(row.w, g));
  return acc;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 fma4(float4 acc, float4 v, float s) {
  acc.x = fmaf(v.x, s, acc.x); acc.y = fmaf(v.y, s, acc.y); acc.z = fmaf(v.z, s, acc.z); acc.w = fmaf(v.w, s, acc.w);
  return acc;
}
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

inline int grid_for_warps(int64_t n_warp_items, int warps_per_cta, int ctas_per_sm) {
  int64_t want = (n_warp_items + warps_per_cta - 1) / warps_per_cta;
  int64_t cap = (int64_t)RS_NUM_SMS * ctas_per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace rs
