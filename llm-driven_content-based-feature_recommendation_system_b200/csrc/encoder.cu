// Sequence-encoder body on PACKED valid tokens (U-rows' caller: tower_code/v1_refine_usertower.py:458-466).
//
// The reference runs nn.TransformerEncoder over the padded [B, L=50] grid with a causal + key-padding mask.  A
// valid position only ever attends to valid positions of its own sequence, and everything else in the layer is
// position-wise, so the valid rows of the output are a function of the valid rows alone: here the encoder runs on
// the ~25 % of the grid that is not padding (packed, `cu_seqlens` delimits the sequences).  The matmuls stay
// library GEMMs (stock nn.Linear in the reference as well); these kernels are the glue around them:
//
//   attn_fwd / attn_bwd      causal softmax(QK^T/sqrt(d))V for sequences of <= 64 tokens, head_dim 32, straight from
//                            the packed in_proj output [T, 3, H, 32] (no head transposes, no [B,H,L,L] bias tensor,
//                            no stored probabilities: the backward recomputes them from the saved row log-sum-exp);
//                            attention dropout from a counter-based hash (no mask tensor)
//   ln_fwd / ln_bwd          LayerNorm(128) with fp32 statistics, optional row gather (packing) on the way in,
//                            optional dropout on the way out, 16-bit or fp32 output
//   dropout_add fwd / bwd    x + dropout(y)            (residual stream stays fp32 as under the reference's autocast)
//   gelu_dropout fwd / bwd   dropout(gelu(z))
//
// All of it is HBM-bound row streaming: one warp per row / per (sequence, head), 128-bit accesses.
#include "common.cuh"
#include "../../include/rs_twotower.h"

namespace rs {

#define ENC_HD 32          // head dim
#define ENC_D 128          // model dim of the row kernels

__device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
  return h;
}
// Dropout epoch: one device-side counter mixed into every seed.  rs_rng_advance() increments it from the stream, so a
// CUDA graph that captured a train step (seeds are by-value kernel arguments, frozen at capture) still draws fresh
// masks on every replay, while the forward and backward kernels of one step see the same value.
__device__ unsigned long long g_rng_epoch = 0ull;
__global__ void rng_advance_kernel() { g_rng_epoch += 1ull; }
__device__ __forceinline__ uint64_t epoch_seed(uint64_t seed) {
  return seed ^ (*reinterpret_cast<volatile unsigned long long*>(&g_rng_epoch) * 0x9E3779B97F4A7C15ull);
}

// 32 uniform bits for element (a, b) of the stream `seed` (two rounds of an avalanche hash: ample for dropout)
__device__ __forceinline__ uint32_t rnd32(uint64_t seed, uint32_t a, uint32_t b) {
  uint32_t h = mix32(a * 0x9E3779B1u + (uint32_t)seed);
  return mix32(h ^ (b * 0x85EBCA77u + (uint32_t)(seed >> 32)));
}

template <int DT> __device__ __forceinline__ float4 ld4(const void* base, int64_t off) {
  if constexpr (DT == RS_F32) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off));
  else {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + off));
    float2 a, b;
    if constexpr (DT == RS_BF16) { a = unpack_bf16(u.x); b = unpack_bf16(u.y); }
    else { a = unpack_f16(u.x); b = unpack_f16(u.y); }
    return make_float4(a.x, a.y, b.x, b.y);
  }
}
template <int DT> __device__ __forceinline__ void st4(void* base, int64_t off, float4 v) {
  if constexpr (DT == RS_F32) *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = v;
  else {
    uint2 u;
    if constexpr (DT == RS_BF16) { u.x = pack_bf16(v.x, v.y); u.y = pack_bf16(v.z, v.w); }
    else { u.x = pack_f16(v.x, v.y); u.y = pack_f16(v.z, v.w); }
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + off) = u;
  }
}

template <int DT> __device__ __forceinline__ float round_dt(float x) {
  if constexpr (DT == RS_BF16) return __bfloat162float(__float2bfloat16_rn(x));
  else if constexpr (DT == RS_F16) return __half2float(__float2half_rn(x));
  else return x;
}

// element type as stored (only used to write zeros)
template <int DT> struct DTStoreT { typedef float type; };
template <> struct DTStoreT<RS_F16> { typedef __half type; };
template <> struct DTStoreT<RS_BF16> { typedef __nv_bfloat16 type; };
template <int DT> using DTStore = typename DTStoreT<DT>::type;

// ------------------------------------------------------------------------------------------------ attention
struct AttnParams {
  const int* cu;            // [n_seq + 1] token offsets
  const float* bias;        // [3*H*32] in_proj bias, added on load (NULL: none)
  int64_t n_seq;
  int H, max_len;
  int64_t zero_from;        // sequences b >= zero_from are fully masked queries: output 0, no gradient
  int64_t full_to;          // sequences in [full_to, zero_from) are only read at ONE token each (attn_one_* kernels):
                            // the main kernels skip them
  int short_split;          // backward: one-tile sequences (<= 16 tokens) are left to attn3_bwd_short_kernel
  const int64_t* one_row;   // [zero_from - full_to] packed row of that token (NULL: the last token; a row outside the
                            // sequence: none -- the sequence produces zeros only)
  float scale;
  uint32_t drop_thresh;     // keep iff rnd32 >= thresh   (0: no dropout)
  float inv_keep;
  uint64_t seed;
};


// a lane's slice of a row: N consecutive values (N = 4, 8 or 16)
template <int DT, int N>
__device__ __forceinline__ void load_slice(float (&r)[N], const void* src, int64_t off, const float* bias, int64_t boff) {
#pragma unroll
  for (int q = 0; q < N / 4; ++q) {
    float4 v = ld4<DT>(src, off + 4 * q);
    if (bias) { const float4 b4 = ldg_f4(bias + boff + 4 * q); v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w; }
    r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
  }
}
// sum over the R consecutive lanes that share a row
template <int R> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = R / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ================================================================================================ attention, SIMT (fp32 operands)
// One warp per (sequence, head), R lanes per query row (each lane owns 32/R of the 32 dims):
//   * keys (forward, dQ phase) / queries (dK-dV phase) are staged in blocks of 16 rows: 4 KB of shared memory per warp
//     whatever the sequence length, online softmax across blocks;
//   * two keys (queries) per iteration -- independent dot/exp chains -- and a branch-free running-max update;
//   * packed fp32 pairs (FFMA2) for every dot product and accumulation.
#define ATT_BLK 16
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk2(float a, float b) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t ffma2(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t fmul2(f2_t a, f2_t b) { f2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2_t fadd2(f2_t a, f2_t b) { f2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// stage `n` rows (<= 16) as fp32 [16][32]; the row after the last one is zero-filled when n is odd (pairs are processed)
template <int DT>
__device__ __forceinline__ void stage_block(float* dst, const void* src, int64_t row0, int n, int64_t row_stride,
                                            int64_t col0, const float* bias, int lane) {
  const int sub = lane >> 3, d4 = (lane & 7) * 4;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) b4 = ldg_f4(bias + col0 + d4);
  const int n2 = (n + 1) & ~1;
#pragma unroll
  for (int j0 = 0; j0 < ATT_BLK; j0 += 4) {
    const int j = j0 + sub;
    if (j < n) {
      const float4 v = ld4<DT>(src, (row0 + j) * row_stride + col0 + d4);
      *reinterpret_cast<float4*>(dst + j * ENC_HD + d4) = make_float4(v.x + b4.x, v.y + b4.y, v.z + b4.z, v.w + b4.w);
    } else if (j < n2) {
      *reinterpret_cast<float4*>(dst + j * ENC_HD + d4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
template <int N> __device__ __forceinline__ void load_pairs(f2_t (&r)[N / 2], const float* s) {
#pragma unroll
  for (int q = 0; q < N / 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(s + 4 * q);
    r[2 * q] = pk2(v.x, v.y); r[2 * q + 1] = pk2(v.z, v.w);
  }
}
template <int N> __device__ __forceinline__ float dot_pairs(const f2_t (&a)[N / 2], const f2_t (&b)[N / 2]) {
  f2_t s0 = fmul2(a[0], b[0]), s1 = fmul2(a[1], b[1]);
#pragma unroll
  for (int q = 2; q < N / 2; q += 2) { s0 = ffma2(a[q], b[q], s0); s1 = ffma2(a[q + 1], b[q + 1], s1); }
  float x, y;
  upk2(fadd2(s0, s1), x, y);
  return x + y;
}
template <int N> __device__ __forceinline__ float dot_smem(const f2_t (&a)[N / 2], const float* s) {
  f2_t b[N / 2];
  load_pairs<N>(b, s);
  return dot_pairs<N>(a, b);
}
template <int N> __device__ __forceinline__ void pack_slice(f2_t (&r)[N / 2], const float (&x)[N], float scale) {
#pragma unroll
  for (int q = 0; q < N / 2; ++q) r[q] = pk2(x[2 * q] * scale, x[2 * q + 1] * scale);
}
template <int DT, int N>
__device__ __forceinline__ void store_pairs(void* dst, int64_t off, const f2_t (&r)[N / 2], float s) {
#pragma unroll
  for (int q = 0; q < N / 4; ++q) {
    float a, b, c, d;
    upk2(r[2 * q], a, b);
    upk2(r[2 * q + 1], c, d);
    st4<DT>(dst, off + 4 * q, make_float4(a * s, b * s, c * s, d * s));
  }
}

template <int DT, int R>
__device__ __forceinline__ void attn2_fwd_item(const void* __restrict__ qkv, const AttnParams& p, float* sK, float* sV,
                                               int64_t t0, int len, int h, int lane, void* __restrict__ out,
                                               float* __restrict__ lse) {
  constexpr int N = ENC_HD / R, RPP = 32 / R, NP = N / 2;
  const int sub = lane % R, rl = lane / R, d0 = sub * N;
  const int64_t rs_ = 3 * (int64_t)p.H * ENC_HD, os_ = (int64_t)p.H * ENC_HD;
  for (int r0 = 0; r0 < len; r0 += RPP) {
    const int i = r0 + rl;
    const bool act = i < len;
    f2_t q[NP], acc[NP];
    {
      float qf[N];
#pragma unroll
      for (int d = 0; d < N; ++d) qf[d] = 0.f;
      if (act) load_slice<DT, N>(qf, qkv, (t0 + i) * rs_ + h * ENC_HD + d0, p.bias, h * ENC_HD + d0);
      pack_slice<N>(q, qf, p.scale);
    }
#pragma unroll
    for (int d = 0; d < NP; ++d) acc[d] = 0ull;
    float m = -INFINITY, l = 0.f;
    const uint32_t rid = (uint32_t)((t0 + i) * p.H + h);
    const int jend = min(len, r0 + RPP);
    for (int kb = 0; kb < jend; kb += ATT_BLK) {
      const int nk = min(ATT_BLK, jend - kb);
      __syncwarp();
      stage_block<DT>(sK, qkv, t0 + kb, nk, rs_, os_ + h * ENC_HD, p.bias, lane);
      stage_block<DT>(sV, qkv, t0 + kb, nk, rs_, 2 * os_ + h * ENC_HD, p.bias, lane);
      __syncwarp();
      for (int jj = 0; jj < nk; jj += 2) {
        const int j0 = kb + jj;
        float s0 = group_sum<R>(dot_smem<N>(q, sK + jj * ENC_HD + d0));
        float s1 = group_sum<R>(dot_smem<N>(q, sK + (jj + 1) * ENC_HD + d0));
        const bool ok0 = act && j0 <= i, ok1 = act && (jj + 1 < nk) && (j0 + 1 <= i);
        s0 = ok0 ? s0 : -INFINITY;
        s1 = ok1 ? s1 : -INFINITY;
        const float mn = fmaxf(m, fmaxf(s0, s1));
        const float mu = (mn == -INFINITY) ? 0.f : mn;
        const float c = __expf(m - mu), p0 = __expf(s0 - mu), p1 = __expf(s1 - mu);
        l = fmaf(l, c, p0 + p1);
        m = mn;
        float k0 = p0, k1 = p1;
        if (p.drop_thresh) {
          k0 = (rnd32(p.seed, rid, (uint32_t)j0) >= p.drop_thresh) ? p0 * p.inv_keep : 0.f;
          k1 = (rnd32(p.seed, rid, (uint32_t)(j0 + 1)) >= p.drop_thresh) ? p1 * p.inv_keep : 0.f;
        }
        const f2_t c2 = pk2(c, c), a0 = pk2(k0, k0), a1 = pk2(k1, k1);
        f2_t v0[NP], v1[NP];
        load_pairs<N>(v0, sV + jj * ENC_HD + d0);
        load_pairs<N>(v1, sV + (jj + 1) * ENC_HD + d0);
#pragma unroll
        for (int d = 0; d < NP; ++d) acc[d] = ffma2(a1, v1[d], ffma2(a0, v0[d], fmul2(acc[d], c2)));
      }
    }
    if (act) {
      store_pairs<DT, N>(out, (t0 + i) * os_ + h * ENC_HD + d0, acc, 1.f / l);
      if (sub == 0) lse[(t0 + i) * p.H + h] = m + __logf(l);
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(256, 3) attn2_fwd_kernel(const void* __restrict__ qkv, AttnParams p,
                                                        void* __restrict__ out, float* __restrict__ lse) {
  __shared__ __align__(16) float smem[8 * 2 * ATT_BLK * ENC_HD];
  p.seed = epoch_seed(p.seed);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  float* sK = smem + warp * 2 * ATT_BLK * ENC_HD;
  float* sV = sK + ATT_BLK * ENC_HD;
  const int64_t n_items = p.n_seq * p.H;
  const int64_t os_ = (int64_t)p.H * ENC_HD;
  for (int64_t item = (int64_t)blockIdx.x * wpc + warp; item < n_items; item += (int64_t)gridDim.x * wpc) {
    const int64_t b = item / p.H;
    const int h = (int)(item % p.H);
    const int64_t t0 = __ldg(p.cu + b);
    const int len = min((int)(__ldg(p.cu + b + 1) - t0), p.max_len);
    if (b >= p.full_to && b < p.zero_from) continue;    // single-row sequences: attn_one_fwd_kernel
    if (b >= p.zero_from) {                             // a query whose every key is masked (see rs_twotower.h)
      for (int i = 0; i < len; ++i) {
        reinterpret_cast<DTStore<DT>*>(out)[(t0 + i) * os_ + h * ENC_HD + lane] = DTStore<DT>(0);
        if (lane == 0) lse[(t0 + i) * p.H + h] = 0.f;
      }
      continue;
    }
    if (len <= 8) attn2_fwd_item<DT, 4>(qkv, p, sK, sV, t0, len, h, lane, out, lse);
    else attn2_fwd_item<DT, 2>(qkv, p, sK, sV, t0, len, h, lane, out, lse);
  }
}

// dQ phase: rows = queries; also leaves lse_i and delta_i = <dO_i, O_i> of every row in shared memory for the dK/dV phase
template <int DT, int R>
__device__ __forceinline__ void attn2_bwd_q(const void* __restrict__ qkv, const void* __restrict__ d_out,
                                            const void* __restrict__ out, const float* __restrict__ lse,
                                            const AttnParams& p, float* sA, float* sB, float* sLse, float* sDelta,
                                            int64_t t0, int len, int h, int lane, void* __restrict__ d_qkv) {
  constexpr int N = ENC_HD / R, RPP = 32 / R, NP = N / 2;
  const int sub = lane % R, rl = lane / R, d0 = sub * N;
  const int64_t rs_ = 3 * (int64_t)p.H * ENC_HD, os_ = (int64_t)p.H * ENC_HD;
  for (int r0 = 0; r0 < len; r0 += RPP) {
    const int i = r0 + rl;
    const bool act = i < len;
    f2_t q[NP], g[NP], dq[NP];
    float di, li = 0.f;
    {
      float qf[N], gf[N], of[N];
#pragma unroll
      for (int d = 0; d < N; ++d) { qf[d] = 0.f; gf[d] = 0.f; of[d] = 0.f; }
      if (act) {
        load_slice<DT, N>(qf, qkv, (t0 + i) * rs_ + h * ENC_HD + d0, p.bias, h * ENC_HD + d0);
        load_slice<DT, N>(gf, d_out, (t0 + i) * os_ + h * ENC_HD + d0, nullptr, 0);
        load_slice<DT, N>(of, out, (t0 + i) * os_ + h * ENC_HD + d0, nullptr, 0);
        li = __ldg(lse + (t0 + i) * p.H + h);
      }
      float part = 0.f;
#pragma unroll
      for (int d = 0; d < N; ++d) part = fmaf(gf[d], of[d], part);
      di = group_sum<R>(part);
      pack_slice<N>(q, qf, p.scale);
      pack_slice<N>(g, gf, 1.f);
    }
    if (act && sub == 0) { sLse[i] = li; sDelta[i] = di; }
#pragma unroll
    for (int d = 0; d < NP; ++d) dq[d] = 0ull;
    const uint32_t rid = (uint32_t)((t0 + i) * p.H + h);
    const int jend = min(len, r0 + RPP);
    for (int kb = 0; kb < jend; kb += ATT_BLK) {
      const int nk = min(ATT_BLK, jend - kb);
      __syncwarp();
      stage_block<DT>(sA, qkv, t0 + kb, nk, rs_, os_ + h * ENC_HD, p.bias, lane);        // K
      stage_block<DT>(sB, qkv, t0 + kb, nk, rs_, 2 * os_ + h * ENC_HD, p.bias, lane);    // V
      __syncwarp();
      for (int jj = 0; jj < nk; jj += 2) {
        const int j0 = kb + jj;
        f2_t k0[NP], k1[NP];
        load_pairs<N>(k0, sA + jj * ENC_HD + d0);
        load_pairs<N>(k1, sA + (jj + 1) * ENC_HD + d0);
        const float s0 = group_sum<R>(dot_pairs<N>(q, k0)), s1 = group_sum<R>(dot_pairs<N>(q, k1));
        float dp0 = group_sum<R>(dot_smem<N>(g, sB + jj * ENC_HD + d0));
        float dp1 = group_sum<R>(dot_smem<N>(g, sB + (jj + 1) * ENC_HD + d0));
        const bool ok0 = act && j0 <= i, ok1 = act && (jj + 1 < nk) && (j0 + 1 <= i);
        const float pr0 = ok0 ? __expf(s0 - li) : 0.f, pr1 = ok1 ? __expf(s1 - li) : 0.f;
        if (p.drop_thresh) {
          dp0 = (rnd32(p.seed, rid, (uint32_t)j0) >= p.drop_thresh) ? dp0 * p.inv_keep : 0.f;
          dp1 = (rnd32(p.seed, rid, (uint32_t)(j0 + 1)) >= p.drop_thresh) ? dp1 * p.inv_keep : 0.f;
        }
        const float e0 = pr0 * (dp0 - di), e1 = pr1 * (dp1 - di);
        const f2_t a0 = pk2(e0, e0), a1 = pk2(e1, e1);
#pragma unroll
        for (int d = 0; d < NP; ++d) dq[d] = ffma2(a1, k1[d], ffma2(a0, k0[d], dq[d]));
      }
    }
    if (act) store_pairs<DT, N>(d_qkv, (t0 + i) * rs_ + h * ENC_HD + d0, dq, p.scale);
  }
}

// dK / dV phase: rows = keys, query blocks staged (Q in sA, dO in sB)
template <int DT, int R>
__device__ __forceinline__ void attn2_bwd_kv(const void* __restrict__ qkv, const void* __restrict__ d_out,
                                             const AttnParams& p, float* sA, float* sB, const float* sLse,
                                             const float* sDelta, int64_t t0, int len, int h, int lane,
                                             void* __restrict__ d_qkv) {
  constexpr int N = ENC_HD / R, RPP = 32 / R, NP = N / 2;
  const int sub = lane % R, rl = lane / R, d0 = sub * N;
  const int64_t rs_ = 3 * (int64_t)p.H * ENC_HD, os_ = (int64_t)p.H * ENC_HD;
  for (int r0 = 0; r0 < len; r0 += RPP) {
    const int j = r0 + rl;
    const bool act = j < len;
    f2_t k[NP], v[NP], dk[NP], dv[NP];
    {
      float kf[N], vf[N];
#pragma unroll
      for (int d = 0; d < N; ++d) { kf[d] = 0.f; vf[d] = 0.f; }
      if (act) {
        load_slice<DT, N>(kf, qkv, (t0 + j) * rs_ + os_ + h * ENC_HD + d0, p.bias, os_ + h * ENC_HD + d0);
        load_slice<DT, N>(vf, qkv, (t0 + j) * rs_ + 2 * os_ + h * ENC_HD + d0, p.bias, 2 * os_ + h * ENC_HD + d0);
      }
      pack_slice<N>(k, kf, p.scale);
      pack_slice<N>(v, vf, 1.f);
    }
#pragma unroll
    for (int d = 0; d < NP; ++d) { dk[d] = 0ull; dv[d] = 0ull; }
    for (int ib = (r0 / ATT_BLK) * ATT_BLK; ib < len; ib += ATT_BLK) {
      const int nq = min(ATT_BLK, len - ib);
      __syncwarp();
      stage_block<DT>(sA, qkv, t0 + ib, nq, rs_, h * ENC_HD, p.bias, lane);              // Q
      stage_block<DT>(sB, d_out, t0 + ib, nq, os_, h * ENC_HD, nullptr, lane);           // dO
      __syncwarp();
#pragma unroll 2
      for (int ii = 0; ii < nq; ++ii) {
        const int i = ib + ii;
        f2_t qr[NP], gr[NP];
        load_pairs<N>(qr, sA + ii * ENC_HD + d0);
        load_pairs<N>(gr, sB + ii * ENC_HD + d0);
        const float s = group_sum<R>(dot_pairs<N>(k, qr));
        float dp = group_sum<R>(dot_pairs<N>(v, gr));
        const bool ok = act && i >= j;
        const float pr = ok ? __expf(s - sLse[i]) : 0.f;
        float pk = pr;
        if (p.drop_thresh) {
          const bool keep = rnd32(p.seed, (uint32_t)((t0 + i) * p.H + h), (uint32_t)j) >= p.drop_thresh;
          dp = keep ? dp * p.inv_keep : 0.f;
          pk = keep ? pr * p.inv_keep : 0.f;
        }
        const float e = pr * (dp - sDelta[i]);
        const f2_t a = pk2(pk, pk), b = pk2(e, e);
#pragma unroll
        for (int d = 0; d < NP; ++d) { dv[d] = ffma2(a, gr[d], dv[d]); dk[d] = ffma2(b, qr[d], dk[d]); }
      }
    }
    if (act) {
      store_pairs<DT, N>(d_qkv, (t0 + j) * rs_ + os_ + h * ENC_HD + d0, dk, p.scale);
      store_pairs<DT, N>(d_qkv, (t0 + j) * rs_ + 2 * os_ + h * ENC_HD + d0, dv, 1.f);
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(256, 2) attn2_bwd_kernel(const void* __restrict__ qkv, const void* __restrict__ d_out,
                                                        const void* __restrict__ out, const float* __restrict__ lse,
                                                        AttnParams p, void* __restrict__ d_qkv) {
  __shared__ __align__(16) float smem[8 * (2 * ATT_BLK * ENC_HD + 128)];
  p.seed = epoch_seed(p.seed);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  float* sA = smem + warp * (2 * ATT_BLK * ENC_HD + 128);
  float* sB = sA + ATT_BLK * ENC_HD;
  float* sLse = sB + ATT_BLK * ENC_HD;
  float* sDelta = sLse + 64;
  const int64_t n_items = p.n_seq * p.H;
  const int64_t rs_ = 3 * (int64_t)p.H * ENC_HD, os_ = (int64_t)p.H * ENC_HD;
  for (int64_t item = (int64_t)blockIdx.x * wpc + warp; item < n_items; item += (int64_t)gridDim.x * wpc) {
    const int64_t b = item / p.H;
    const int h = (int)(item % p.H);
    const int64_t t0 = __ldg(p.cu + b);
    const int len = min((int)(__ldg(p.cu + b + 1) - t0), p.max_len);
    if (b >= p.full_to && b < p.zero_from) continue;    // single-row sequences: attn_one_bwd_kernel
    if (b >= p.zero_from) {
      for (int i = 0; i < len; ++i)
        for (int c = 0; c < 3; ++c)
          reinterpret_cast<DTStore<DT>*>(d_qkv)[(t0 + i) * rs_ + c * os_ + h * ENC_HD + lane] = DTStore<DT>(0);
      continue;
    }
    __syncwarp();                                       // the previous item's lse / delta are no longer read
    if (len <= 8) {
      attn2_bwd_q<DT, 4>(qkv, d_out, out, lse, p, sA, sB, sLse, sDelta, t0, len, h, lane, d_qkv);
      attn2_bwd_kv<DT, 4>(qkv, d_out, p, sA, sB, sLse, sDelta, t0, len, h, lane, d_qkv);
    } else {
      attn2_bwd_q<DT, 2>(qkv, d_out, out, lse, p, sA, sB, sLse, sDelta, t0, len, h, lane, d_qkv);
      attn2_bwd_kv<DT, 2>(qkv, d_out, p, sA, sB, sLse, sDelta, t0, len, h, lane, d_qkv);
    }
  }
}

#include "attn_mma.cuh"

// ------------------------------------------------------------------------------------------------ single-row attention
// Sequences whose output is only read at ONE token (the second dropout view in the last encoder layer: it feeds nothing
// but its DuoRec row, v1_usertower_train.py:789,830-842 -- position len-1 of the LEFT-padded grid, i.e. some token in
// the middle of the sequence).  One query against <= 64 keys is a matrix-vector product: one warp per (sequence,
// head), lane j scores keys j and j + 32 against the broadcast query, lane d then accumulates output dimension d over
// the keys (coalesced 64-byte V reads).  The other rows of `out` are written as zeros (they are never read where it
// matters: rows with loss weight 0 at most).  Same bias / rounding convention as the tile kernels (bias added in the
// operand precision); probabilities stay fp32.
template <int DT> __device__ __forceinline__ float ld1(const void* base, int64_t off) {
  if constexpr (DT == RS_F32) return __ldg(reinterpret_cast<const float*>(base) + off);
  else if constexpr (DT == RS_BF16) return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(base) + off));
  else return __half2float(__ldg(reinterpret_cast<const __half*>(base) + off));
}
template <int DT> __device__ __forceinline__ void st1(void* base, int64_t off, float v) {
  reinterpret_cast<DTStore<DT>*>(base)[off] = DTStore<DT>(v);
}
// row `tok`, 32 dims starting at column `col0`, + bias (rounded to DT like a Linear's output) -> r[32] in every lane
template <int DT>
__device__ __forceinline__ void load_row32(float (&r)[ENC_HD], const void* src, int64_t off, const float* bias, int64_t boff) {
#pragma unroll
  for (int c = 0; c < ENC_HD / 4; ++c) {
    float4 v = ld4<DT>(src, off + 4 * c);
    if (bias) {
      const float4 b4 = ldg_f4(bias + boff + 4 * c);
      v = make_float4(round_dt<DT>(v.x + b4.x), round_dt<DT>(v.y + b4.y), round_dt<DT>(v.z + b4.z), round_dt<DT>(v.w + b4.w));
    }
    r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
  }
}
// the one query row of sequence `b` (index within the sequence), or -1
__device__ __forceinline__ int one_row_of(const AttnParams& p, int64_t b, int64_t t0, int len) {
  if (!p.one_row) return len - 1;
  const int64_t r = __ldg(p.one_row + (b - p.full_to)) - t0;
  return (r >= 0 && r < len) ? (int)r : -1;
}

template <int DT>
__global__ void __launch_bounds__(256) attn_one_fwd_kernel(const void* __restrict__ qkv, AttnParams p,
                                                           void* __restrict__ out, float* __restrict__ lse) {
  __shared__ float sP[8][64];
  p.seed = epoch_seed(p.seed);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int64_t n_items = (p.zero_from - p.full_to) * p.H;
  const int64_t rs_ = 3 * (int64_t)p.H * ENC_HD, os_ = (int64_t)p.H * ENC_HD;
  for (int64_t item = (int64_t)blockIdx.x * wpc + warp; item < n_items; item += (int64_t)gridDim.x * wpc) {
    const int64_t b = p.full_to + item / p.H;
    const int h = (int)(item % p.H);
    const int64_t t0 = __ldg(p.cu + b);
    const int len = min((int)(__ldg(p.cu + b + 1) - t0), p.max_len);
    if (len <= 0) continue;
    const int i = one_row_of(p, b, t0, len);
    for (int j = 0; j < len; ++j) {
      if (j == i) continue;
      st1<DT>(out, (t0 + j) * os_ + h * ENC_HD + lane, 0.f);
      if (lane == 0) lse[(t0 + j) * p.H + h] = 0.f;
    }
    if (i < 0) continue;
    float q[ENC_HD];
    load_row32<DT>(q, qkv, (t0 + i) * rs_ + h * ENC_HD, p.bias, h * ENC_HD);
    float sc[2];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int j = lane + 32 * kk;
      sc[kk] = -INFINITY;
      if (j <= i) {
        float k[ENC_HD];
        load_row32<DT>(k, qkv, (t0 + j) * rs_ + os_ + h * ENC_HD, p.bias, os_ + h * ENC_HD);
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < ENC_HD; ++c) d = fmaf(q[c], k[c], d);
        sc[kk] = d * p.scale;
      }
    }
    const float m = warp_max(fmaxf(sc[0], sc[1]));
    const float e0 = __expf(sc[0] - m), e1 = __expf(sc[1] - m);        // (-inf -> 0)
    const float l = warp_sum(e0 + e1);
    const float inv = 1.f / l;
    __syncwarp();
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int j = lane + 32 * kk;
      float pr = (kk ? e1 : e0) * inv;
      if (p.drop_thresh) pr = (rnd32(p.seed, (uint32_t)((t0 + i) * p.H + h), (uint32_t)j) >= p.drop_thresh) ? pr * p.inv_keep : 0.f;
      sP[warp][j] = pr;
    }
    __syncwarp();
    const float bv = p.bias ? __ldg(p.bias + 2 * os_ + h * ENC_HD + lane) : 0.f;
    float acc = 0.f;
    for (int j = 0; j <= i; ++j) {
      float v = ld1<DT>(qkv, (t0 + j) * rs_ + 2 * os_ + h * ENC_HD + lane);
      if (p.bias) v = round_dt<DT>(v + bv);
      acc = fmaf(sP[warp][j], v, acc);
    }
    st1<DT>(out, (t0 + i) * os_ + h * ENC_HD + lane, acc);
    if (lane == 0) lse[(t0 + i) * p.H + h] = m + __logf(l);
  }
}

// d_qkv of the same sequences from the gradient of their one row alone: dq (that row; zeros elsewhere), dk_j = ds_j q,
// dv_j = p_j dO for the keys j <= i, zeros behind it
template <int DT>
__global__ void __launch_bounds__(256) attn_one_bwd_kernel(const void* __restrict__ qkv, const void* __restrict__ d_out,
                                                           const void* __restrict__ out, const float* __restrict__ lse,
                                                           AttnParams p, void* __restrict__ d_qkv) {
  __shared__ float sS[8][2][64];                       // [0] ds_j   [1] dropped p_j
  p.seed = epoch_seed(p.seed);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int64_t n_items = (p.zero_from - p.full_to) * p.H;
  const int64_t rs_ = 3 * (int64_t)p.H * ENC_HD, os_ = (int64_t)p.H * ENC_HD;
  for (int64_t item = (int64_t)blockIdx.x * wpc + warp; item < n_items; item += (int64_t)gridDim.x * wpc) {
    const int64_t b = p.full_to + item / p.H;
    const int h = (int)(item % p.H);
    const int64_t t0 = __ldg(p.cu + b);
    const int len = min((int)(__ldg(p.cu + b + 1) - t0), p.max_len);
    if (len <= 0) continue;
    const int i = one_row_of(p, b, t0, len);
    for (int j = i + 1; j < len; ++j) {                // keys behind the query (all keys when there is no query)
      const int64_t row = (t0 + j) * rs_ + h * ENC_HD + lane;
      st1<DT>(d_qkv, row, 0.f); st1<DT>(d_qkv, row + os_, 0.f); st1<DT>(d_qkv, row + 2 * os_, 0.f);
    }
    if (i < 0) continue;
    float q[ENC_HD], go[ENC_HD];
    load_row32<DT>(q, qkv, (t0 + i) * rs_ + h * ENC_HD, p.bias, h * ENC_HD);
    load_row32<DT>(go, d_out, (t0 + i) * os_ + h * ENC_HD, nullptr, 0);
    const float g_d = ld1<DT>(d_out, (t0 + i) * os_ + h * ENC_HD + lane);
    const float o_d = ld1<DT>(out, (t0 + i) * os_ + h * ENC_HD + lane);
    const float delta = warp_sum(g_d * o_d);
    const float li = __ldg(lse + (t0 + i) * p.H + h);
    __syncwarp();                                      // the previous item's sS reads are done
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int j = lane + 32 * kk;
      float ds = 0.f, pd = 0.f;
      if (j <= i) {
        float k[ENC_HD];
        load_row32<DT>(k, qkv, (t0 + j) * rs_ + os_ + h * ENC_HD, p.bias, os_ + h * ENC_HD);
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < ENC_HD; ++c) d = fmaf(q[c], k[c], d);
        const float pr = __expf(d * p.scale - li);
        float v[ENC_HD];
        load_row32<DT>(v, qkv, (t0 + j) * rs_ + 2 * os_ + h * ENC_HD, p.bias, 2 * os_ + h * ENC_HD);
        float dp = 0.f;
#pragma unroll
        for (int c = 0; c < ENC_HD; ++c) dp = fmaf(go[c], v[c], dp);
        pd = pr;
        if (p.drop_thresh) {
          const bool keep = rnd32(p.seed, (uint32_t)((t0 + i) * p.H + h), (uint32_t)j) >= p.drop_thresh;
          dp = keep ? dp * p.inv_keep : 0.f;
          pd = keep ? pr * p.inv_keep : 0.f;
        }
        ds = pr * (dp - delta);
      }
      sS[warp][0][j] = ds;
      sS[warp][1][j] = pd;
    }
    __syncwarp();
    // lane = feature d: dq_d = scale sum_j ds_j k_jd ; rows j: dk_jd = scale ds_j q_d, dv_jd = pd_j dO_d
    const float bk = p.bias ? __ldg(p.bias + os_ + h * ENC_HD + lane) : 0.f;
    const float bq = p.bias ? __ldg(p.bias + h * ENC_HD + lane) : 0.f;
    float q_d = ld1<DT>(qkv, (t0 + i) * rs_ + h * ENC_HD + lane);
    if (p.bias) q_d = round_dt<DT>(q_d + bq);
    float dq = 0.f;
    for (int j = 0; j <= i; ++j) {
      float k = ld1<DT>(qkv, (t0 + j) * rs_ + os_ + h * ENC_HD + lane);
      if (p.bias) k = round_dt<DT>(k + bk);
      const float ds = sS[warp][0][j];
      dq = fmaf(ds, k, dq);
      const int64_t row = (t0 + j) * rs_ + h * ENC_HD + lane;
      if (j < i) st1<DT>(d_qkv, row, 0.f);
      st1<DT>(d_qkv, row + os_, ds * q_d * p.scale);
      st1<DT>(d_qkv, row + 2 * os_, sS[warp][1][j] * g_d);
    }
    st1<DT>(d_qkv, (t0 + i) * rs_ + h * ENC_HD + lane, dq * p.scale);
  }
}

// out[c] = sum_r x[r, c]  -- column sums of a [n_rows, n_cols] matrix (bias gradients), two deterministic stages:
// thread t of a CTA owns column group (t % (n_cols/4)) and walks rows t / (n_cols/4), + row-groups ...; per-CTA partials
template <int DT>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const void* __restrict__ x, int64_t n_rows, int n_cols,
                                                             float* __restrict__ part /*[grid][n_cols]*/) {
  extern __shared__ float4 cred[];                 // [rows_per_iter][n_cols/4]
  const int c4n = n_cols >> 2;
  const int rpi = blockDim.x / c4n;                // rows handled per iteration by this CTA
  const int c4 = threadIdx.x % c4n, rr = threadIdx.x / c4n;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rr < rpi) {
    const int64_t step = (int64_t)gridDim.x * rpi;
    int64_t r = (int64_t)blockIdx.x * rpi + rr;
    for (; r + 3 * step < n_rows; r += 4 * step) {            // four independent row loads in flight, fixed add order
      const float4 v0 = ld4<DT>(x, r * n_cols + 4 * c4), v1 = ld4<DT>(x, (r + step) * n_cols + 4 * c4);
      const float4 v2 = ld4<DT>(x, (r + 2 * step) * n_cols + 4 * c4), v3 = ld4<DT>(x, (r + 3 * step) * n_cols + 4 * c4);
      s.x += (v0.x + v1.x) + (v2.x + v3.x); s.y += (v0.y + v1.y) + (v2.y + v3.y);
      s.z += (v0.z + v1.z) + (v2.z + v3.z); s.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; r < n_rows; r += step) {
      const float4 v = ld4<DT>(x, r * n_cols + 4 * c4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  if (rr < rpi) cred[rr * c4n + c4] = s;
  __syncthreads();
  if (rr == 0) {
    for (int k = 1; k < rpi; ++k) {
      const float4 t = cred[k * c4n + c4];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    *reinterpret_cast<float4*>(part + (int64_t)blockIdx.x * n_cols + 4 * c4) = s;
  }
}
// ------------------------------------------------------------------------------------------------ LayerNorm(128)
// y[r] = dropout(LN(x[index ? index[r] : r])) ; one warp per row, lane owns 4 consecutive features
template <int DTI, int DTO>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const void* __restrict__ x, const int64_t* __restrict__ index,
                                                     int64_t n_rows, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float eps, uint32_t drop_thresh,
                                                     float inv_keep, uint64_t seed, void* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd) {
  seed = epoch_seed(seed);
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float4 w4 = ldg_f4(w + 4 * lane), b4 = ldg_f4(bias + 4 * lane);
  for (int64_t r = warp; r < n_rows; r += nw) {
    const int64_t src = index ? __ldg(index + r) : r;
    const float4 v = ld4<DTI>(x, src * ENC_D + 4 * lane);
    const float mu = warp_sum(v.x + v.y + v.z + v.w) * (1.f / ENC_D);
    const float4 c = make_float4(v.x - mu, v.y - mu, v.z - mu, v.w - mu);
    const float var = warp_sum(c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w) * (1.f / ENC_D);
    const float rs = rsqrtf(var + eps);
    float4 o = make_float4(c.x * rs * w4.x + b4.x, c.y * rs * w4.y + b4.y, c.z * rs * w4.z + b4.z, c.w * rs * w4.w + b4.w);
    if (drop_thresh) {
      const uint32_t e = (uint32_t)(r * (ENC_D / 4) + lane);
      const uint32_t h = rnd32(seed, e, 0x5bd1e995u);
      // four keep decisions from four byte-rotations of one hash would correlate: draw four hashes
      o.x = (h >= drop_thresh) ? o.x * inv_keep : 0.f;
      o.y = (rnd32(seed, e, 1u) >= drop_thresh) ? o.y * inv_keep : 0.f;
      o.z = (rnd32(seed, e, 2u) >= drop_thresh) ? o.z * inv_keep : 0.f;
      o.w = (rnd32(seed, e, 3u) >= drop_thresh) ? o.w * inv_keep : 0.f;
    }
    st4<DTO>(y, r * ENC_D + 4 * lane, o);
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
  }
}

// dx[r] = LN backward of dy[r] w.r.t. the row it normalised, x[index ? index[r] : r] (PACKED like dy: an index may
// name the same source row several times -- two dropout views -- so the caller scatter-adds); per-CTA partial dw / db
template <int DTI, int DTO>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x,
                                                     const int64_t* __restrict__ index, int64_t n_rows,
                                                     const float* __restrict__ w, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, uint32_t drop_thresh,
                                                     float inv_keep, uint64_t seed, void* __restrict__ dx,
                                                     const float* __restrict__ res /*nullable: added to dx*/,
                                                     float* __restrict__ part /*[grid][2][128]*/) {
  __shared__ float4 red[2][8][32];
  seed = epoch_seed(seed);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float4 w4 = ldg_f4(w + 4 * lane);
  float4 dw = make_float4(0.f, 0.f, 0.f, 0.f), db = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = warp; r < n_rows; r += nw) {
    const int64_t src = index ? __ldg(index + r) : r;
    const float4 v = ld4<DTI>(x, src * ENC_D + 4 * lane);
    float4 g = ld4<DTO>(dy, r * ENC_D + 4 * lane);
    if (drop_thresh) {
      const uint32_t e = (uint32_t)(r * (ENC_D / 4) + lane);
      g.x = (rnd32(seed, e, 0x5bd1e995u) >= drop_thresh) ? g.x * inv_keep : 0.f;
      g.y = (rnd32(seed, e, 1u) >= drop_thresh) ? g.y * inv_keep : 0.f;
      g.z = (rnd32(seed, e, 2u) >= drop_thresh) ? g.z * inv_keep : 0.f;
      g.w = (rnd32(seed, e, 3u) >= drop_thresh) ? g.w * inv_keep : 0.f;
    }
    const float mu = mean[r], rs = rstd[r];
    const float4 xh = make_float4((v.x - mu) * rs, (v.y - mu) * rs, (v.z - mu) * rs, (v.w - mu) * rs);
    dw.x += g.x * xh.x; dw.y += g.y * xh.y; dw.z += g.z * xh.z; dw.w += g.w * xh.w;
    db.x += g.x; db.y += g.y; db.z += g.z; db.w += g.w;
    const float4 gw = make_float4(g.x * w4.x, g.y * w4.y, g.z * w4.z, g.w * w4.w);
    const float m1 = warp_sum(gw.x + gw.y + gw.z + gw.w) * (1.f / ENC_D);
    const float m2 = warp_sum(gw.x * xh.x + gw.y * xh.y + gw.z * xh.z + gw.w * xh.w) * (1.f / ENC_D);
    float4 o = make_float4(rs * (gw.x - m1 - xh.x * m2), rs * (gw.y - m1 - xh.y * m2), rs * (gw.z - m1 - xh.z * m2),
                           rs * (gw.w - m1 - xh.w * m2));
    if (res) {                                      // the gradient arriving through the residual connection around the LN
      const float4 rg = ldg_f4(res + r * ENC_D + 4 * lane);
      o.x += rg.x; o.y += rg.y; o.z += rg.z; o.w += rg.w;
    }
    st4<DTI>(dx, r * ENC_D + 4 * lane, o);
  }
  red[0][wib][lane] = dw;
  red[1][wib][lane] = db;
  __syncthreads();
  if (wib < 2) {                                   // warp 0 folds dw, warp 1 folds db, in a fixed order
    float4 s = red[wib][0][lane];
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      const float4 t = red[wib][k][lane];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    *reinterpret_cast<float4*>(part + ((int64_t)blockIdx.x * 2 + wib) * ENC_D + 4 * lane) = s;
  }
}

// out[c] = sum_b part[b][c] over n_blocks partial rows of n_cols_total columns, fixed order (deterministic).
// One CTA of 32 warps per 32 columns; warp w adds partial rows w, w+32, ... (coalesced 128 B reads, 8 independent rows
// in flight), then the 32 warps are folded.  (profiles/r02d_launches.md: with 8 warps per CTA a call was a chain of ~19
// exposed memory latencies, ~20 us cold, 19 calls per step.)
#define PS_WARPS 32
__global__ void __launch_bounds__(PS_WARPS * 32) partial_sum_kernel(const float* __restrict__ part, int n_blocks,
                                                                    int n_cols_total, float* __restrict__ out0,
                                                                    float* __restrict__ out1, int split) {
  __shared__ float red[PS_WARPS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float s[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) s[q] = 0.f;
  if (c < n_cols_total) {
    for (int b = w; b < n_blocks; b += 8 * PS_WARPS) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int bb = b + PS_WARPS * q;
        if (bb < n_blocks) s[q] += part[(int64_t)bb * n_cols_total + c];
      }
    }
  }
  red[w][lane] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  __syncthreads();
  if (w == 0 && c < n_cols_total) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < PS_WARPS; k += 4) { t[0] += red[k][lane]; t[1] += red[k + 1][lane]; t[2] += red[k + 2][lane]; t[3] += red[k + 3][lane]; }
    const float r = (t[0] + t[1]) + (t[2] + t[3]);
    if (c < split) out0[c] = r; else if (out1) out1[c - split] = r;
  }
}

// the same fold for [n_blocks][NV][128] partials -> NV 128-vectors (NV = 3 or 4)
template <int NV>
__global__ void __launch_bounds__(PS_WARPS * 32) partial_sumv_kernel(const float* __restrict__ part, int n_blocks,
                                                                     float* __restrict__ out0, float* __restrict__ out1,
                                                                     float* __restrict__ out2, float* __restrict__ out3) {
  __shared__ float red[PS_WARPS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;                 // < NV * 128
  float s[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) s[q] = 0.f;
  for (int b = w; b < n_blocks; b += 8 * PS_WARPS) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int bb = b + PS_WARPS * q;
      if (bb < n_blocks) s[q] += part[(int64_t)bb * (NV * ENC_D) + c];
    }
  }
  red[w][lane] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  __syncthreads();
  if (w == 0) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < PS_WARPS; k += 4) { t[0] += red[k][lane]; t[1] += red[k + 1][lane]; t[2] += red[k + 2][lane]; t[3] += red[k + 3][lane]; }
    const float r = (t[0] + t[1]) + (t[2] + t[3]);
    float* const outs[4] = {out0, out1, out2, out3};
    outs[c / ENC_D][c % ENC_D] = r;
  }
}

// ------------------------------------------------------------------------------------------------ elementwise
// out = x + dropout(y + bias)     x, out fp32 (or DTX), y DTY, bias fp32 [4*c4n] or NULL
template <int DTX, int DTY>
__global__ void __launch_bounds__(256) dropout_add_fwd_kernel(const void* __restrict__ x, const void* __restrict__ y,
                                                              const float* __restrict__ bias, int c4n, int64_t n4,
                                                              uint32_t drop_thresh, float inv_keep, uint64_t seed,
                                                              void* __restrict__ out) {
  seed = epoch_seed(seed);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = ld4<DTX>(x, 4 * i);
    float4 b = ld4<DTY>(y, 4 * i);
    if (bias) {
      const float4 b4 = ldg_f4(bias + 4 * (int)(i % c4n));
      b.x += b4.x; b.y += b4.y; b.z += b4.z; b.w += b4.w;
    }
    if (drop_thresh) {
      const uint32_t e = (uint32_t)i;
      b.x = (rnd32(seed, e, 0u) >= drop_thresh) ? b.x * inv_keep : 0.f;
      b.y = (rnd32(seed, e, 1u) >= drop_thresh) ? b.y * inv_keep : 0.f;
      b.z = (rnd32(seed, e, 2u) >= drop_thresh) ? b.z * inv_keep : 0.f;
      b.w = (rnd32(seed, e, 3u) >= drop_thresh) ? b.w * inv_keep : 0.f;
    }
    st4<DTX>(out, 4 * i, make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w));
  }
}
// Column sums fused into a grid-stride elementwise kernel over a [rows, 4*c4n] matrix walked as float4 groups: with
// 256 % c4n == 0 the grid stride is a multiple of c4n, so a thread meets ONE column group (threadIdx.x % c4n) and keeps
// its sum in registers; the CTA folds the 256 / c4n threads of a group in a fixed order -> part[blockIdx.x][4*c4n].
__device__ __forceinline__ void cta_colsum(float4 cs, int c4n, float4* cred, float* __restrict__ part) {
  cred[threadIdx.x] = cs;
  __syncthreads();
  if ((int)threadIdx.x < c4n) {
    float4 s = cred[threadIdx.x];
    for (int k = threadIdx.x + c4n; k < 256; k += c4n) {
      const float4 t = cred[k];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    *reinterpret_cast<float4*>(part + ((int64_t)blockIdx.x * c4n + threadIdx.x) * 4) = s;
  }
}

// dy = dropout_mask(g)     g DTX (fp32), dy DTY    (+ per-CTA column sums of dy -> part when given)
template <int DTX, int DTY>
__global__ void __launch_bounds__(256) dropout_bwd_kernel(const void* __restrict__ g, int64_t n4, uint32_t drop_thresh,
                                                          float inv_keep, uint64_t seed, void* __restrict__ dy, int c4n,
                                                          float* __restrict__ part) {
  __shared__ float4 cred[256];
  seed = epoch_seed(seed);
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 b = ld4<DTX>(g, 4 * i);
    if (drop_thresh) {
      const uint32_t e = (uint32_t)i;
      b.x = (rnd32(seed, e, 0u) >= drop_thresh) ? b.x * inv_keep : 0.f;
      b.y = (rnd32(seed, e, 1u) >= drop_thresh) ? b.y * inv_keep : 0.f;
      b.z = (rnd32(seed, e, 2u) >= drop_thresh) ? b.z * inv_keep : 0.f;
      b.w = (rnd32(seed, e, 3u) >= drop_thresh) ? b.w * inv_keep : 0.f;
    }
    st4<DTY>(dy, 4 * i, b);
    cs.x += b.x; cs.y += b.y; cs.z += b.z; cs.w += b.w;
  }
  if (part) cta_colsum(cs, c4n, cred, part);
}

// exact (erf) GELU, as nn.TransformerEncoderLayer(activation="gelu") -> F.gelu(approximate="none")
__device__ __forceinline__ float gelu_f(float z) { return 0.5f * z * (1.f + erff(z * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float z) {
  return 0.5f * (1.f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * __expf(-0.5f * z * z);
}
// 16-bit activations: Phi(z) = 0.5 erfc(-z / sqrt 2) with erfc by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, three
// orders of magnitude below the 2^-9 rounding of the result): one ex2 + one rcp + 6 FMAs instead of libdevice's erff, and
// the backward's exp(-z^2/2) is the same exponential.  The kernels were issue-bound on erff + the dropout hashes
// (profiles/r02d_launches.md: 2.3 TB/s of a 6.5 TB/s stream).  fp32 activations keep erff.
__device__ __forceinline__ void phi_fast(float z, float& Phi, float& e) {
  const float x = fabsf(z) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, x, 1.f));
  e = __expf(-x * x);                                                    // = exp(-z^2 / 2)
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  const float half_erfc = 0.5f * poly * e;                               // 0.5 erfc(|z| / sqrt 2)
  Phi = z >= 0.f ? 1.f - half_erfc : half_erfc;
}
template <int DT> __device__ __forceinline__ float gelu_dt(float z) {
  if constexpr (DT == RS_F32) return gelu_f(z);
  float Phi, e;
  phi_fast(z, Phi, e);
  return z * Phi;
}
template <int DT> __device__ __forceinline__ float gelu_grad_dt(float z) {
  if constexpr (DT == RS_F32) return gelu_grad_f(z);
  float Phi, e;
  phi_fast(z, Phi, e);
  return fmaf(z * 0.3989422804014327f, e, Phi);
}
// out = dropout(gelu(z + bias)) ; BWD: dz = dropout_mask(g) * gelu'(z + bias)  (+ per-CTA column sums of dz -> part,
// the bias gradient, when `part` is given: thread t of every CTA always meets column group t % c4n)
template <int DT, bool BWD>
__global__ void __launch_bounds__(256) gelu_dropout_kernel(const void* __restrict__ z, const void* __restrict__ g,
                                                           const float* __restrict__ bias, int c4n, int64_t n4,
                                                           uint32_t drop_thresh, float inv_keep, uint64_t seed,
                                                           void* __restrict__ out, float* __restrict__ part) {
  __shared__ float4 cred[256];
  seed = epoch_seed(seed);
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = ld4<DT>(z, 4 * i);
    if (bias) {
      const float4 b4 = ldg_f4(bias + 4 * (int)(i % c4n));
      // the reference's Linear rounds z + b to the activation dtype before the GELU
      a = make_float4(round_dt<DT>(a.x + b4.x), round_dt<DT>(a.y + b4.y), round_dt<DT>(a.z + b4.z), round_dt<DT>(a.w + b4.w));
    }
    float4 o;
    if (BWD) {
      const float4 gg = ld4<DT>(g, 4 * i);
      o = make_float4(gg.x * gelu_grad_dt<DT>(a.x), gg.y * gelu_grad_dt<DT>(a.y), gg.z * gelu_grad_dt<DT>(a.z),
                      gg.w * gelu_grad_dt<DT>(a.w));
    } else {
      // the reference's gelu output is rounded to the activation dtype BEFORE its dropout scales it: do the same
      o = make_float4(round_dt<DT>(gelu_dt<DT>(a.x)), round_dt<DT>(gelu_dt<DT>(a.y)), round_dt<DT>(gelu_dt<DT>(a.z)),
                      round_dt<DT>(gelu_dt<DT>(a.w)));
    }
    if (drop_thresh) {
      // two 32-bit hashes for four keep decisions (16 bits each: p is resolved to 2^-16, |p_eff - p| < 1.6e-5): the kernel
      // is issue-bound and the hashes were a sixth of its instructions
      const uint32_t e = (uint32_t)i, t16 = drop_thresh >> 16;
      const uint32_t h0 = rnd32(seed, e, 0u), h1 = rnd32(seed, e, 1u);
      o.x = ((h0 & 0xffffu) >= t16) ? o.x * inv_keep : 0.f;
      o.y = ((h0 >> 16) >= t16) ? o.y * inv_keep : 0.f;
      o.z = ((h1 & 0xffffu) >= t16) ? o.z * inv_keep : 0.f;
      o.w = ((h1 >> 16) >= t16) ? o.w * inv_keep : 0.f;
    }
    st4<DT>(out, 4 * i, o);
    if (BWD) { cs.x += o.x; cs.y += o.y; cs.z += o.z; cs.w += o.w; }
  }
  if (BWD && part) cta_colsum(cs, c4n, cred, part);
}

// ------------------------------------------------------------------------------------------------ LN + activation
// y[r] = dropout(act(LN(x[r] + add[add_rows[r]] + lin_bias))), act = exact GELU (ACT == 1) or identity.  The late-fusion
// head (v1_refine_usertower.py:394-399, :499-510) is Linear(256 -> 128) on cat([sequence row, profile row of its user]) -> LayerNorm
// -> GELU: the Linear splits into a GEMM on the sequence rows (x) and a small GEMM on the distinct profile rows (add, fp32,
// one row per user), the concat never exists, and LN / GELU / the cast for the next Linear are one pass over the rows.
// One warp per row, lane owns 4 consecutive features.
template <int DTI, int DTO, int ACT>
__global__ void __launch_bounds__(256) ln_act_fwd_kernel(const void* __restrict__ x, const float* __restrict__ add,
                                                         const int64_t* __restrict__ add_rows,
                                                         const float* __restrict__ lin_bias, int64_t n_rows,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         float eps, uint32_t drop_thresh, float inv_keep, uint64_t seed,
                                                         void* __restrict__ y, float* __restrict__ mean,
                                                         float* __restrict__ rstd) {
  seed = epoch_seed(seed);
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float4 w4 = ldg_f4(w + 4 * lane), b4 = ldg_f4(bias + 4 * lane);
  const float4 lb = lin_bias ? ldg_f4(lin_bias + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = warp; r < n_rows; r += nw) {
    float4 v = ld4<DTI>(x, r * ENC_D + 4 * lane);
    if (add) {
      const int64_t u = add_rows ? __ldg(add_rows + r) : r;
      const float4 a = ldg_f4(add + u * ENC_D + 4 * lane);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    v.x += lb.x; v.y += lb.y; v.z += lb.z; v.w += lb.w;
    const float mu = warp_sum(v.x + v.y + v.z + v.w) * (1.f / ENC_D);
    const float4 c = make_float4(v.x - mu, v.y - mu, v.z - mu, v.w - mu);
    const float var = warp_sum(c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w) * (1.f / ENC_D);
    const float rs = rsqrtf(var + eps);
    float4 o = make_float4(c.x * rs * w4.x + b4.x, c.y * rs * w4.y + b4.y, c.z * rs * w4.z + b4.z, c.w * rs * w4.w + b4.w);
    if (ACT == 1) o = make_float4(gelu_dt<DTO>(o.x), gelu_dt<DTO>(o.y), gelu_dt<DTO>(o.z), gelu_dt<DTO>(o.w));
    if (drop_thresh) {
      const uint32_t e = (uint32_t)(r * (ENC_D / 4) + lane);
      o.x = (rnd32(seed, e, 0x5bd1e995u) >= drop_thresh) ? o.x * inv_keep : 0.f;
      o.y = (rnd32(seed, e, 1u) >= drop_thresh) ? o.y * inv_keep : 0.f;
      o.z = (rnd32(seed, e, 2u) >= drop_thresh) ? o.z * inv_keep : 0.f;
      o.w = (rnd32(seed, e, 3u) >= drop_thresh) ? o.w * inv_keep : 0.f;
    }
    st4<DTO>(y, r * ENC_D + 4 * lane, o);
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
  }
}

// dx[r] = gradient w.r.t. the LN input of row r (== gradient of x[r]; the caller sums it over the rows that share a
// profile row to get d add, and over all rows to get d lin_bias); per-CTA partial d gamma / d beta
template <int DTI, int DTG, int ACT>
__global__ void __launch_bounds__(256) ln_act_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x,
                                                         const float* __restrict__ add,
                                                         const int64_t* __restrict__ add_rows,
                                                         const float* __restrict__ lin_bias, int64_t n_rows,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         uint32_t drop_thresh, float inv_keep, uint64_t seed,
                                                         void* __restrict__ dx, float* __restrict__ part /*[grid][2][128]*/) {
  __shared__ float4 red[2][8][32];
  seed = epoch_seed(seed);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float4 w4 = ldg_f4(w + 4 * lane), b4 = ldg_f4(bias + 4 * lane);
  const float4 lb = lin_bias ? ldg_f4(lin_bias + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 dw = make_float4(0.f, 0.f, 0.f, 0.f), db = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = warp; r < n_rows; r += nw) {
    float4 v = ld4<DTI>(x, r * ENC_D + 4 * lane);
    if (add) {
      const int64_t u = add_rows ? __ldg(add_rows + r) : r;
      const float4 a = ldg_f4(add + u * ENC_D + 4 * lane);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    v.x += lb.x; v.y += lb.y; v.z += lb.z; v.w += lb.w;
    float4 g = ld4<DTG>(dy, r * ENC_D + 4 * lane);
    if (drop_thresh) {
      const uint32_t e = (uint32_t)(r * (ENC_D / 4) + lane);
      g.x = (rnd32(seed, e, 0x5bd1e995u) >= drop_thresh) ? g.x * inv_keep : 0.f;
      g.y = (rnd32(seed, e, 1u) >= drop_thresh) ? g.y * inv_keep : 0.f;
      g.z = (rnd32(seed, e, 2u) >= drop_thresh) ? g.z * inv_keep : 0.f;
      g.w = (rnd32(seed, e, 3u) >= drop_thresh) ? g.w * inv_keep : 0.f;
    }
    const float mu = mean[r], rs = rstd[r];
    const float4 xh = make_float4((v.x - mu) * rs, (v.y - mu) * rs, (v.z - mu) * rs, (v.w - mu) * rs);
    if (ACT == 1) {
      g.x *= gelu_grad_dt<DTG>(xh.x * w4.x + b4.x); g.y *= gelu_grad_dt<DTG>(xh.y * w4.y + b4.y);
      g.z *= gelu_grad_dt<DTG>(xh.z * w4.z + b4.z); g.w *= gelu_grad_dt<DTG>(xh.w * w4.w + b4.w);
    }
    dw.x += g.x * xh.x; dw.y += g.y * xh.y; dw.z += g.z * xh.z; dw.w += g.w * xh.w;
    db.x += g.x; db.y += g.y; db.z += g.z; db.w += g.w;
    const float4 gw = make_float4(g.x * w4.x, g.y * w4.y, g.z * w4.z, g.w * w4.w);
    const float m1 = warp_sum(gw.x + gw.y + gw.z + gw.w) * (1.f / ENC_D);
    const float m2 = warp_sum(gw.x * xh.x + gw.y * xh.y + gw.z * xh.z + gw.w * xh.w) * (1.f / ENC_D);
    st4<DTI>(dx, r * ENC_D + 4 * lane,
             make_float4(rs * (gw.x - m1 - xh.x * m2), rs * (gw.y - m1 - xh.y * m2), rs * (gw.z - m1 - xh.z * m2),
                         rs * (gw.w - m1 - xh.w * m2)));
  }
  red[0][wib][lane] = dw;
  red[1][wib][lane] = db;
  __syncthreads();
  if (wib < 2) {                                   // warp 0 folds dw, warp 1 folds db, in a fixed order
    float4 s = red[wib][0][lane];
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      const float4 t = red[wib][k][lane];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    *reinterpret_cast<float4*>(part + ((int64_t)blockIdx.x * 2 + wib) * ENC_D + 4 * lane) = s;
  }
}

// ------------------------------------------------------------------------------------------------ residual add + LN
// A pre-norm block ends with  x1 = x + dropout(y + bias)  and the next block starts with  h = LayerNorm(x1): one pass
// instead of two (x1 is written once and not read back), and one backward pass instead of two:
//   dx1 = res + LN'(dh)   (res: the gradient reaching x1 through the residual path)      -> dx   (fp32, to x)
//   dy  = mask(dx1) / keep                                                                 -> dyy  (y's dtype)
// with per-CTA partials of d gamma, d beta and d bias (= column sums of dy).  One warp per row, lane owns 4 features.
template <int DTY, int DTO>
__global__ void __launch_bounds__(256) dropout_add_ln_fwd_kernel(const float* __restrict__ x, const void* __restrict__ y,
                                                                 const float* __restrict__ lin_bias, int64_t n_rows,
                                                                 uint32_t drop_thresh, float inv_keep, uint64_t seed,
                                                                 const float* __restrict__ w, const float* __restrict__ bias,
                                                                 float eps, float* __restrict__ x1, void* __restrict__ h,
                                                                 float* __restrict__ mean, float* __restrict__ rstd) {
  seed = epoch_seed(seed);
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float4 w4 = ldg_f4(w + 4 * lane), b4 = ldg_f4(bias + 4 * lane);
  const float4 lb = lin_bias ? ldg_f4(lin_bias + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = warp; r < n_rows; r += nw) {
    const float4 a = ldg_f4(x + r * ENC_D + 4 * lane);
    float4 t = ld4<DTY>(y, r * ENC_D + 4 * lane);
    t.x += lb.x; t.y += lb.y; t.z += lb.z; t.w += lb.w;
    if (drop_thresh) {
      const uint32_t e = (uint32_t)(r * (ENC_D / 4) + lane);
      t.x = (rnd32(seed, e, 0u) >= drop_thresh) ? t.x * inv_keep : 0.f;
      t.y = (rnd32(seed, e, 1u) >= drop_thresh) ? t.y * inv_keep : 0.f;
      t.z = (rnd32(seed, e, 2u) >= drop_thresh) ? t.z * inv_keep : 0.f;
      t.w = (rnd32(seed, e, 3u) >= drop_thresh) ? t.w * inv_keep : 0.f;
    }
    const float4 v = make_float4(a.x + t.x, a.y + t.y, a.z + t.z, a.w + t.w);
    *reinterpret_cast<float4*>(x1 + r * ENC_D + 4 * lane) = v;
    const float mu = warp_sum(v.x + v.y + v.z + v.w) * (1.f / ENC_D);
    const float4 c = make_float4(v.x - mu, v.y - mu, v.z - mu, v.w - mu);
    const float var = warp_sum(c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w) * (1.f / ENC_D);
    const float rs = rsqrtf(var + eps);
    st4<DTO>(h, r * ENC_D + 4 * lane,
             make_float4(c.x * rs * w4.x + b4.x, c.y * rs * w4.y + b4.y, c.z * rs * w4.z + b4.z, c.w * rs * w4.w + b4.w));
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
  }
}

template <int DTY, int DTO>
__global__ void __launch_bounds__(256) ln_bwd_dropout_kernel(const void* __restrict__ dh, const float* __restrict__ x1,
                                                             const float* __restrict__ res, int64_t n_rows,
                                                             const float* __restrict__ w, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, uint32_t drop_thresh,
                                                             float inv_keep, uint64_t seed, float* __restrict__ dx,
                                                             void* __restrict__ dyy, float* __restrict__ part /*[grid][3][128]*/) {
  __shared__ float4 red[3][8][32];
  seed = epoch_seed(seed);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float4 w4 = ldg_f4(w + 4 * lane);
  float4 dw = make_float4(0.f, 0.f, 0.f, 0.f), db = dw, dl = dw;
  for (int64_t r = warp; r < n_rows; r += nw) {
    const float4 v = ldg_f4(x1 + r * ENC_D + 4 * lane);
    const float4 g = ld4<DTO>(dh, r * ENC_D + 4 * lane);
    const float mu = mean[r], rs = rstd[r];
    const float4 xh = make_float4((v.x - mu) * rs, (v.y - mu) * rs, (v.z - mu) * rs, (v.w - mu) * rs);
    dw.x += g.x * xh.x; dw.y += g.y * xh.y; dw.z += g.z * xh.z; dw.w += g.w * xh.w;
    db.x += g.x; db.y += g.y; db.z += g.z; db.w += g.w;
    const float4 gw = make_float4(g.x * w4.x, g.y * w4.y, g.z * w4.z, g.w * w4.w);
    const float m1 = warp_sum(gw.x + gw.y + gw.z + gw.w) * (1.f / ENC_D);
    const float m2 = warp_sum(gw.x * xh.x + gw.y * xh.y + gw.z * xh.z + gw.w * xh.w) * (1.f / ENC_D);
    float4 o = make_float4(rs * (gw.x - m1 - xh.x * m2), rs * (gw.y - m1 - xh.y * m2), rs * (gw.z - m1 - xh.z * m2),
                           rs * (gw.w - m1 - xh.w * m2));
    if (res) {
      const float4 rg = ldg_f4(res + r * ENC_D + 4 * lane);
      o.x += rg.x; o.y += rg.y; o.z += rg.z; o.w += rg.w;
    }
    *reinterpret_cast<float4*>(dx + r * ENC_D + 4 * lane) = o;
    if (drop_thresh) {
      const uint32_t e = (uint32_t)(r * (ENC_D / 4) + lane);
      o.x = (rnd32(seed, e, 0u) >= drop_thresh) ? o.x * inv_keep : 0.f;
      o.y = (rnd32(seed, e, 1u) >= drop_thresh) ? o.y * inv_keep : 0.f;
      o.z = (rnd32(seed, e, 2u) >= drop_thresh) ? o.z * inv_keep : 0.f;
      o.w = (rnd32(seed, e, 3u) >= drop_thresh) ? o.w * inv_keep : 0.f;
    }
    st4<DTY>(dyy, r * ENC_D + 4 * lane, o);
    dl.x += o.x; dl.y += o.y; dl.z += o.z; dl.w += o.w;
  }
  red[0][wib][lane] = dw;
  red[1][wib][lane] = db;
  red[2][wib][lane] = dl;
  __syncthreads();
  if (wib < 3) {                                   // warps 0 / 1 / 2 fold d gamma / d beta / d bias in a fixed order
    float4 s = red[wib][0][lane];
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      const float4 t = red[wib][k][lane];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    *reinterpret_cast<float4*>(part + ((int64_t)blockIdx.x * 3 + wib) * ENC_D + 4 * lane) = s;
  }
}

// ------------------------------------------------------------------------------------------------ embedding LN + first LN
// The encoder's input is x0 = dropout(LayerNorm_emb(e[index])) and its first layer starts with h = LayerNorm_1(x0)
// (v1_refine_usertower.py:458-459 + the pre-norm layer): one pass over the packed rows instead of two.  In the backward
// every source row u is read by exactly two packed rows (inv1[u], inv2[u]: the two dropout views), and LayerNorm_emb's
// backward is linear in its incoming gradient: the two views' gradients are masked, added and pushed through it ONCE per
// source row -- half the work of the per-packed-row backward and no separate fold pass.
template <int DTI, int DTO>
__global__ void __launch_bounds__(256) emb_ln2_fwd_kernel(const void* __restrict__ x, const int64_t* __restrict__ index,
                                                          int64_t n_rows, const float* __restrict__ w0,
                                                          const float* __restrict__ b0, float eps0, uint32_t drop_thresh,
                                                          float inv_keep, uint64_t seed, const float* __restrict__ w1,
                                                          const float* __restrict__ b1, float eps1, float* __restrict__ x0,
                                                          void* __restrict__ h, float* __restrict__ mean0,
                                                          float* __restrict__ rstd0, float* __restrict__ mean1,
                                                          float* __restrict__ rstd1) {
  seed = epoch_seed(seed);
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float4 wa = ldg_f4(w0 + 4 * lane), ba = ldg_f4(b0 + 4 * lane);
  const float4 wb = ldg_f4(w1 + 4 * lane), bb = ldg_f4(b1 + 4 * lane);
  for (int64_t r = warp; r < n_rows; r += nw) {
    const int64_t src = __ldg(index + r);
    const float4 v = ld4<DTI>(x, src * ENC_D + 4 * lane);
    const float mu = warp_sum(v.x + v.y + v.z + v.w) * (1.f / ENC_D);
    const float4 c = make_float4(v.x - mu, v.y - mu, v.z - mu, v.w - mu);
    const float rs = rsqrtf(warp_sum(c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w) * (1.f / ENC_D) + eps0);
    float4 o = make_float4(c.x * rs * wa.x + ba.x, c.y * rs * wa.y + ba.y, c.z * rs * wa.z + ba.z, c.w * rs * wa.w + ba.w);
    if (drop_thresh) {
      const uint32_t e = (uint32_t)(r * (ENC_D / 4) + lane);
      o.x = (rnd32(seed, e, 0x5bd1e995u) >= drop_thresh) ? o.x * inv_keep : 0.f;
      o.y = (rnd32(seed, e, 1u) >= drop_thresh) ? o.y * inv_keep : 0.f;
      o.z = (rnd32(seed, e, 2u) >= drop_thresh) ? o.z * inv_keep : 0.f;
      o.w = (rnd32(seed, e, 3u) >= drop_thresh) ? o.w * inv_keep : 0.f;
    }
    *reinterpret_cast<float4*>(x0 + r * ENC_D + 4 * lane) = o;
    const float mu1 = warp_sum(o.x + o.y + o.z + o.w) * (1.f / ENC_D);
    const float4 d = make_float4(o.x - mu1, o.y - mu1, o.z - mu1, o.w - mu1);
    const float rs1 = rsqrtf(warp_sum(d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w) * (1.f / ENC_D) + eps1);
    st4<DTO>(h, r * ENC_D + 4 * lane,
             make_float4(d.x * rs1 * wb.x + bb.x, d.y * rs1 * wb.y + bb.y, d.z * rs1 * wb.z + bb.z, d.w * rs1 * wb.w + bb.w));
    if (lane == 0) { mean0[r] = mu; rstd0[r] = rs; mean1[r] = mu1; rstd1[r] = rs1; }
  }
}

// one warp per SOURCE row u: the gradients of its two packed copies -> d e[u]; partials [grid][4][128] = d gamma_1,
// d beta_1, d gamma_emb, d beta_emb
template <int DTI, int DTO>
__global__ void __launch_bounds__(256) emb_ln2_bwd_kernel(const void* __restrict__ dh, const void* __restrict__ x,
                                                          const float* __restrict__ x0, const float* __restrict__ res,
                                                          const int64_t* __restrict__ inv1, const int64_t* __restrict__ inv2,
                                                          int64_t n_src, const float* __restrict__ w0,
                                                          const float* __restrict__ w1, const float* __restrict__ mean0,
                                                          const float* __restrict__ rstd0, const float* __restrict__ mean1,
                                                          const float* __restrict__ rstd1, uint32_t drop_thresh,
                                                          float inv_keep, uint64_t seed, void* __restrict__ dx,
                                                          float* __restrict__ part) {
  __shared__ float4 red[4][8][32];
  seed = epoch_seed(seed);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float4 wa = ldg_f4(w0 + 4 * lane), wb = ldg_f4(w1 + 4 * lane);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 dw1 = z4, db1 = z4, dw0 = z4, db0 = z4;
  for (int64_t u = warp; u < n_src; u += nw) {
    const int64_t rr[2] = {__ldg(inv1 + u), __ldg(inv2 + u)};
    const float4 ev = ld4<DTI>(x, u * ENC_D + 4 * lane);
    float4 acc = z4;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int64_t r = rr[q];
      const float4 v = ldg_f4(x0 + r * ENC_D + 4 * lane);
      const float4 g = ld4<DTO>(dh, r * ENC_D + 4 * lane);
      const float mu = mean1[r], rs = rstd1[r];
      const float4 xh = make_float4((v.x - mu) * rs, (v.y - mu) * rs, (v.z - mu) * rs, (v.w - mu) * rs);
      dw1.x += g.x * xh.x; dw1.y += g.y * xh.y; dw1.z += g.z * xh.z; dw1.w += g.w * xh.w;
      db1.x += g.x; db1.y += g.y; db1.z += g.z; db1.w += g.w;
      const float4 gw = make_float4(g.x * wb.x, g.y * wb.y, g.z * wb.z, g.w * wb.w);
      const float m1 = warp_sum(gw.x + gw.y + gw.z + gw.w) * (1.f / ENC_D);
      const float m2 = warp_sum(gw.x * xh.x + gw.y * xh.y + gw.z * xh.z + gw.w * xh.w) * (1.f / ENC_D);
      float4 o = make_float4(rs * (gw.x - m1 - xh.x * m2), rs * (gw.y - m1 - xh.y * m2), rs * (gw.z - m1 - xh.z * m2),
                             rs * (gw.w - m1 - xh.w * m2));
      if (res) {
        const float4 rg = ldg_f4(res + r * ENC_D + 4 * lane);
        o.x += rg.x; o.y += rg.y; o.z += rg.z; o.w += rg.w;
      }
      if (drop_thresh) {
        const uint32_t e = (uint32_t)(r * (ENC_D / 4) + lane);
        o.x = (rnd32(seed, e, 0x5bd1e995u) >= drop_thresh) ? o.x * inv_keep : 0.f;
        o.y = (rnd32(seed, e, 1u) >= drop_thresh) ? o.y * inv_keep : 0.f;
        o.z = (rnd32(seed, e, 2u) >= drop_thresh) ? o.z * inv_keep : 0.f;
        o.w = (rnd32(seed, e, 3u) >= drop_thresh) ? o.w * inv_keep : 0.f;
      }
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    const float mu = mean0[rr[0]], rs = rstd0[rr[0]];
    const float4 xh = make_float4((ev.x - mu) * rs, (ev.y - mu) * rs, (ev.z - mu) * rs, (ev.w - mu) * rs);
    dw0.x += acc.x * xh.x; dw0.y += acc.y * xh.y; dw0.z += acc.z * xh.z; dw0.w += acc.w * xh.w;
    db0.x += acc.x; db0.y += acc.y; db0.z += acc.z; db0.w += acc.w;
    const float4 gw = make_float4(acc.x * wa.x, acc.y * wa.y, acc.z * wa.z, acc.w * wa.w);
    const float m1 = warp_sum(gw.x + gw.y + gw.z + gw.w) * (1.f / ENC_D);
    const float m2 = warp_sum(gw.x * xh.x + gw.y * xh.y + gw.z * xh.z + gw.w * xh.w) * (1.f / ENC_D);
    st4<DTI>(dx, u * ENC_D + 4 * lane,
             make_float4(rs * (gw.x - m1 - xh.x * m2), rs * (gw.y - m1 - xh.y * m2), rs * (gw.z - m1 - xh.z * m2),
                         rs * (gw.w - m1 - xh.w * m2)));
  }
  red[0][wib][lane] = dw1; red[1][wib][lane] = db1; red[2][wib][lane] = dw0; red[3][wib][lane] = db0;
  __syncthreads();
  if (wib < 4) {
    float4 s = red[wib][0][lane];
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      const float4 t = red[wib][k][lane];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    *reinterpret_cast<float4*>(part + ((int64_t)blockIdx.x * 4 + wib) * ENC_D + 4 * lane) = s;
  }
}

static inline void drop_consts(float p, uint32_t& thresh, float& inv_keep) {
  if (p <= 0.f) { thresh = 0u; inv_keep = 1.f; return; }
  double t = (double)p * 4294967296.0;
  if (t < 1.0) t = 1.0;
  if (t > 4294967295.0) t = 4294967295.0;
  thresh = (uint32_t)t;
  inv_keep = 1.f / (1.f - p);
}

// ------------------------------------------------------------------------------------------------ L2 normalise (128)
// y = x / max(||x||_2, eps)  (F.normalize(p=2, dim=-1), v1_refine_usertower.py:510 and the loop's :807): one warp per
// row, one pass.  inv[r] = 1 / max(||x||, eps), NEGATED when the clamp was active (the backward then has no projection).
template <int DTI, int DTO>
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const void* __restrict__ x, int64_t n_rows, float eps,
                                                         void* __restrict__ y, float* __restrict__ inv) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp; r < n_rows; r += nw) {
    const float4 v = ld4<DTI>(x, r * ENC_D + 4 * lane);
    const float nrm = sqrtf(warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w));
    const float iv = 1.f / fmaxf(nrm, eps);
    st4<DTO>(y, r * ENC_D + 4 * lane, make_float4(v.x * iv, v.y * iv, v.z * iv, v.w * iv));
    if (lane == 0) inv[r] = nrm > eps ? iv : -iv;
  }
}
// dx = (g - y <g, y>) * inv   (clamped rows: dx = g * |inv|)
template <int DTG, int DTY, int DTX>
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const void* __restrict__ g, const void* __restrict__ y,
                                                         const float* __restrict__ inv, int64_t n_rows,
                                                         void* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp; r < n_rows; r += nw) {
    const float4 gv = ld4<DTG>(g, r * ENC_D + 4 * lane), yv = ld4<DTY>(y, r * ENC_D + 4 * lane);
    const float iv = __ldg(inv + r);
    float d = warp_sum(gv.x * yv.x + gv.y * yv.y + gv.z * yv.z + gv.w * yv.w);
    if (iv < 0.f) d = 0.f;
    const float a = fabsf(iv);
    st4<DTX>(dx, r * ENC_D + 4 * lane,
             make_float4((gv.x - yv.x * d) * a, (gv.y - yv.y * d) * a, (gv.z - yv.z * d) * a, (gv.w - yv.w * d) * a));
  }
}

}  // namespace rs

using namespace rs;

#define ENC_DISPATCH1(dt, NAME, ...)                                    \
  switch (dt) {                                                         \
    case RS_F32: { constexpr int NAME = RS_F32; __VA_ARGS__; break; }   \
    case RS_F16: { constexpr int NAME = RS_F16; __VA_ARGS__; break; }   \
    case RS_BF16: { constexpr int NAME = RS_BF16; __VA_ARGS__; break; } \
    default: return RS_ERR_BAD_ARG;                                     \
  }

extern "C" int rs_rng_advance(void* stream) {
  rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>();
  RS_LAUNCH_CHECK();
  return RS_OK;
}

static int attn_check(int64_t n_seq, int64_t total, int H, int hd, int max_len, float p, int64_t zero_tail) {
  if (n_seq <= 0 || total < 0 || H <= 0 || max_len <= 0 || zero_tail < 0 || zero_tail > n_seq) return RS_ERR_BAD_ARG;
  if (hd != ENC_HD || max_len > 64) return RS_ERR_UNSUPPORTED;
  if (p < 0.f || p >= 1.f) return RS_ERR_BAD_ARG;
  if (total * H >= ((int64_t)1 << 32)) return RS_ERR_UNSUPPORTED;      // dropout stream index is 32 bits
  return RS_OK;
}

extern "C" int rs_attn_varlen_fwd(const void* qkv, int dtype, const float* bias, const int32_t* cu_seqlens, int64_t n_seq,
                                  int64_t total_tokens, int n_heads, int head_dim, int max_len, int64_t zero_tail,
                                  int64_t one_row_from, const int64_t* one_rows, float scale, float dropout_p,
                                  uint64_t seed, void* out, float* lse, void* stream) {
  int rc = attn_check(n_seq, total_tokens, n_heads, head_dim, max_len, dropout_p, zero_tail);
  if (rc != RS_OK) return rc;
  if (one_row_from < 0 || one_row_from > n_seq - zero_tail) one_row_from = n_seq - zero_tail;
  if (total_tokens == 0) return RS_OK;
  if (!qkv || !cu_seqlens || !out || !lse) return RS_ERR_BAD_ARG;
  AttnParams p;
  max_len = (max_len + 3) & ~3;      // keeps every per-warp shared-memory region 16-byte aligned
  p.cu = cu_seqlens; p.bias = bias; p.n_seq = n_seq; p.H = n_heads; p.max_len = max_len; p.scale = scale; p.seed = seed;
  p.zero_from = n_seq - zero_tail;
  p.full_to = one_row_from;
  p.one_row = one_rows;
  p.short_split = 0;
  drop_consts(dropout_p, p.drop_thresh, p.inv_keep);
  const int grid = grid_for_warps((dtype == RS_F32 ? n_seq : (p.full_to > 0 ? p.full_to : 1)) * n_heads, 8, 8);
  cudaStream_t st = (cudaStream_t)stream;
  // 16-bit operands: tensor-core tiles (attn_mma.cuh); fp32: the SIMT kernel
  if (dtype == RS_BF16) attn3_fwd_kernel<RS_BF16><<<grid, 256, 0, st>>>(qkv, p, out, lse);
  else if (dtype == RS_F16) attn3_fwd_kernel<RS_F16><<<grid, 256, 0, st>>>(qkv, p, out, lse);
  else if (dtype == RS_F32) attn2_fwd_kernel<RS_F32><<<grid, 256, 0, st>>>(qkv, p, out, lse);
  else return RS_ERR_BAD_ARG;
  RS_LAUNCH_CHECK();
  if (p.full_to < p.zero_from) {
    const int g2 = grid_for_warps((p.zero_from - p.full_to) * n_heads, 8, 8);
    ENC_DISPATCH1(dtype, DT, (attn_one_fwd_kernel<DT><<<g2, 256, 0, st>>>(qkv, p, out, lse)));
    RS_LAUNCH_CHECK();
  }
  return RS_OK;
}

static const int ATTN_SHORT_SPLIT = [] { const char* e = getenv("RS_ATTN_SHORT_SPLIT"); return (e && e[0] == '0') ? 0 : 1; }();

extern "C" int rs_attn_varlen_bwd(const void* qkv, const void* d_out, const void* out, int dtype, const float* bias,
                                  const float* lse,
                                  const int32_t* cu_seqlens, int64_t n_seq, int64_t total_tokens, int n_heads,
                                  int head_dim, int max_len, int64_t zero_tail, int64_t one_row_from,
                                  const int64_t* one_rows, float scale, float dropout_p, uint64_t seed, void* d_qkv,
                                  void* stream) {
  int rc = attn_check(n_seq, total_tokens, n_heads, head_dim, max_len, dropout_p, zero_tail);
  if (rc != RS_OK) return rc;
  if (one_row_from < 0 || one_row_from > n_seq - zero_tail) one_row_from = n_seq - zero_tail;
  if (total_tokens == 0) return RS_OK;
  if (!qkv || !d_out || !out || !lse || !cu_seqlens || !d_qkv) return RS_ERR_BAD_ARG;
  AttnParams p;
  max_len = (max_len + 3) & ~3;      // keeps every per-warp shared-memory region 16-byte aligned
  p.cu = cu_seqlens; p.bias = bias; p.n_seq = n_seq; p.H = n_heads; p.max_len = max_len; p.scale = scale; p.seed = seed;
  p.zero_from = n_seq - zero_tail;
  p.full_to = one_row_from;
  p.one_row = one_rows;
  // one-tile sequences go to their own kernel (16-bit operands, bias already inside qkv)
  p.short_split = (dtype != RS_F32 && bias == nullptr && ATTN_SHORT_SPLIT) ? 1 : 0;
  drop_consts(dropout_p, p.drop_thresh, p.inv_keep);
  const int grid = grid_for_warps((dtype == RS_F32 ? n_seq : (p.full_to > 0 ? p.full_to : 1)) * n_heads, 8, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (p.short_split) {
    // once-per-pair kernels: <= 16 tokens (one tile, registers only) and 17..64 tokens (dQ summed in shared memory); the
    // two-phase kernel below then only clears the zero tail
    const int smem_long = 8 * AT3L_WARP_BYTES;
    if (dtype == RS_BF16) {
      attn3_bwd_short_kernel<RS_BF16><<<grid, 256, 0, st>>>(qkv, d_out, out, lse, p, d_qkv);
      cudaError_t e = cudaFuncSetAttribute(attn3_bwd_long_kernel<RS_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_long);
      if (e != cudaSuccess) return (int)e;
      attn3_bwd_long_kernel<RS_BF16><<<grid, 256, smem_long, st>>>(qkv, d_out, out, lse, p, d_qkv);
    } else {
      attn3_bwd_short_kernel<RS_F16><<<grid, 256, 0, st>>>(qkv, d_out, out, lse, p, d_qkv);
      cudaError_t e = cudaFuncSetAttribute(attn3_bwd_long_kernel<RS_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_long);
      if (e != cudaSuccess) return (int)e;
      attn3_bwd_long_kernel<RS_F16><<<grid, 256, smem_long, st>>>(qkv, d_out, out, lse, p, d_qkv);
    }
    RS_LAUNCH_CHECK_N(2);
  }
  if (dtype == RS_BF16) attn3_bwd_kernel<RS_BF16><<<grid, 256, 0, st>>>(qkv, d_out, out, lse, p, d_qkv);
  else if (dtype == RS_F16) attn3_bwd_kernel<RS_F16><<<grid, 256, 0, st>>>(qkv, d_out, out, lse, p, d_qkv);
  else if (dtype == RS_F32) attn2_bwd_kernel<RS_F32><<<grid, 256, 0, st>>>(qkv, d_out, out, lse, p, d_qkv);
  else return RS_ERR_BAD_ARG;
  RS_LAUNCH_CHECK();
  if (p.full_to < p.zero_from) {
    const int g2 = grid_for_warps((p.zero_from - p.full_to) * n_heads, 8, 8);
    ENC_DISPATCH1(dtype, DT, (attn_one_bwd_kernel<DT><<<g2, 256, 0, st>>>(qkv, d_out, out, lse, p, d_qkv)));
    RS_LAUNCH_CHECK();
  }
  return RS_OK;
}

static int colsum_grid(int64_t n_rows, int rpi) {
  int64_t g = (n_rows + rpi * 8 - 1) / (rpi * 8);
  const int64_t cap = RS_NUM_SMS * 4;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
extern "C" size_t rs_colsum_workspace_bytes(int64_t n_rows, int64_t n_cols) {
  return (size_t)RS_NUM_SMS * 4 * (size_t)n_cols * sizeof(float);
}
extern "C" int rs_colsum(const void* x, int dtype, int64_t n_rows, int64_t n_cols, float* out, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (!x || !out || !workspace || n_rows < 0 || n_cols <= 0 || (n_cols & 3) || n_cols > 1024) return RS_ERR_BAD_ARG;
  if (workspace_bytes < rs_colsum_workspace_bytes(n_rows, n_cols)) return RS_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int c4n = (int)(n_cols / 4), rpi = 256 / c4n;
  const int grid = colsum_grid(n_rows, rpi);
  const size_t smem = (size_t)rpi * c4n * sizeof(float4);
  float* part = (float*)workspace;
  ENC_DISPATCH1(dtype, DT, (colsum_partial_kernel<DT><<<grid, 256, smem, st>>>(x, n_rows, (int)n_cols, part)));
  RS_LAUNCH_CHECK();
  partial_sum_kernel<<<(int)((n_cols + 31) / 32), PS_WARPS * 32, 0, st>>>(part, grid, (int)n_cols, out, nullptr, (int)n_cols);
  RS_LAUNCH_CHECK();
  return RS_OK;
}

#define LN_GRID_CAP (RS_NUM_SMS * 8)
static int ln_grid(int64_t n_rows) { return grid_for_warps(n_rows, 8, 8); }

extern "C" size_t rs_ln_bwd_workspace_bytes(int64_t n_rows) { return (size_t)ln_grid(n_rows) * 2 * ENC_D * sizeof(float); }

extern "C" int rs_ln_fwd(const void* x, int x_dtype, const int64_t* index, int64_t n_rows, int64_t dim, const float* w,
                         const float* b, float eps, float dropout_p, uint64_t seed, void* y, int y_dtype, float* mean,
                         float* rstd, void* stream) {
  if (n_rows == 0) return RS_OK;
  if (!x || !w || !b || !y || !mean || !rstd || n_rows < 0) return RS_ERR_BAD_ARG;
  if (dim != ENC_D) return RS_ERR_UNSUPPORTED;
  if (n_rows * (ENC_D / 4) >= ((int64_t)1 << 32)) return RS_ERR_UNSUPPORTED;
  uint32_t th; float ik;
  drop_consts(dropout_p, th, ik);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(n_rows);
  ENC_DISPATCH1(x_dtype, DTI, ENC_DISPATCH1(y_dtype, DTO, (ln_fwd_kernel<DTI, DTO><<<grid, 256, 0, st>>>(
      x, index, n_rows, w, b, eps, th, ik, seed, y, mean, rstd))));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_ln_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const int64_t* index, int64_t n_rows,
                         int64_t dim, const float* w, const float* mean, const float* rstd, float dropout_p,
                         uint64_t seed, void* dx, const float* residual_grad, float* dw, float* db, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (n_rows == 0) return RS_OK;
  if (!dy || !x || !w || !mean || !rstd || !dx || !dw || !db || !workspace) return RS_ERR_BAD_ARG;
  if (dim != ENC_D) return RS_ERR_UNSUPPORTED;
  if (workspace_bytes < rs_ln_bwd_workspace_bytes(n_rows)) return RS_ERR_WORKSPACE;
  uint32_t th; float ik;
  drop_consts(dropout_p, th, ik);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(n_rows);
  float* part = (float*)workspace;
  ENC_DISPATCH1(x_dtype, DTI, ENC_DISPATCH1(dy_dtype, DTO, (ln_bwd_kernel<DTI, DTO><<<grid, 256, 0, st>>>(
      dy, x, index, n_rows, w, mean, rstd, th, ik, seed, dx, residual_grad, part))));
  RS_LAUNCH_CHECK();
  partial_sum_kernel<<<(2 * ENC_D + 31) / 32, PS_WARPS * 32, 0, st>>>(part, grid, 2 * ENC_D, dw, db, ENC_D);
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_ln_act_fwd(const void* x, int x_dtype, const float* add, const int64_t* add_rows,
                             const float* lin_bias, int64_t n_rows, int64_t dim, const float* w, const float* b,
                             float eps, int act, float dropout_p, uint64_t seed, void* y, int y_dtype, float* mean,
                             float* rstd, void* stream) {
  if (n_rows == 0) return RS_OK;
  if (!x || !w || !b || !y || !mean || !rstd || n_rows < 0 || (add_rows && !add)) return RS_ERR_BAD_ARG;
  if (dim != ENC_D || (act != 0 && act != 1)) return RS_ERR_UNSUPPORTED;
  if (n_rows * (ENC_D / 4) >= ((int64_t)1 << 32)) return RS_ERR_UNSUPPORTED;
  uint32_t th; float ik;
  drop_consts(dropout_p, th, ik);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(n_rows);
#define LN_ACT_FWD(ACT)                                                                                          \
  ENC_DISPATCH1(x_dtype, DTI, ENC_DISPATCH1(y_dtype, DTO, (ln_act_fwd_kernel<DTI, DTO, ACT><<<grid, 256, 0, st>>>( \
      x, add, add_rows, lin_bias, n_rows, w, b, eps, th, ik, seed, y, mean, rstd))))
  if (act == 1) { LN_ACT_FWD(1); } else { LN_ACT_FWD(0); }
#undef LN_ACT_FWD
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_ln_act_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* add,
                             const int64_t* add_rows, const float* lin_bias, int64_t n_rows, int64_t dim,
                             const float* w, const float* b, const float* mean, const float* rstd, int act,
                             float dropout_p, uint64_t seed, void* dx, float* dw, float* db, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (n_rows == 0) return RS_OK;
  if (!dy || !x || !w || !b || !mean || !rstd || !dx || !dw || !db || !workspace || (add_rows && !add))
    return RS_ERR_BAD_ARG;
  if (dim != ENC_D || (act != 0 && act != 1)) return RS_ERR_UNSUPPORTED;
  if (workspace_bytes < rs_ln_bwd_workspace_bytes(n_rows)) return RS_ERR_WORKSPACE;
  uint32_t th; float ik;
  drop_consts(dropout_p, th, ik);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(n_rows);
  float* part = (float*)workspace;
#define LN_ACT_BWD(ACT)                                                                                           \
  ENC_DISPATCH1(x_dtype, DTI, ENC_DISPATCH1(dy_dtype, DTG, (ln_act_bwd_kernel<DTI, DTG, ACT><<<grid, 256, 0, st>>>( \
      dy, x, add, add_rows, lin_bias, n_rows, w, b, mean, rstd, th, ik, seed, dx, part))))
  if (act == 1) { LN_ACT_BWD(1); } else { LN_ACT_BWD(0); }
#undef LN_ACT_BWD
  RS_LAUNCH_CHECK();
  partial_sum_kernel<<<(2 * ENC_D + 31) / 32, PS_WARPS * 32, 0, st>>>(part, grid, 2 * ENC_D, dw, db, ENC_D);
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_dropout_add_ln_fwd(const float* x, const void* y, int y_dtype, const float* lin_bias, int64_t n_rows,
                                     int64_t dim, float dropout_p, uint64_t seed, const float* w, const float* b,
                                     float eps, float* x1, void* h, int h_dtype, float* mean, float* rstd,
                                     void* stream) {
  if (n_rows == 0) return RS_OK;
  if (!x || !y || !w || !b || !x1 || !h || !mean || !rstd || n_rows < 0) return RS_ERR_BAD_ARG;
  if (dim != ENC_D) return RS_ERR_UNSUPPORTED;
  if (n_rows * (ENC_D / 4) >= ((int64_t)1 << 32)) return RS_ERR_UNSUPPORTED;
  uint32_t th; float ik;
  drop_consts(dropout_p, th, ik);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(n_rows);
  ENC_DISPATCH1(y_dtype, DTY, ENC_DISPATCH1(h_dtype, DTO, (dropout_add_ln_fwd_kernel<DTY, DTO><<<grid, 256, 0, st>>>(
      x, y, lin_bias, n_rows, th, ik, seed, w, b, eps, x1, h, mean, rstd))));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" size_t rs_ln_bwd_dropout_workspace_bytes(int64_t n_rows) {
  return (size_t)ln_grid(n_rows) * 3 * ENC_D * sizeof(float);
}

extern "C" int rs_ln_bwd_dropout(const void* dh, int dh_dtype, const float* x1, const float* residual_grad,
                                 int64_t n_rows, int64_t dim, const float* w, const float* mean, const float* rstd,
                                 float dropout_p, uint64_t seed, float* dx, void* dy, int dy_dtype, float* dw, float* db,
                                 float* d_lin_bias, void* workspace, size_t workspace_bytes, void* stream) {
  if (n_rows == 0) return RS_OK;
  if (!dh || !x1 || !w || !mean || !rstd || !dx || !dy || !dw || !db || !d_lin_bias || !workspace) return RS_ERR_BAD_ARG;
  if (dim != ENC_D) return RS_ERR_UNSUPPORTED;
  if (workspace_bytes < rs_ln_bwd_dropout_workspace_bytes(n_rows)) return RS_ERR_WORKSPACE;
  uint32_t th; float ik;
  drop_consts(dropout_p, th, ik);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(n_rows);
  float* part = (float*)workspace;
  ENC_DISPATCH1(dy_dtype, DTY, ENC_DISPATCH1(dh_dtype, DTO, (ln_bwd_dropout_kernel<DTY, DTO><<<grid, 256, 0, st>>>(
      dh, x1, residual_grad, n_rows, w, mean, rstd, th, ik, seed, dx, dy, part))));
  RS_LAUNCH_CHECK();
  // part rows are [d gamma | d beta | d bias]: the first two go to (dw, db), the third to d_lin_bias
  partial_sumv_kernel<3><<<(3 * ENC_D + 31) / 32, PS_WARPS * 32, 0, st>>>(part, grid, dw, db, d_lin_bias, nullptr);
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_emb_ln2_fwd(const void* x, int x_dtype, const int64_t* index, int64_t n_rows, int64_t dim, const float* w0,
                              const float* b0, float eps0, float dropout_p, uint64_t seed, const float* w1,
                              const float* b1, float eps1, float* x0, void* h, int h_dtype, float* mean0, float* rstd0,
                              float* mean1, float* rstd1, void* stream) {
  if (n_rows == 0) return RS_OK;
  if (!x || !index || !w0 || !b0 || !w1 || !b1 || !x0 || !h || !mean0 || !rstd0 || !mean1 || !rstd1 || n_rows < 0)
    return RS_ERR_BAD_ARG;
  if (dim != ENC_D) return RS_ERR_UNSUPPORTED;
  if (n_rows * (ENC_D / 4) >= ((int64_t)1 << 32)) return RS_ERR_UNSUPPORTED;
  uint32_t th; float ik;
  drop_consts(dropout_p, th, ik);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(n_rows);
  ENC_DISPATCH1(x_dtype, DTI, ENC_DISPATCH1(h_dtype, DTO, (emb_ln2_fwd_kernel<DTI, DTO><<<grid, 256, 0, st>>>(
      x, index, n_rows, w0, b0, eps0, th, ik, seed, w1, b1, eps1, x0, h, mean0, rstd0, mean1, rstd1))));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" size_t rs_emb_ln2_bwd_workspace_bytes(int64_t n_src) {
  return (size_t)ln_grid(n_src) * 4 * ENC_D * sizeof(float);
}

extern "C" int rs_emb_ln2_bwd(const void* dh, int dh_dtype, const void* x, int x_dtype, const float* x0,
                              const float* residual_grad, const int64_t* inv1, const int64_t* inv2, int64_t n_src,
                              int64_t dim, const float* w0, const float* w1, const float* mean0, const float* rstd0,
                              const float* mean1, const float* rstd1, float dropout_p, uint64_t seed, void* dx, float* dw0,
                              float* db0, float* dw1, float* db1, void* workspace, size_t workspace_bytes, void* stream) {
  if (n_src == 0) return RS_OK;
  if (!dh || !x || !x0 || !inv1 || !inv2 || !w0 || !w1 || !mean0 || !rstd0 || !mean1 || !rstd1 || !dx || !dw0 || !db0 ||
      !dw1 || !db1 || !workspace)
    return RS_ERR_BAD_ARG;
  if (dim != ENC_D) return RS_ERR_UNSUPPORTED;
  if (workspace_bytes < rs_emb_ln2_bwd_workspace_bytes(n_src)) return RS_ERR_WORKSPACE;
  uint32_t th; float ik;
  drop_consts(dropout_p, th, ik);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(n_src);
  float* part = (float*)workspace;
  ENC_DISPATCH1(x_dtype, DTI, ENC_DISPATCH1(dh_dtype, DTO, (emb_ln2_bwd_kernel<DTI, DTO><<<grid, 256, 0, st>>>(
      dh, x, x0, residual_grad, inv1, inv2, n_src, w0, w1, mean0, rstd0, mean1, rstd1, th, ik, seed, dx, part))));
  RS_LAUNCH_CHECK();
  // part rows: [d gamma_1 | d beta_1 | d gamma_emb | d beta_emb]
  partial_sumv_kernel<4><<<(4 * ENC_D + 31) / 32, PS_WARPS * 32, 0, st>>>(part, grid, dw1, db1, dw0, db0);
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_l2_normalize_fwd(const void* x, int x_dtype, int64_t n_rows, int64_t dim, float eps, void* y,
                                   int y_dtype, float* inv_norm, void* stream) {
  if (n_rows == 0) return RS_OK;
  if (!x || !y || !inv_norm || n_rows < 0) return RS_ERR_BAD_ARG;
  if (dim != ENC_D) return RS_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(n_rows);
  ENC_DISPATCH1(x_dtype, DTI, ENC_DISPATCH1(y_dtype, DTO, (l2norm_fwd_kernel<DTI, DTO><<<grid, 256, 0, st>>>(
      x, n_rows, eps, y, inv_norm))));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_l2_normalize_bwd(const void* g, int g_dtype, const void* y, int y_dtype, const float* inv_norm,
                                   int64_t n_rows, int64_t dim, void* dx, int dx_dtype, void* stream) {
  if (n_rows == 0) return RS_OK;
  if (!g || !y || !inv_norm || !dx || n_rows < 0) return RS_ERR_BAD_ARG;
  if (dim != ENC_D) return RS_ERR_UNSUPPORTED;
  if (y_dtype != RS_F32) return RS_ERR_UNSUPPORTED;          // the normalised rows are kept in fp32 (as under autocast)
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(n_rows);
  ENC_DISPATCH1(g_dtype, DTG, ENC_DISPATCH1(dx_dtype, DTX, (l2norm_bwd_kernel<DTG, RS_F32, DTX><<<grid, 256, 0, st>>>(
      g, y, inv_norm, n_rows, dx))));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

static int ew_grid(int64_t n4) {
  int64_t g = (n4 + 255) / 256;
  const int64_t cap = (int64_t)RS_NUM_SMS * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" int rs_dropout_add_fwd(const void* x, int x_dtype, const void* y, int y_dtype, const float* bias,
                                  int64_t n_cols, int64_t n, float dropout_p, uint64_t seed, void* out, void* stream) {
  if (n == 0) return RS_OK;
  if (!x || !y || !out || n < 0 || (n & 3) || n_cols <= 0 || (n_cols & 3) || n % n_cols) return RS_ERR_BAD_ARG;
  if (n / 4 >= ((int64_t)1 << 32)) return RS_ERR_UNSUPPORTED;
  uint32_t th; float ik;
  drop_consts(dropout_p, th, ik);
  cudaStream_t st = (cudaStream_t)stream;
  ENC_DISPATCH1(x_dtype, DTX, ENC_DISPATCH1(y_dtype, DTY, (dropout_add_fwd_kernel<DTX, DTY><<<ew_grid(n / 4), 256, 0, st>>>(
      x, y, bias, (int)(n_cols / 4), n / 4, th, ik, seed, out))));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_dropout_bwd(const void* g, int g_dtype, int64_t n, float dropout_p, uint64_t seed, void* dy,
                              int dy_dtype, void* stream) {
  if (n == 0) return RS_OK;
  if (!g || !dy || n < 0 || (n & 3)) return RS_ERR_BAD_ARG;
  uint32_t th; float ik;
  drop_consts(dropout_p, th, ik);
  cudaStream_t st = (cudaStream_t)stream;
  ENC_DISPATCH1(g_dtype, DTX, ENC_DISPATCH1(dy_dtype, DTY, (dropout_bwd_kernel<DTX, DTY><<<ew_grid(n / 4), 256, 0, st>>>(
      g, n / 4, th, ik, seed, dy, 0, nullptr))));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

// the same two backward kernels with the bias gradient (column sums of their output) folded in: no second pass over dy
static bool cs_cols_ok(int64_t n_cols) { return n_cols > 0 && (n_cols & 3) == 0 && 256 % (n_cols / 4) == 0; }
extern "C" size_t rs_ew_colsum_workspace_bytes(int64_t n, int64_t n_cols) {
  return (size_t)ew_grid(n / 4) * (size_t)n_cols * sizeof(float);
}
extern "C" int rs_dropout_bwd_bias(const void* g, int g_dtype, int64_t n, int64_t n_cols, float dropout_p, uint64_t seed,
                                   void* dy, int dy_dtype, float* d_bias, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  if (!d_bias || n_cols <= 0) return RS_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) { cudaMemsetAsync(d_bias, 0, (size_t)n_cols * sizeof(float), st); return RS_OK; }
  if (!g || !dy || !workspace || n < 0 || (n & 3) || n % n_cols) return RS_ERR_BAD_ARG;
  if (!cs_cols_ok(n_cols)) return RS_ERR_UNSUPPORTED;
  if (workspace_bytes < rs_ew_colsum_workspace_bytes(n, n_cols)) return RS_ERR_WORKSPACE;
  uint32_t th; float ik;
  drop_consts(dropout_p, th, ik);
  const int grid = ew_grid(n / 4);
  float* part = (float*)workspace;
  ENC_DISPATCH1(g_dtype, DTX, ENC_DISPATCH1(dy_dtype, DTY, (dropout_bwd_kernel<DTX, DTY><<<grid, 256, 0, st>>>(
      g, n / 4, th, ik, seed, dy, (int)(n_cols / 4), part))));
  RS_LAUNCH_CHECK();
  partial_sum_kernel<<<(int)((n_cols + 31) / 32), PS_WARPS * 32, 0, st>>>(part, grid, (int)n_cols, d_bias, nullptr, (int)n_cols);
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_gelu_dropout_bwd_bias(const void* z, const void* g, int dtype, const float* bias, int64_t n_cols,
                                        int64_t n, float dropout_p, uint64_t seed, void* dz, float* d_bias,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  if (!d_bias || n_cols <= 0) return RS_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) { cudaMemsetAsync(d_bias, 0, (size_t)n_cols * sizeof(float), st); return RS_OK; }
  if (!z || !g || !dz || !workspace || n < 0 || (n & 3) || n % n_cols) return RS_ERR_BAD_ARG;
  if (!cs_cols_ok(n_cols)) return RS_ERR_UNSUPPORTED;
  if (workspace_bytes < rs_ew_colsum_workspace_bytes(n, n_cols)) return RS_ERR_WORKSPACE;
  uint32_t th; float ik;
  drop_consts(dropout_p, th, ik);
  const int grid = ew_grid(n / 4);
  float* part = (float*)workspace;
  ENC_DISPATCH1(dtype, DT, (gelu_dropout_kernel<DT, true><<<grid, 256, 0, st>>>(z, g, bias, (int)(n_cols / 4), n / 4, th, ik, seed, dz, part)));
  RS_LAUNCH_CHECK();
  partial_sum_kernel<<<(int)((n_cols + 31) / 32), PS_WARPS * 32, 0, st>>>(part, grid, (int)n_cols, d_bias, nullptr, (int)n_cols);
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_gelu_dropout_fwd(const void* z, int dtype, const float* bias, int64_t n_cols, int64_t n,
                                   float dropout_p, uint64_t seed, void* out, void* stream) {
  if (n == 0) return RS_OK;
  if (!z || !out || n < 0 || (n & 3) || n_cols <= 0 || (n_cols & 3) || n % n_cols) return RS_ERR_BAD_ARG;
  if (n / 4 >= ((int64_t)1 << 32)) return RS_ERR_UNSUPPORTED;
  uint32_t th; float ik;
  drop_consts(dropout_p, th, ik);
  cudaStream_t st = (cudaStream_t)stream;
  ENC_DISPATCH1(dtype, DT, (gelu_dropout_kernel<DT, false><<<ew_grid(n / 4), 256, 0, st>>>(z, nullptr, bias, (int)(n_cols / 4), n / 4, th, ik, seed, out, nullptr)));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_gelu_dropout_bwd(const void* z, const void* g, int dtype, const float* bias, int64_t n_cols, int64_t n,
                                   float dropout_p, uint64_t seed, void* dz, void* stream) {
  if (n == 0) return RS_OK;
  if (!z || !g || !dz || n < 0 || (n & 3) || n_cols <= 0 || (n_cols & 3) || n % n_cols) return RS_ERR_BAD_ARG;
  uint32_t th; float ik;
  drop_consts(dropout_p, th, ik);
  cudaStream_t st = (cudaStream_t)stream;
  ENC_DISPATCH1(dtype, DT, (gelu_dropout_kernel<DT, true><<<ew_grid(n / 4), 256, 0, st>>>(z, g, bias, (int)(n_cols / 4), n / 4, th, ik, seed, dz, nullptr)));
  RS_LAUNCH_CHECK();
  return RS_OK;
}
