// Attention v3 for 16-bit operands (included by encoder.cu, inside namespace rs): the same causal short-sequence
// attention, with the two contractions of every 16 x 16 (query, key) tile on the tensor cores.
//
// Why mma.sync and not tcgen05 here: a sequence has 13 tokens on average (<= 64), one (sequence, head) is a 16 x 16
// x 32 problem -- a tcgen05 tile is 128 rows and lives in TMEM behind mbarriers; the warp-level m16n8k16 instruction
// is the tensor-core shape that fits.  The SIMT version (attn2_*) spends ~78 instructions per (query, key) pair on the
// dot products; here a pair costs ~1 instruction of tensor-core work and what remains is the elementwise softmax /
// dropout on the accumulator fragments: 5-6x fewer instructions per item, the kernels are instruction-bound.
//
// One warp per (sequence, head).  Operands come straight from the packed in_proj output [T, 3, H, 32]:
//   * A fragments and "col" B fragments are pairs of adjacent 16-bit elements of one token row -> 32-bit global loads;
//   * B fragments whose k index is a TOKEN index (V in P.V, K in dS.K, Q / dO in the dK-dV phase) are read with
//     ldmatrix.trans from a [16][32] tile staged in shared memory (80-byte rows: conflict-free).
// The in_proj bias (kept out of the GEMM so that its gradient is a column sum of d_qkv) is added on load in the
// operand precision (packed 16-bit add).  Dropout: the same counter-based hash of (row id, key) as everywhere else,
// so forward, dQ phase and dK/dV phase agree on the mask.
#pragma once

#define AT3_LD 40          // tile row pitch in 16-bit elements (80 B)

// rnd32(seed, a, b) == rnd_col(seed_hi, rnd_row(seed, a), b): the row half is hoisted out of the element loops
__device__ __forceinline__ uint32_t rnd_row(uint64_t seed, uint32_t a) { return mix32(a * 0x9E3779B1u + (uint32_t)seed); }
__device__ __forceinline__ uint32_t rnd_col(uint32_t seed_hi, uint32_t hrow, uint32_t b) {
  return mix32(hrow ^ (b * 0x85EBCA77u + seed_hi));
}

template <int DT>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (DT == RS_BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const uint16_t* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <int DT> __device__ __forceinline__ uint32_t pack16(float a, float b) {
  return DT == RS_BF16 ? pack_bf16(a, b) : pack_f16(a, b);
}
template <int DT> __device__ __forceinline__ float2 unpack16(uint32_t u) {
  return DT == RS_BF16 ? unpack_bf16(u) : unpack_f16(u);
}
template <int DT> __device__ __forceinline__ uint32_t add16x2(uint32_t x, uint32_t y) {
  if constexpr (DT == RS_BF16) {
    __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&x), *reinterpret_cast<__nv_bfloat162*>(&y));
    return *reinterpret_cast<uint32_t*>(&r);
  } else {
    __half2 r = __hadd2(*reinterpret_cast<__half2*>(&x), *reinterpret_cast<__half2*>(&y));
    return *reinterpret_cast<uint32_t*>(&r);
  }
}
__device__ __forceinline__ uint32_t ldg_u32(const void* base, int64_t elem_off) {
  return __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(base) + elem_off));
}
// bias pairs of the fragment columns of this lane: [ks][half] <-> columns 16 ks + 2 t + 8 half (+0, +1)
template <int DT>
__device__ __forceinline__ void frag_bias(uint32_t (&b)[2][2], const float* bias, int64_t col0, int t) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int c = 16 * ks + 2 * t + 8 * hf;
      b[ks][hf] = bias ? pack16<DT>(__ldg(bias + col0 + c), __ldg(bias + col0 + c + 1)) : 0u;
    }
}
// A fragments of a [16 rows x 32] operand: rows (row0 + g, row0 + g + 8) of `src`, both k-steps; rows >= limit read 0
template <int DT>
__device__ __forceinline__ void load_a(uint32_t (&a)[2][4], const void* src, int64_t tok0, int row0, int limit,
                                       int64_t row_stride, int64_t col0, const uint32_t (&b)[2][2], bool has_bias, int g,
                                       int t) {
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int r = row0 + g + 8 * rr;
    const bool ok = r < limit;
    const int64_t base = (tok0 + r) * row_stride + col0 + 2 * t;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t v = ok ? ldg_u32(src, base + 16 * ks + 8 * hf) : 0u;
        if (has_bias && ok) v = add16x2<DT>(v, b[ks][hf]);
        a[ks][rr + 2 * hf] = v;
      }
  }
}
// "col" B fragments of a [16 rows (n) x 32 (k)] operand: b[nt][ks][half], n = row0 + 8 nt + g
template <int DT>
__device__ __forceinline__ void load_b(uint32_t (&bf)[2][2][2], const void* src, int64_t tok0, int row0, int limit,
                                       int64_t row_stride, int64_t col0, const uint32_t (&b)[2][2], bool has_bias, int g,
                                       int t) {
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int r = row0 + 8 * nt + g;
    const bool ok = r < limit;
    const int64_t base = (tok0 + r) * row_stride + col0 + 2 * t;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t v = ok ? ldg_u32(src, base + 16 * ks + 8 * hf) : 0u;
        if (has_bias && ok) v = add16x2<DT>(v, b[ks][hf]);
        bf[nt][ks][hf] = v;
      }
  }
}
// C[2 nt][4] = A[16 x 32] . B[16 x 32]^T
template <int DT>
__device__ __forceinline__ void mma_abt(float (&c)[2][4], const uint32_t (&a)[2][4], const uint32_t (&b)[2][2][2]) {
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) c[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) mma16816<DT>(c[nt], a[ks], b[nt][ks][0], b[nt][ks][1]);
  }
}
// stage rows [row0, row0 + 16) x 32 columns (+ bias) as a 16-bit [16][AT3_LD] tile; rows >= limit are zero.  Split in two
// so that the global loads are issued at the top of a tile iteration and their latency hides behind the fragment loads,
// the first product and the elementwise work; the tile is only written (and read back by ldmatrix) at the end.
struct TileRegs { uint4 v0, v1; };
template <int DT>
__device__ __forceinline__ TileRegs tile_load(const void* src, int64_t tok0, int row0, int limit, int64_t row_stride,
                                              int64_t col0, const float* bias, int lane) {
  const int row = lane >> 1, c0 = (lane & 1) * 16;
  TileRegs t;
  t.v0 = make_uint4(0, 0, 0, 0);
  t.v1 = t.v0;
  if (row0 + row < limit) {
    const uint4* s = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(src) + (tok0 + row0 + row) * row_stride +
                                                    col0 + c0);
    t.v0 = __ldg(s);
    t.v1 = __ldg(s + 1);
    if (bias) {
      const float4 b0 = ldg_f4(bias + col0 + c0), b1 = ldg_f4(bias + col0 + c0 + 4), b2 = ldg_f4(bias + col0 + c0 + 8),
                   b3 = ldg_f4(bias + col0 + c0 + 12);
      t.v0.x = add16x2<DT>(t.v0.x, pack16<DT>(b0.x, b0.y)); t.v0.y = add16x2<DT>(t.v0.y, pack16<DT>(b0.z, b0.w));
      t.v0.z = add16x2<DT>(t.v0.z, pack16<DT>(b1.x, b1.y)); t.v0.w = add16x2<DT>(t.v0.w, pack16<DT>(b1.z, b1.w));
      t.v1.x = add16x2<DT>(t.v1.x, pack16<DT>(b2.x, b2.y)); t.v1.y = add16x2<DT>(t.v1.y, pack16<DT>(b2.z, b2.w));
      t.v1.z = add16x2<DT>(t.v1.z, pack16<DT>(b3.x, b3.y)); t.v1.w = add16x2<DT>(t.v1.w, pack16<DT>(b3.z, b3.w));
    }
  }
  return t;
}
__device__ __forceinline__ void tile_store(uint16_t* dst, const TileRegs& t, int lane) {
  const int row = lane >> 1, c0 = (lane & 1) * 16;
  *reinterpret_cast<uint4*>(dst + row * AT3_LD + c0) = t.v0;
  *reinterpret_cast<uint4*>(dst + row * AT3_LD + c0 + 8) = t.v1;
}
// acc[4 n-tiles of 8 dims][4] += A[16 x 16 tokens] . T[16 tokens x 32 dims], T staged in shared memory
template <int DT>
__device__ __forceinline__ void mma_a_tile(float (&acc)[4][4], const uint32_t (&a)[4], const uint16_t* tile, int lane) {
  const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int np = 0; np < 2; ++np) {
    uint32_t r[4];
    ldsm_x4_trans(r, tile + ((mi & 1) * 8 + rr) * AT3_LD + 8 * (2 * np + (mi >> 1)));
    mma16816<DT>(acc[2 * np], a, r[0], r[1]);
    mma16816<DT>(acc[2 * np + 1], a, r[2], r[3]);
  }
}
// accumulator fragments [2 nt][4] (16 x 16, fp32) -> A fragments of the next product
template <int DT> __device__ __forceinline__ void c_to_a(uint32_t (&a)[4], const float (&c)[2][4]) {
  a[0] = pack16<DT>(c[0][0], c[0][1]);
  a[1] = pack16<DT>(c[0][2], c[0][3]);
  a[2] = pack16<DT>(c[1][0], c[1][1]);
  a[3] = pack16<DT>(c[1][2], c[1][3]);
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ------------------------------------------------------------------------------------------------ forward
template <int DT>
__global__ void __launch_bounds__(256, 3) attn3_fwd_kernel(const void* __restrict__ qkv, AttnParams p,
                                                        void* __restrict__ out, float* __restrict__ lse) {
  __shared__ __align__(16) uint16_t smem[8 * 16 * AT3_LD];
  p.seed = epoch_seed(p.seed);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  uint16_t* sT = smem + warp * 16 * AT3_LD;
  const int64_t n_items = p.full_to * p.H;              // the zero tail (queries with every key masked) is bulk-filled;
                                                        // sequences in [full_to, zero_from): attn_one_fwd_kernel
  const int64_t rs_ = 3 * (int64_t)p.H * ENC_HD, os_ = (int64_t)p.H * ENC_HD;
  const bool hb = p.bias != nullptr;
  {
    // sequences >= zero_from are the LAST ones: their tokens are the contiguous range [cu[zero_from], cu[n_seq])
    const int64_t z0 = __ldg(p.cu + p.zero_from), z1 = __ldg(p.cu + p.n_seq);
    const int64_t n2 = (z1 - z0) * os_ / 2;              // 32-bit words of `out`
    uint32_t* zo = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(out) + z0 * os_);
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += (int64_t)gridDim.x * blockDim.x) zo[k] = 0u;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < (z1 - z0) * p.H; k += (int64_t)gridDim.x * blockDim.x)
      lse[z0 * p.H + k] = 0.f;
  }
  const int64_t stride = (int64_t)gridDim.x * wpc;
  int64_t item = (int64_t)blockIdx.x * wpc + warp;
  int64_t t0n = 0, t1n = 0;
  if (item < n_items) { t0n = __ldg(p.cu + item / p.H); t1n = __ldg(p.cu + item / p.H + 1); }
  for (; item < n_items; item += stride) {
    const int h = (int)(item % p.H);
    const int64_t t0 = t0n;
    const int len = min((int)(t1n - t0n), p.max_len);
    if (item + stride < n_items) {                      // next item's bounds: in flight during this item
      t0n = __ldg(p.cu + (item + stride) / p.H);
      t1n = __ldg(p.cu + (item + stride) / p.H + 1);
    }
    uint32_t bq[2][2], bk[2][2];
    frag_bias<DT>(bq, p.bias, h * ENC_HD, t);
    frag_bias<DT>(bk, p.bias, os_ + h * ENC_HD, t);
    for (int qt = 0; qt * 16 < len; ++qt) {
      uint32_t qa[2][4];
      load_a<DT>(qa, qkv, t0, qt * 16, len, rs_, h * ENC_HD, bq, hb, g, t);
      float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f}, o[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
      const int i0 = qt * 16 + g;
      const uint32_t shi = (uint32_t)(p.seed >> 32);
      const uint32_t hr[2] = {rnd_row(p.seed, (uint32_t)((t0 + i0) * p.H + h)), rnd_row(p.seed, (uint32_t)((t0 + i0 + 8) * p.H + h))};
      for (int kt = 0; kt <= qt; ++kt) {
        const TileRegs vt = tile_load<DT>(qkv, t0, kt * 16, len, rs_, 2 * os_ + h * ENC_HD, p.bias, lane);    // V
        uint32_t kb[2][2][2];
        load_b<DT>(kb, qkv, t0, kt * 16, len, rs_, os_ + h * ENC_HD, bk, hb, g, t);
        float s[2][4];
        mma_abt<DT>(s, qa, kb);
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = i0 + 8 * (e >> 1), j = kt * 16 + 8 * nt + 2 * t + (e & 1);
            const float v = (j <= i && i < len) ? s[nt][e] * p.scale : -INFINITY;
            s[nt][e] = v;
            mx[e >> 1] = fmaxf(mx[e >> 1], v);
          }
        float c[2], mu[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float mn = fmaxf(m[r], quad_max(mx[r]));
          mu[r] = (mn == -INFINITY) ? 0.f : mn;
          c[r] = __expf(m[r] - mu[r]);
          m[r] = mn;
          l[r] *= c[r];
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = e >> 1;
            float pr = __expf(s[nt][e] - mu[r]);
            l[r] += pr;
            if (p.drop_thresh) {
              const int j = kt * 16 + 8 * nt + 2 * t + (e & 1);
              pr = (rnd_col(shi, hr[r], (uint32_t)j) >= p.drop_thresh) ? pr * p.inv_keep : 0.f;
            }
            s[nt][e] = pr;
          }
#pragma unroll
        for (int n = 0; n < 4; ++n) { o[n][0] *= c[0]; o[n][1] *= c[0]; o[n][2] *= c[1]; o[n][3] *= c[1]; }
        uint32_t pa[4];
        c_to_a<DT>(pa, s);
        __syncwarp();                                   // the previous tile's ldmatrix reads are done
        tile_store(sT, vt, lane);
        __syncwarp();
        mma_a_tile<DT>(o, pa, sT, lane);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = i0 + 8 * r;
        const float lt = quad_sum(l[r]);
        if (i < len) {
          const float inv = 1.f / lt;
          uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(out) + (t0 + i) * os_ + h * ENC_HD + 2 * t);
#pragma unroll
          for (int n = 0; n < 4; ++n) dst[4 * n] = pack16<DT>(o[n][2 * r] * inv, o[n][2 * r + 1] * inv);
          if (t == 0) lse[(t0 + i) * p.H + h] = m[r] + __logf(lt);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
template <int DT>
__global__ void __launch_bounds__(256, 2) attn3_bwd_kernel(const void* __restrict__ qkv, const void* __restrict__ d_out,
                                                        const void* __restrict__ out, const float* __restrict__ lse,
                                                        AttnParams p, void* __restrict__ d_qkv) {
  __shared__ __align__(16) uint16_t smem[8 * 2 * 16 * AT3_LD];
  __shared__ float sstat[8 * 128];
  p.seed = epoch_seed(p.seed);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  uint16_t* sT0 = smem + warp * 2 * 16 * AT3_LD;
  uint16_t* sT1 = sT0 + 16 * AT3_LD;
  float* sLse = sstat + warp * 128;
  float* sDelta = sLse + 64;
  const int64_t n_items = p.full_to * p.H;
  const int64_t rs_ = 3 * (int64_t)p.H * ENC_HD, os_ = (int64_t)p.H * ENC_HD;
  const bool hb = p.bias != nullptr;
  const uint32_t nob[2][2] = {{0u, 0u}, {0u, 0u}};
  {
    const int64_t z0 = __ldg(p.cu + p.zero_from), z1 = __ldg(p.cu + p.n_seq);
    const int64_t n2 = (z1 - z0) * rs_ / 2;
    uint32_t* zo = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(d_qkv) + z0 * rs_);
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += (int64_t)gridDim.x * blockDim.x) zo[k] = 0u;
  }
  const int64_t stride = (int64_t)gridDim.x * wpc;
  int64_t item = (int64_t)blockIdx.x * wpc + warp;
  int64_t t0n = 0, t1n = 0;
  if (item < n_items) { t0n = __ldg(p.cu + item / p.H); t1n = __ldg(p.cu + item / p.H + 1); }
  for (; item < n_items; item += stride) {
    const int h = (int)(item % p.H);
    const int64_t t0 = t0n;
    const int len = min((int)(t1n - t0n), p.max_len);
    if (item + stride < n_items) {
      t0n = __ldg(p.cu + (item + stride) / p.H);
      t1n = __ldg(p.cu + (item + stride) / p.H + 1);
    }
    if (p.short_split) continue;                        // (every sequence is served by attn3_bwd_short / _long_kernel)
    uint32_t bq[2][2], bk[2][2], bv[2][2];
    frag_bias<DT>(bq, p.bias, h * ENC_HD, t);
    frag_bias<DT>(bk, p.bias, os_ + h * ENC_HD, t);
    frag_bias<DT>(bv, p.bias, 2 * os_ + h * ENC_HD, t);
    __syncwarp();                                       // the previous item's row statistics are no longer read
    // ---------------- dQ phase (rows = queries); leaves lse_i and delta_i = <dO_i, O_i> in shared memory
    for (int qt = 0; qt * 16 < len; ++qt) {
      uint32_t qa[2][4], ga[2][4];
      load_a<DT>(qa, qkv, t0, qt * 16, len, rs_, h * ENC_HD, bq, hb, g, t);
      load_a<DT>(ga, d_out, t0, qt * 16, len, os_, h * ENC_HD, nob, false, g, t);
      const int i0 = qt * 16 + g;
      float li[2], di[2];
      {
        uint32_t oa[2][4];
        load_a<DT>(oa, out, t0, qt * 16, len, os_, h * ENC_HD, nob, false, g, t);
        float part[2] = {0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 x = unpack16<DT>(ga[ks][e]), y = unpack16<DT>(oa[ks][e]);
            part[e & 1] = fmaf(x.x, y.x, fmaf(x.y, y.y, part[e & 1]));
          }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          di[r] = quad_sum(part[r]);
          const int i = i0 + 8 * r;
          li[r] = (i < len) ? __ldg(lse + (t0 + i) * p.H + h) : 0.f;
          if (t == 0 && i < len) { sLse[i] = li[r]; sDelta[i] = di[r]; }
        }
      }
      float dq[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) dq[n][e] = 0.f;
      const uint32_t shi = (uint32_t)(p.seed >> 32);
      const uint32_t hr[2] = {rnd_row(p.seed, (uint32_t)((t0 + i0) * p.H + h)), rnd_row(p.seed, (uint32_t)((t0 + i0 + 8) * p.H + h))};
      for (int kt = 0; kt <= qt; ++kt) {
        const TileRegs kt_ = tile_load<DT>(qkv, t0, kt * 16, len, rs_, os_ + h * ENC_HD, p.bias, lane);       // K
        uint32_t kb[2][2][2], vb[2][2][2];
        load_b<DT>(kb, qkv, t0, kt * 16, len, rs_, os_ + h * ENC_HD, bk, hb, g, t);
        load_b<DT>(vb, qkv, t0, kt * 16, len, rs_, 2 * os_ + h * ENC_HD, bv, hb, g, t);
        float s[2][4], dp[2][4];
        mma_abt<DT>(s, qa, kb);
        mma_abt<DT>(dp, ga, vb);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = e >> 1;
            const int i = i0 + 8 * r, j = kt * 16 + 8 * nt + 2 * t + (e & 1);
            const float pr = (j <= i && i < len) ? __expf(s[nt][e] * p.scale - li[r]) : 0.f;
            float d = dp[nt][e];
            if (p.drop_thresh)
              d = (rnd_col(shi, hr[r], (uint32_t)j) >= p.drop_thresh) ? d * p.inv_keep : 0.f;
            s[nt][e] = pr * (d - di[r]);
          }
        uint32_t da[4];
        c_to_a<DT>(da, s);
        __syncwarp();
        tile_store(sT0, kt_, lane);
        __syncwarp();
        mma_a_tile<DT>(dq, da, sT0, lane);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = i0 + 8 * r;
        if (i < len) {
          uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(d_qkv) + (t0 + i) * rs_ + h * ENC_HD + 2 * t);
#pragma unroll
          for (int n = 0; n < 4; ++n) dst[4 * n] = pack16<DT>(dq[n][2 * r] * p.scale, dq[n][2 * r + 1] * p.scale);
        }
      }
    }
    __syncwarp();
    // ---------------- dK / dV phase (rows = keys): S^T = K Q^T, dP^T = V dO^T
    for (int kt = 0; kt * 16 < len; ++kt) {
      uint32_t ka[2][4], va[2][4];
      load_a<DT>(ka, qkv, t0, kt * 16, len, rs_, os_ + h * ENC_HD, bk, hb, g, t);
      load_a<DT>(va, qkv, t0, kt * 16, len, rs_, 2 * os_ + h * ENC_HD, bv, hb, g, t);
      const int j0 = kt * 16 + g;
      float dk[4][4], dv[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dk[n][e] = 0.f; dv[n][e] = 0.f; }
      for (int qt = kt; qt * 16 < len; ++qt) {
        const TileRegs qt_ = tile_load<DT>(qkv, t0, qt * 16, len, rs_, h * ENC_HD, p.bias, lane);             // Q
        const TileRegs gt_ = tile_load<DT>(d_out, t0, qt * 16, len, os_, h * ENC_HD, nullptr, lane);          // dO
        uint32_t qb[2][2][2], gb[2][2][2];
        load_b<DT>(qb, qkv, t0, qt * 16, len, rs_, h * ENC_HD, bq, hb, g, t);
        load_b<DT>(gb, d_out, t0, qt * 16, len, os_, h * ENC_HD, nob, false, g, t);
        float s[2][4], dp[2][4];
        mma_abt<DT>(s, ka, qb);
        mma_abt<DT>(dp, va, gb);
        float pk[2][4];
        const uint32_t shi = (uint32_t)(p.seed >> 32);
        uint32_t hq[2][2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int c = 0; c < 2; ++c) hq[nt][c] = rnd_row(p.seed, (uint32_t)((t0 + qt * 16 + 8 * nt + 2 * t + c) * p.H + h));
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = j0 + 8 * (e >> 1), i = qt * 16 + 8 * nt + 2 * t + (e & 1);
            const bool ok = (j <= i) && (i < len);
            const float pr = ok ? __expf(s[nt][e] * p.scale - sLse[i & 63]) : 0.f;
            float d = dp[nt][e], q = pr;
            if (p.drop_thresh) {
              const bool keep = rnd_col(shi, hq[nt][e & 1], (uint32_t)j) >= p.drop_thresh;
              d = keep ? d * p.inv_keep : 0.f;
              q = keep ? pr * p.inv_keep : 0.f;
            }
            pk[nt][e] = q;
            s[nt][e] = ok ? pr * (d - sDelta[i & 63]) : 0.f;
          }
        uint32_t pa[4], da[4];
        c_to_a<DT>(pa, pk);
        c_to_a<DT>(da, s);
        __syncwarp();
        tile_store(sT0, qt_, lane);
        tile_store(sT1, gt_, lane);
        __syncwarp();
        mma_a_tile<DT>(dv, pa, sT1, lane);
        mma_a_tile<DT>(dk, da, sT0, lane);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int j = j0 + 8 * r;
        if (j < len) {
          uint16_t* row = reinterpret_cast<uint16_t*>(d_qkv) + (t0 + j) * rs_ + h * ENC_HD + 2 * t;
          uint32_t* dkp = reinterpret_cast<uint32_t*>(row + os_);
          uint32_t* dvp = reinterpret_cast<uint32_t*>(row + 2 * os_);
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            dkp[4 * n] = pack16<DT>(dk[n][2 * r] * p.scale, dk[n][2 * r + 1] * p.scale);
            dvp[4 * n] = pack16<DT>(dv[n][2 * r], dv[n][2 * r + 1]);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward, one tile
// 72 % of the sequences have <= 16 tokens: one (query, key) tile.  The two-phase kernel above computes S and dP twice for
// them (rows = queries for dQ, rows = keys for dK / dV).  Here they are computed once; the A operands of the key-side
// products, P^T and dS^T, come from transposing the 8 x 8 blocks of the packed fragments (movmatrix) -- 12 fewer mma,
// 32 fewer fragment loads, half the exponentials and dropout hashes per item.  (No in_proj bias: since r02f it rides in
// the GEMM epilogue; the host falls back to the two-phase kernel when a bias is given.)
__device__ __forceinline__ uint32_t movm_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
template <int DT>
__global__ void __launch_bounds__(256, 2) attn3_bwd_short_kernel(const void* __restrict__ qkv, const void* __restrict__ d_out,
                                                                 const void* __restrict__ out, const float* __restrict__ lse,
                                                                 AttnParams p, void* __restrict__ d_qkv) {
  __shared__ __align__(16) uint16_t smem[8 * 3 * 16 * AT3_LD];
  p.seed = epoch_seed(p.seed);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  uint16_t* sK = smem + warp * 3 * 16 * AT3_LD;
  uint16_t* sQ = sK + 16 * AT3_LD;
  uint16_t* sG = sQ + 16 * AT3_LD;
  const int64_t n_items = p.full_to * p.H;
  const int64_t rs_ = 3 * (int64_t)p.H * ENC_HD, os_ = (int64_t)p.H * ENC_HD;
  const uint32_t nob[2][2] = {{0u, 0u}, {0u, 0u}};
  const int64_t stride = (int64_t)gridDim.x * wpc;
  int64_t item = (int64_t)blockIdx.x * wpc + warp;
  int64_t t0n = 0, t1n = 0;
  if (item < n_items) { t0n = __ldg(p.cu + item / p.H); t1n = __ldg(p.cu + item / p.H + 1); }
  for (; item < n_items; item += stride) {
    const int h = (int)(item % p.H);
    const int64_t t0 = t0n;
    const int len = min((int)(t1n - t0n), p.max_len);
    if (item + stride < n_items) {
      t0n = __ldg(p.cu + (item + stride) / p.H);
      t1n = __ldg(p.cu + (item + stride) / p.H + 1);
    }
    if (len > 16 || len <= 0) continue;
    // every global load of the item is issued here
    const TileRegs kt_ = tile_load<DT>(qkv, t0, 0, len, rs_, os_ + h * ENC_HD, nullptr, lane);
    const TileRegs qt_ = tile_load<DT>(qkv, t0, 0, len, rs_, h * ENC_HD, nullptr, lane);
    const TileRegs gt_ = tile_load<DT>(d_out, t0, 0, len, os_, h * ENC_HD, nullptr, lane);
    uint32_t qa[2][4], ga[2][4], oa[2][4], kb[2][2][2], vb[2][2][2];
    load_a<DT>(qa, qkv, t0, 0, len, rs_, h * ENC_HD, nob, false, g, t);
    load_a<DT>(ga, d_out, t0, 0, len, os_, h * ENC_HD, nob, false, g, t);
    load_a<DT>(oa, out, t0, 0, len, os_, h * ENC_HD, nob, false, g, t);
    load_b<DT>(kb, qkv, t0, 0, len, rs_, os_ + h * ENC_HD, nob, false, g, t);
    load_b<DT>(vb, qkv, t0, 0, len, rs_, 2 * os_ + h * ENC_HD, nob, false, g, t);
    float li[2], di[2];
    {
      float part[2] = {0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 x = unpack16<DT>(ga[ks][e]), y = unpack16<DT>(oa[ks][e]);
          part[e & 1] = fmaf(x.x, y.x, fmaf(x.y, y.y, part[e & 1]));
        }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        di[r] = quad_sum(part[r]);
        const int i = g + 8 * r;
        li[r] = (i < len) ? __ldg(lse + (t0 + i) * p.H + h) : 0.f;
      }
    }
    __syncwarp();                                       // the previous item's ldmatrix reads are done
    tile_store(sK, kt_, lane);
    tile_store(sQ, qt_, lane);
    tile_store(sG, gt_, lane);
    float s[2][4], dp[2][4], pk[2][4];
    mma_abt<DT>(s, qa, kb);
    mma_abt<DT>(dp, ga, vb);
    const uint32_t shi = (uint32_t)(p.seed >> 32);
    const uint32_t hr[2] = {rnd_row(p.seed, (uint32_t)((t0 + g) * p.H + h)), rnd_row(p.seed, (uint32_t)((t0 + g + 8) * p.H + h))};
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const int i = g + 8 * r, j = 8 * nt + 2 * t + (e & 1);
        const float pr = (j <= i && i < len) ? __expf(s[nt][e] * p.scale - li[r]) : 0.f;
        float d = dp[nt][e], q = pr;
        if (p.drop_thresh) {
          const bool keep = rnd_col(shi, hr[r], (uint32_t)j) >= p.drop_thresh;
          d = keep ? d * p.inv_keep : 0.f;
          q = keep ? pr * p.inv_keep : 0.f;
        }
        pk[nt][e] = q;
        s[nt][e] = pr * (d - di[r]);
      }
    uint32_t pa[4], da[4];
    c_to_a<DT>(pa, pk);
    c_to_a<DT>(da, s);
    const uint32_t paT[4] = {movm_trans(pa[0]), movm_trans(pa[2]), movm_trans(pa[1]), movm_trans(pa[3])};
    const uint32_t daT[4] = {movm_trans(da[0]), movm_trans(da[2]), movm_trans(da[1]), movm_trans(da[3])};
    float dq[4][4], dk[4][4], dv[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) { dq[n][e] = 0.f; dk[n][e] = 0.f; dv[n][e] = 0.f; }
    __syncwarp();                                       // the three tiles are in shared memory
    mma_a_tile<DT>(dq, da, sK, lane);
    mma_a_tile<DT>(dv, paT, sG, lane);
    mma_a_tile<DT>(dk, daT, sQ, lane);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = g + 8 * r;
      if (i < len) {
        uint16_t* row = reinterpret_cast<uint16_t*>(d_qkv) + (t0 + i) * rs_ + h * ENC_HD + 2 * t;
        uint32_t* dqp = reinterpret_cast<uint32_t*>(row);
        uint32_t* dkp = reinterpret_cast<uint32_t*>(row + os_);
        uint32_t* dvp = reinterpret_cast<uint32_t*>(row + 2 * os_);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          dqp[4 * n] = pack16<DT>(dq[n][2 * r] * p.scale, dq[n][2 * r + 1] * p.scale);
          dkp[4 * n] = pack16<DT>(dk[n][2 * r] * p.scale, dk[n][2 * r + 1] * p.scale);
          dvp[4 * n] = pack16<DT>(dv[n][2 * r], dv[n][2 * r + 1]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward, 17..64 tokens
// The same idea for longer sequences: every (query tile, key tile) pair is visited ONCE (outer loop over key tiles, inner
// over the query tiles at or behind it): S, dP, P and dS are formed once per pair and feed dQ (dS . K), dV (P^T . dO) and
// dK (dS^T . Q) right away.  dK / dV of the key tile stay in registers across the inner loop; the dQ contributions of
// the different key tiles are added up in a per-warp fp32 buffer in shared memory (each lane only ever touches the
// fragment elements it owns: no synchronisation), and lse / delta of every row are computed during the first key tile
// and kept in shared memory.  Per pair: 20 mma instead of 28, half the fragment loads, half the exponentials / hashes.
#define AT3L_DQ_LD 34
#define AT3L_WARP_BYTES (3 * 16 * AT3_LD * 2 + 64 * AT3L_DQ_LD * 4 + 2 * 64 * 4)
template <int DT>
__global__ void __launch_bounds__(256, 2) attn3_bwd_long_kernel(const void* __restrict__ qkv, const void* __restrict__ d_out,
                                                                const void* __restrict__ out, const float* __restrict__ lse,
                                                                AttnParams p, void* __restrict__ d_qkv) {
  extern __shared__ __align__(16) unsigned char at3l_smem[];
  p.seed = epoch_seed(p.seed);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  unsigned char* wbase = at3l_smem + (size_t)warp * AT3L_WARP_BYTES;
  uint16_t* sK = reinterpret_cast<uint16_t*>(wbase);
  uint16_t* sQ = sK + 16 * AT3_LD;
  uint16_t* sG = sQ + 16 * AT3_LD;
  float* sDq = reinterpret_cast<float*>(sG + 16 * AT3_LD);
  float* sLse = sDq + 64 * AT3L_DQ_LD;
  float* sDelta = sLse + 64;
  const int64_t n_items = p.full_to * p.H;
  const int64_t rs_ = 3 * (int64_t)p.H * ENC_HD, os_ = (int64_t)p.H * ENC_HD;
  const uint32_t nob[2][2] = {{0u, 0u}, {0u, 0u}};
  const uint32_t shi = (uint32_t)(p.seed >> 32);
  const int64_t stride = (int64_t)gridDim.x * wpc;
  int64_t item = (int64_t)blockIdx.x * wpc + warp;
  int64_t t0n = 0, t1n = 0;
  if (item < n_items) { t0n = __ldg(p.cu + item / p.H); t1n = __ldg(p.cu + item / p.H + 1); }
  for (; item < n_items; item += stride) {
    const int h = (int)(item % p.H);
    const int64_t t0 = t0n;
    const int len = min((int)(t1n - t0n), p.max_len);
    if (item + stride < n_items) {
      t0n = __ldg(p.cu + (item + stride) / p.H);
      t1n = __ldg(p.cu + (item + stride) / p.H + 1);
    }
    if (len <= 16) continue;                            // attn3_bwd_short_kernel
    const int ntile = (len + 15) >> 4;
    for (int kt = 0; kt < ntile; ++kt) {
      const TileRegs kt_ = tile_load<DT>(qkv, t0, kt * 16, len, rs_, os_ + h * ENC_HD, nullptr, lane);
      uint32_t kb[2][2][2], vb[2][2][2];
      load_b<DT>(kb, qkv, t0, kt * 16, len, rs_, os_ + h * ENC_HD, nob, false, g, t);
      load_b<DT>(vb, qkv, t0, kt * 16, len, rs_, 2 * os_ + h * ENC_HD, nob, false, g, t);
      float dk[4][4], dv[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dk[n][e] = 0.f; dv[n][e] = 0.f; }
      __syncwarp();                                     // the previous key tile's ldmatrix reads of sK are done
      tile_store(sK, kt_, lane);
      for (int qt = kt; qt < ntile; ++qt) {
        const TileRegs qt_ = tile_load<DT>(qkv, t0, qt * 16, len, rs_, h * ENC_HD, nullptr, lane);
        const TileRegs gt_ = tile_load<DT>(d_out, t0, qt * 16, len, os_, h * ENC_HD, nullptr, lane);
        uint32_t qa[2][4], ga[2][4];
        load_a<DT>(qa, qkv, t0, qt * 16, len, rs_, h * ENC_HD, nob, false, g, t);
        load_a<DT>(ga, d_out, t0, qt * 16, len, os_, h * ENC_HD, nob, false, g, t);
        const int i0 = qt * 16 + g;
        float li[2], di[2];
        if (kt == 0) {                                  // first visit of this query tile: row statistics
          uint32_t oa[2][4];
          load_a<DT>(oa, out, t0, qt * 16, len, os_, h * ENC_HD, nob, false, g, t);
          float part[2] = {0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 x = unpack16<DT>(ga[ks][e]), y = unpack16<DT>(oa[ks][e]);
              part[e & 1] = fmaf(x.x, y.x, fmaf(x.y, y.y, part[e & 1]));
            }
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            di[r] = quad_sum(part[r]);
            const int i = i0 + 8 * r;
            li[r] = (i < len) ? __ldg(lse + (t0 + i) * p.H + h) : 0.f;
            if (t == 0) { sLse[i] = li[r]; sDelta[i] = di[r]; }
          }
        } else {
#pragma unroll
          for (int r = 0; r < 2; ++r) { li[r] = sLse[i0 + 8 * r]; di[r] = sDelta[i0 + 8 * r]; }
        }
        float s[2][4], dp[2][4], pk[2][4];
        mma_abt<DT>(s, qa, kb);
        mma_abt<DT>(dp, ga, vb);
        const uint32_t hr[2] = {rnd_row(p.seed, (uint32_t)((t0 + i0) * p.H + h)), rnd_row(p.seed, (uint32_t)((t0 + i0 + 8) * p.H + h))};
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = e >> 1;
            const int i = i0 + 8 * r, j = kt * 16 + 8 * nt + 2 * t + (e & 1);
            const float pr = (j <= i && i < len) ? __expf(s[nt][e] * p.scale - li[r]) : 0.f;
            float d = dp[nt][e], q = pr;
            if (p.drop_thresh) {
              const bool keep = rnd_col(shi, hr[r], (uint32_t)j) >= p.drop_thresh;
              d = keep ? d * p.inv_keep : 0.f;
              q = keep ? pr * p.inv_keep : 0.f;
            }
            pk[nt][e] = q;
            s[nt][e] = pr * (d - di[r]);
          }
        uint32_t pa[4], da[4];
        c_to_a<DT>(pa, pk);
        c_to_a<DT>(da, s);
        __syncwarp();                                   // the previous pair's ldmatrix reads of sQ / sG are done
        tile_store(sQ, qt_, lane);
        tile_store(sG, gt_, lane);
        __syncwarp();
        {
          float dq[4][4];
#pragma unroll
          for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) dq[n][e] = 0.f;
          mma_a_tile<DT>(dq, da, sK, lane);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            float* row = sDq + (i0 + 8 * r) * AT3L_DQ_LD + 2 * t;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              float2 v = make_float2(dq[n][2 * r], dq[n][2 * r + 1]);
              if (kt > 0) { const float2 o = *reinterpret_cast<const float2*>(row + 8 * n); v.x += o.x; v.y += o.y; }
              *reinterpret_cast<float2*>(row + 8 * n) = v;
            }
          }
        }
        const uint32_t paT[4] = {movm_trans(pa[0]), movm_trans(pa[2]), movm_trans(pa[1]), movm_trans(pa[3])};
        const uint32_t daT[4] = {movm_trans(da[0]), movm_trans(da[2]), movm_trans(da[1]), movm_trans(da[3])};
        mma_a_tile<DT>(dv, paT, sG, lane);
        mma_a_tile<DT>(dk, daT, sQ, lane);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int j = kt * 16 + g + 8 * r;
        if (j < len) {
          uint16_t* row = reinterpret_cast<uint16_t*>(d_qkv) + (t0 + j) * rs_ + h * ENC_HD + 2 * t;
          uint32_t* dkp = reinterpret_cast<uint32_t*>(row + os_);
          uint32_t* dvp = reinterpret_cast<uint32_t*>(row + 2 * os_);
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            dkp[4 * n] = pack16<DT>(dk[n][2 * r] * p.scale, dk[n][2 * r + 1] * p.scale);
            dvp[4 * n] = pack16<DT>(dv[n][2 * r], dv[n][2 * r + 1]);
          }
        }
      }
    }
    // dQ: every lane writes back the fragment elements it accumulated (its own addresses: program order suffices)
    for (int qt = 0; qt < ntile; ++qt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = qt * 16 + g + 8 * r;
        if (i < len) {
          const float* row = sDq + i * AT3L_DQ_LD + 2 * t;
          uint32_t* dqp = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(d_qkv) + (t0 + i) * rs_ + h * ENC_HD + 2 * t);
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const float2 v = *reinterpret_cast<const float2*>(row + 8 * n);
            dqp[4 * n] = pack16<DT>(v.x * p.scale, v.y * p.scale);
          }
        }
      }
  }
}
