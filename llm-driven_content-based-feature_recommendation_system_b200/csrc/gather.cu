// Row gathers and the fused tower "fronts" (forward side).  HBM-bound byte/row movers:
// 128-bit coalesced row loads, one warp (or a power-of-two lane group for narrow rows) per
// row, streaming (no-L1-allocate) loads for read-once operands, grid = multiple of 148 SMs.
#include "common.cuh"
#include "../../include/rs_twotower.h"

namespace rs {

// lane group geometry for rows of `dim` fp32 elements (dim % 4 == 0)
struct RowMap {
  int vecs;  // float4 per row
  int lpr;   // lanes per row (power of two <= 32)
  int rpw;   // rows per warp
};
__host__ __device__ inline RowMap row_map(int64_t dim) {
  RowMap m;
  m.vecs = (int)(dim >> 2);
  m.lpr = 1;
  while (m.lpr < m.vecs && m.lpr < 32) m.lpr <<= 1;
  m.rpw = 32 / m.lpr;
  return m;
}
__device__ __forceinline__ float group_sum(float v, int lpr) {
  for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ gather_rows
template <int TD, int OD>
__global__ void __launch_bounds__(256) gather_rows_kernel(const void* __restrict__ table, int64_t rows, int64_t dim,
                                                          const int64_t* __restrict__ ids, int64_t n,
                                                          int64_t clamp_max, void* __restrict__ out,
                                                          int* __restrict__ oob) {
  const RowMap m = row_map(dim);
  const int lane = threadIdx.x & 31;
  const int grp = lane / m.lpr, gl = lane % m.lpr;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r0 = warp * m.rpw; r0 < n; r0 += nwarps * m.rpw) {
    const int64_t r = r0 + grp;
    if (r >= n) continue;
    int64_t id = __ldg(ids + r);
    if (clamp_max >= 0 && id > clamp_max) id = clamp_max;
    const bool ok = (id >= 0) && (id < rows);
    if (!ok && id != -1 && oob && gl == 0) *oob = 1;      // id -1 is the null id: reads as zeros, never flagged
    for (int v = gl; v < m.vecs; v += m.lpr) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        if constexpr (TD == RS_F32) x = ldg_f4(reinterpret_cast<const float*>(table) + id * dim + 4 * v);
        else x = load4<TD>(table, id * dim + 4 * v);
      }
      store4<OD>(out, r * dim + 4 * v, x);
    }
  }
}

// ------------------------------------------------------------------ scatter_add_rows (atomic)
template <int GD>
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const void* __restrict__ d_out,
                                                               const int64_t* __restrict__ ids, int64_t n,
                                                               int64_t dim, int64_t rows, int64_t padding_idx,
                                                               int64_t clamp_max, float scale,
                                                               float* __restrict__ d_table, int* __restrict__ oob) {
  const RowMap m = row_map(dim);
  const int lane = threadIdx.x & 31;
  const int grp = lane / m.lpr, gl = lane % m.lpr;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r0 = warp * m.rpw; r0 < n; r0 += nwarps * m.rpw) {
    const int64_t r = r0 + grp;
    if (r >= n) continue;
    int64_t id = __ldg(ids + r);
    if (clamp_max >= 0 && id > clamp_max) id = clamp_max;
    if (id == padding_idx) continue;
    if (id < 0 || id >= rows) {
      if (oob && gl == 0) *oob = 1;
      continue;
    }
    for (int v = gl; v < m.vecs; v += m.lpr) {
      float4 g = load4<GD>(d_out, r * dim + 4 * v);
      g.x *= scale; g.y *= scale; g.z *= scale; g.w *= scale;
      red_add_f4(d_table + id * dim + 4 * v, g);
    }
  }
}

// ------------------------------------------------------------------ seq_front_fwd (U1)
struct SeqFrontParams {
  const int64_t* ids[RS_MAX_TABLES];
  const float* tables[RS_MAX_TABLES];
  int64_t rows[RS_MAX_TABLES];
  int n_tables;
};

template <int BD, int OD>
__global__ void __launch_bounds__(256) seq_front_fwd_kernel(const void* __restrict__ base, SeqFrontParams prm,
                                                            const float* __restrict__ gates,
                                                            const float* __restrict__ pos_table, int64_t L, int64_t P,
                                                            int64_t dim, void* __restrict__ out,
                                                            int* __restrict__ oob) {
  const RowMap m = row_map(dim);
  const int lane = threadIdx.x & 31;
  const int grp = lane / m.lpr, gl = lane % m.lpr;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float g[RS_MAX_TABLES];
#pragma unroll
  for (int t = 0; t < RS_MAX_TABLES; ++t) g[t] = (t < prm.n_tables) ? __ldg(gates + t) : 0.f;

  for (int64_t p0 = warp * m.rpw; p0 < P; p0 += nwarps * m.rpw) {
    const int64_t p = p0 + grp;
    if (p >= P) continue;
    int64_t id[RS_MAX_TABLES];
#pragma unroll
    for (int t = 0; t < RS_MAX_TABLES; ++t) {
      id[t] = -1;
      if (t < prm.n_tables && g[t] != 0.f) {      // a gate that is exactly 0 contributes exactly +0: skip the read
        id[t] = __ldg(prm.ids[t] + p);
        if (id[t] < 0 || id[t] >= prm.rows[t]) {
          if (oob && gl == 0) *oob = 1;
          id[t] = -1;
        }
      }
    }
    const int64_t l = p % L;
    for (int v = gl; v < m.vecs; v += m.lpr) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (base) acc = load4<BD>(base, p * dim + 4 * v);
      float4 row[RS_MAX_TABLES];
#pragma unroll
      for (int t = 0; t < RS_MAX_TABLES; ++t)      // issue every row load before the dependent adds
        if (id[t] >= 0) row[t] = ldg_f4(prm.tables[t] + id[t] * dim + 4 * v);
      float4 pr = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pos_table) pr = ldg_f4(pos_table + l * dim + 4 * v);
#pragma unroll
      for (int t = 0; t < RS_MAX_TABLES; ++t)
        if (id[t] >= 0) acc = mul_add_rn(acc, row[t], g[t]);
      if (pos_table) acc = add4(acc, pr);
      store4<OD>(out, p * dim + 4 * v, acc);
    }
  }
}

// dim == 128 fast path: a warp owns 32 consecutive positions.  The ids of all 32 are fetched with ONE coalesced
// load per live table and handed round by shuffles; rows are then fetched SF_UNROLL positions at a time so that
// each lane keeps SF_UNROLL x (base + NT rows) 128-bit loads in flight (the row address depends on the id:
// without this the kernel is bound by two serial memory latencies per position, not by bandwidth).
// NT = number of tables (compile time, so the per-table state is exactly sized and occupancy stays high).
#define SF_UNROLL 4
template <int BD, int OD, int NT>
__global__ void __launch_bounds__(256, 3) seq_front_fwd128_kernel(const void* __restrict__ base, SeqFrontParams prm,
                                                                  const float* __restrict__ gates,
                                                                  const float* __restrict__ pos_table, int L,
                                                                  int64_t P, void* __restrict__ out,
                                                                  int* __restrict__ oob) {
  constexpr int D = 128;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float g[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) g[t] = __ldg(gates + t);
  for (int64_t pb = warp * 32; pb < P; pb += nwarps * 32) {
    const int64_t mine = pb + lane;
    int myid[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      myid[t] = -1;
      if (g[t] != 0.f && mine < P) {                 // a gate that is exactly 0 contributes exactly +0: not read
        const int64_t x = __ldg(prm.ids[t] + mine);
        if (x < 0 || x >= prm.rows[t]) { if (oob) *oob = 1; }
        else myid[t] = (int)x;
      }
    }
    const int cnt = (int)((P - pb) < 32 ? (P - pb) : 32);
    const int l0 = (int)(pb % L);
    for (int u0 = 0; u0 < cnt; u0 += SF_UNROLL) {
      float4 acc[SF_UNROLL], row[SF_UNROLL][NT];
      int id[SF_UNROLL][NT];
#pragma unroll
      for (int u = 0; u < SF_UNROLL; ++u) {            // every load of the group is issued before any use
        const int64_t p = pb + u0 + u;
        const bool on = u0 + u < cnt;
        acc[u] = (on && base) ? load4<BD>(base, p * D + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          id[u][t] = __shfl_sync(0xffffffffu, myid[t], (u0 + u) & 31);
          if (!on) id[u][t] = -1;
          if (id[u][t] >= 0) row[u][t] = ldg_f4(prm.tables[t] + (int64_t)id[u][t] * D + 4 * lane);
        }
      }
#pragma unroll
      for (int u = 0; u < SF_UNROLL; ++u) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
          if (id[u][t] >= 0) acc[u] = mul_add_rn(acc[u], row[u][t], g[t]);
        if (u0 + u < cnt) {
          // the L position rows stay in L1: fetched at the point of use
          if (pos_table) acc[u] = add4(acc[u], ldg_f4(pos_table + ((l0 + u0 + u) % L) * D + 4 * lane));
          store4<OD>(out, (pb + u0 + u) * D + 4 * lane, acc[u]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ normalized rows (U4)
template <int OD>
__global__ void __launch_bounds__(256) normalized_rows_fwd_kernel(const float* __restrict__ table, int64_t rows,
                                                                  int64_t dim, const int64_t* __restrict__ ids,
                                                                  int64_t n, float eps, void* __restrict__ out,
                                                                  float* __restrict__ inv_norm, int* __restrict__ oob) {
  const RowMap m = row_map(dim);
  const int lane = threadIdx.x & 31;
  const int grp = lane / m.lpr, gl = lane % m.lpr;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r0 = warp * m.rpw; r0 < n; r0 += nwarps * m.rpw) {   // whole warp iterates together (shuffles)
    const int64_t r = r0 + grp;
    const bool live = r < n;
    int64_t id = live ? __ldg(ids + r) : 0;
    bool ok = live && id >= 0 && id < rows;
    if (live && !ok && id != -1 && oob && gl == 0) *oob = 1;      // (-1: null id, zeros)
    float ss = 0.f;
    for (int v = gl; v < m.vecs; v += m.lpr) {
      if (ok) { float4 x = ldg_f4(table + id * dim + 4 * v); ss += dot4(x, x); }
    }
    ss = group_sum(ss, m.lpr);
    const float denom = fmaxf(sqrtf(ss), eps);
    if (live && gl == 0 && inv_norm) inv_norm[r] = 1.0f / denom;
    for (int v = gl; v < m.vecs; v += m.lpr) {
      if (!live) continue;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) x = ldg_f4(table + id * dim + 4 * v);            // second read hits L1/L2
      x.x = __fdiv_rn(x.x, denom); x.y = __fdiv_rn(x.y, denom); x.z = __fdiv_rn(x.z, denom); x.w = __fdiv_rn(x.w, denom);
      store4<OD>(out, r * dim + 4 * v, x);
    }
  }
}

template <int GD>
__global__ void __launch_bounds__(256) normalized_rows_bwd_kernel(const void* __restrict__ d_out,
                                                                  const float* __restrict__ table, int64_t rows,
                                                                  int64_t dim, const int64_t* __restrict__ ids,
                                                                  int64_t n, float eps, float* __restrict__ d_table) {
  const RowMap m = row_map(dim);
  const int lane = threadIdx.x & 31;
  const int grp = lane / m.lpr, gl = lane % m.lpr;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r0 = warp * m.rpw; r0 < n; r0 += nwarps * m.rpw) {
    const int64_t r = r0 + grp;
    const bool live = r < n;
    int64_t id = live ? __ldg(ids + r) : 0;
    const bool ok = live && id >= 0 && id < rows;
    float ss = 0.f, wg = 0.f;
    for (int v = gl; v < m.vecs; v += m.lpr) {
      if (ok) {
        float4 w = ldg_f4(table + id * dim + 4 * v);
        float4 g = load4<GD>(d_out, r * dim + 4 * v);
        ss += dot4(w, w);
        wg += dot4(w, g);
      }
    }
    ss = group_sum(ss, m.lpr);
    wg = group_sum(wg, m.lpr);
    const float nrm = sqrtf(ss);
    const bool clamped = nrm < eps;                   // clamp_min(eps) has zero slope below eps
    const float inv = 1.0f / fmaxf(nrm, eps);
    const float c = clamped ? 0.f : wg * inv * inv * inv;   // <v,g>/denom * v  ==  w * <w,g> / denom^3
    for (int v = gl; v < m.vecs; v += m.lpr) {
      if (!ok) continue;
      float4 w = ldg_f4(table + id * dim + 4 * v);
      float4 g = load4<GD>(d_out, r * dim + 4 * v);
      float4 d = make_float4(g.x * inv - w.x * c, g.y * inv - w.y * c, g.z * inv - w.z * c, g.w * inv - w.w * c);
      red_add_f4(d_table + id * dim + 4 * v, d);
    }
  }
}

// ------------------------------------------------------------------ LayerNorm'd fronts (I1, I2)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser: counter-based RNG for dropout
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint64_t seed, uint64_t idx) {
  return (float)(mix64(seed ^ mix64(idx)) >> 40) * (1.0f / 16777216.0f);
}

// out[r,:] = LN(rowA[idA] + rowB (+ rowC)) ; NV = float4 per lane kept in registers (dim <= 128*NV)
template <int OD, int NV>
__global__ void __launch_bounds__(256) ln_front_kernel(const float* __restrict__ table, int64_t rows, int64_t dim,
                                                       const int64_t* __restrict__ ids, int64_t n,
                                                       const float* __restrict__ addB, int64_t periodB,
                                                       const float* __restrict__ addC,
                                                       const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                       float eps, float dropout_p, uint64_t seed,
                                                       void* __restrict__ out, float* __restrict__ mean_out,
                                                       float* __restrict__ rstd_out, int* __restrict__ oob) {
  const RowMap m = row_map(dim);
  const int lane = threadIdx.x & 31;
  const int grp = lane / m.lpr, gl = lane % m.lpr;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float inv_d = 1.0f / (float)dim;
  const float keep_scale = dropout_p > 0.f ? 1.0f / (1.0f - dropout_p) : 1.0f;
  for (int64_t r0 = warp * m.rpw; r0 < n; r0 += nwarps * m.rpw) {
    const int64_t r = r0 + grp;
    const bool live = r < n;
    int64_t id = live ? __ldg(ids + r) : 0;
    const bool ok = live && id >= 0 && id < rows;
    if (live && !ok && oob && gl == 0) *oob = 1;
    const int64_t rb = periodB > 0 ? (r % periodB) : 0;
    float4 x[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = gl + i * m.lpr;
      x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live && v < m.vecs) {
        if (ok) x[i] = ldg_f4(table + id * dim + 4 * v);
        if (addC) x[i] = add4(x[i], ldg_f4(addC + 4 * v));             // BERT: (word + type) + pos
        if (addB) x[i] = add4(x[i], ldg_f4(addB + rb * dim + 4 * v));
        s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
      }
    }
    const float mean = group_sum(s, m.lpr) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = gl + i * m.lpr;
      if (live && v < m.vecs) {
        float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = rsqrtf(group_sum(q, m.lpr) * inv_d + eps);
    if (live && gl == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = gl + i * m.lpr;
      if (live && v < m.vecs) {
        float4 w = ldg_f4(ln_w + 4 * v), b = ldg_f4(ln_b + 4 * v);
        float4 y = make_float4((x[i].x - mean) * rstd * w.x + b.x, (x[i].y - mean) * rstd * w.y + b.y,
                               (x[i].z - mean) * rstd * w.z + b.z, (x[i].w - mean) * rstd * w.w + b.w);
        if (dropout_p > 0.f) {
          const uint64_t e = (uint64_t)r * (uint64_t)dim + 4ull * v;
          y.x = u01(seed, e + 0) < dropout_p ? 0.f : y.x * keep_scale;
          y.y = u01(seed, e + 1) < dropout_p ? 0.f : y.y * keep_scale;
          y.z = u01(seed, e + 2) < dropout_p ? 0.f : y.z * keep_scale;
          y.w = u01(seed, e + 3) < dropout_p ? 0.f : y.w * keep_scale;
        }
        store4<OD>(out, r * dim + 4 * v, y);
      }
    }
  }
}

// ------------------------------------------------------------------ masked mean (I2 tail)
template <int FD>
__global__ void __launch_bounds__(256) masked_mean_fwd_kernel(const void* __restrict__ feats,
                                                              const int64_t* __restrict__ mask, int64_t n_seq,
                                                              int64_t T, int64_t dim, float* __restrict__ out) {
  const RowMap m = row_map(dim);
  const int lane = threadIdx.x & 31;
  const int grp = lane / m.lpr, gl = lane % m.lpr;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r0 = warp * m.rpw; r0 < n_seq; r0 += nwarps * m.rpw) {
    const int64_t r = r0 + grp;
    if (r >= n_seq) continue;
    for (int v = gl; v < m.vecs; v += m.lpr) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float cnt = 0.f;
      for (int64_t t = 0; t < T; ++t) {
        const float mk = (float)__ldg(mask + r * T + t);
        cnt += mk;
        if (mk != 0.f) acc = fma4(acc, load4<FD>(feats, (r * T + t) * dim + 4 * v), mk);
      }
      const float den = fmaxf(cnt, 1e-9f);
      float4 o = make_float4(acc.x / den, acc.y / den, acc.z / den, acc.w / den);
      *reinterpret_cast<float4*>(out + r * dim + 4 * v) = o;
    }
  }
}
template <int FD>
__global__ void __launch_bounds__(256) masked_mean_bwd_kernel(const float* __restrict__ d_out,
                                                              const int64_t* __restrict__ mask, int64_t n_seq,
                                                              int64_t T, int64_t dim, void* __restrict__ d_feats) {
  const RowMap m = row_map(dim);
  const int lane = threadIdx.x & 31;
  const int grp = lane / m.lpr, gl = lane % m.lpr;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r0 = warp * m.rpw; r0 < n_seq; r0 += nwarps * m.rpw) {
    const int64_t r = r0 + grp;
    if (r >= n_seq) continue;
    float cnt = 0.f;
    for (int64_t t = 0; t < T; ++t) cnt += (float)__ldg(mask + r * T + t);
    const float inv = 1.0f / fmaxf(cnt, 1e-9f);
    for (int v = gl; v < m.vecs; v += m.lpr) {
      const float4 g = ldg_f4(d_out + r * dim + 4 * v);
      for (int64_t t = 0; t < T; ++t) {
        const float s = (float)__ldg(mask + r * T + t) * inv;
        store4<FD>(d_feats, (r * T + t) * dim + 4 * v, make_float4(g.x * s, g.y * s, g.z * s, g.w * s));
      }
    }
  }
}

// ------------------------------------------------------------------ static front (U2)
struct StaticParams {
  const int64_t* ids[9];
  const float* tables[9];
  int64_t rows[9];
};
__constant__ int kStaticDim[9] = {16, 16, 16, 16, 4, 4, 4, 4, 4};
__constant__ int kStaticCol[10] = {0, 16, 32, 48, 64, 68, 72, 76, 80, 84};
__constant__ int kStaticSlot[10] = {0, 176, 352, 528, 704, 720, 736, 748, 760, 772};  // accumulator layout (bwd)

__device__ __forceinline__ void static_col_to_field(int c, int& f, int& d) {
  if (c < 64) { f = c >> 4; d = c & 15; }
  else if (c < 84) { f = 4 + ((c - 64) >> 2); d = (c - 64) & 3; }
  else { f = 9; d = c - 84; }
}

__global__ void __launch_bounds__(256) static_front_fwd_kernel(StaticParams prm, const float* __restrict__ cont,
                                                               const float* __restrict__ cont_w,
                                                               const float* __restrict__ cont_b,
                                                               const float* __restrict__ gates, int64_t B,
                                                               float* __restrict__ out, int* __restrict__ oob) {
  const int64_t total = B * 100;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / 100;
    const int c = (int)(i % 100);
    int f, d;
    static_col_to_field(c, f, d);
    const float g = __ldg(gates + f);
    float v;
    if (f < 9) {
      int64_t id = __ldg(prm.ids[f] + b);
      if (id < 0 || id >= prm.rows[f]) { if (oob) *oob = 1; v = 0.f; }
      else v = __fmul_rn(__ldg(prm.tables[f] + id * kStaticDim[f] + d), g);
    } else {
      float z = __ldg(cont_b + d);
#pragma unroll
      for (int k = 0; k < 4; ++k) z = fmaf(__ldg(cont + b * 4 + k), __ldg(cont_w + d * 4 + k), z);
      v = __fmul_rn(fmaxf(z, 0.f), g);
    }
    out[i] = v;
  }
}

// backward accumulators, one owner thread each (no atomics, fixed summation order):
//   [0,772)  table slots   (field f, row r, dim d) at kStaticSlot[f] + r*dim_f + d
//   [772,782) d_gates[10]   [782,846) d_cont_w[16*4]   [846,862) d_cont_b[16]
#define RS_STATIC_ACC 862
#define RS_STATIC_CHUNK 64
__global__ void __launch_bounds__(896) static_front_bwd_kernel(const float* __restrict__ d_out, StaticParams prm,
                                                               const float* __restrict__ cont,
                                                               const float* __restrict__ cont_w,
                                                               const float* __restrict__ cont_b,
                                                               const float* __restrict__ gates, int64_t B,
                                                               int64_t padding_idx, float* __restrict__ partial) {
  __shared__ float s_g[RS_STATIC_CHUNK][100];
  __shared__ int s_id[RS_STATIC_CHUNK][9];
  __shared__ float s_cont[RS_STATIC_CHUNK][4];
  __shared__ float s_z[RS_STATIC_CHUNK][16];
  const int tid = threadIdx.x;
  // decode which accumulator this thread owns
  int kind = -1, f = 0, r = 0, d = 0;
  if (tid < 772) {
    kind = 0;
    f = 8;
    for (int i = 0; i < 9; ++i) if (tid < kStaticSlot[i + 1]) { f = i; break; }
    const int o = tid - kStaticSlot[f];
    r = o / kStaticDim[f]; d = o % kStaticDim[f];
  } else if (tid < 782) { kind = 1; f = tid - 772; }
  else if (tid < 846) { kind = 2; d = (tid - 782) >> 2; r = (tid - 782) & 3; }
  else if (tid < 862) { kind = 3; d = tid - 846; }
  const float g9 = __ldg(gates + 9);
  const float gf = (kind == 0) ? __ldg(gates + f) : 0.f;
  float acc = 0.f;
  for (int64_t c0 = (int64_t)blockIdx.x * RS_STATIC_CHUNK; c0 < B; c0 += (int64_t)gridDim.x * RS_STATIC_CHUNK) {
    const int cn = (int)((B - c0) < RS_STATIC_CHUNK ? (B - c0) : RS_STATIC_CHUNK);
    __syncthreads();
    for (int i = tid; i < cn * 100; i += blockDim.x) s_g[i / 100][i % 100] = d_out[(c0 + i / 100) * 100 + i % 100];
    for (int i = tid; i < cn * 9; i += blockDim.x) s_id[i / 9][i % 9] = (int)__ldg(prm.ids[i % 9] + c0 + i / 9);
    for (int i = tid; i < cn * 4; i += blockDim.x) s_cont[i / 4][i % 4] = cont[(c0 + i / 4) * 4 + i % 4];
    __syncthreads();
    for (int i = tid; i < cn * 16; i += blockDim.x) {
      const int s = i / 16, j = i % 16;
      float z = __ldg(cont_b + j);
#pragma unroll
      for (int k = 0; k < 4; ++k) z = fmaf(s_cont[s][k], __ldg(cont_w + j * 4 + k), z);
      s_z[s][j] = z;
    }
    __syncthreads();
    if (kind == 0) {
      const int col = kStaticCol[f] + d;
      if ((int64_t)r != padding_idx)
        for (int s = 0; s < cn; ++s) if (s_id[s][f] == r) acc = fmaf(gf, s_g[s][col], acc);
    } else if (kind == 1) {
      if (f < 9) {
        const int dimf = kStaticDim[f], col = kStaticCol[f];
        for (int s = 0; s < cn; ++s) {
          const int id = s_id[s][f];
          if (id >= 0 && id < prm.rows[f])
            for (int k = 0; k < dimf; ++k) acc = fmaf(__ldg(prm.tables[f] + id * dimf + k), s_g[s][col + k], acc);
        }
      } else {
        for (int s = 0; s < cn; ++s)
          for (int j = 0; j < 16; ++j) acc = fmaf(fmaxf(s_z[s][j], 0.f), s_g[s][84 + j], acc);
      }
    } else if (kind == 2) {
      for (int s = 0; s < cn; ++s) if (s_z[s][d] > 0.f) acc = fmaf(g9 * s_g[s][84 + d], s_cont[s][r], acc);
    } else if (kind == 3) {
      for (int s = 0; s < cn; ++s) if (s_z[s][d] > 0.f) acc = fmaf(g9, s_g[s][84 + d], acc);
    }
  }
  if (tid < RS_STATIC_ACC) partial[(int64_t)blockIdx.x * RS_STATIC_ACC + tid] = acc;
}

struct StaticGradPtrs { float* t[9]; };
__global__ void static_front_bwd_finalize(const float* __restrict__ partial, int nparts, StaticGradPtrs gp,
                                          float* __restrict__ d_gates, float* __restrict__ d_w,
                                          float* __restrict__ d_b) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= RS_STATIC_ACC) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(int64_t)p * RS_STATIC_ACC + tid];
  if (tid < 772) {
    int f = 8;
    for (int i = 0; i < 9; ++i) if (tid < kStaticSlot[i + 1]) { f = i; break; }
    if (gp.t[f]) gp.t[f][tid - kStaticSlot[f]] = s;
  } else if (tid < 782) d_gates[tid - 772] = s;
  else if (tid < 846) d_w[tid - 782] = s;
  else d_b[tid - 846] = s;
}

}  // namespace rs

// =============================================================================================
// C ABI
// =============================================================================================
using namespace rs;

#define DISPATCH_DT(dt, NAME, ...)                         \
  switch (dt) {                                            \
    case RS_F32: { constexpr int NAME = RS_F32; __VA_ARGS__; break; }   \
    case RS_F16: { constexpr int NAME = RS_F16; __VA_ARGS__; break; }   \
    case RS_BF16: { constexpr int NAME = RS_BF16; __VA_ARGS__; break; } \
    default: return RS_ERR_BAD_ARG;                        \
  }

static inline bool dim_ok(int64_t dim, int dt_a, int dt_b) {
  if (dim <= 0 || (dim & 3)) return false;
  (void)dt_a; (void)dt_b;
  return true;
}

#include <atomic>
static std::atomic<unsigned long long> g_launches{0};
extern "C" void rs_count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
extern "C" unsigned long long rs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int rs_abi_version(void) { return 1; }

extern "C" const char* rs_error_string(int code) {
  switch (code) {
    case RS_OK: return "ok";
    case RS_ERR_BAD_ARG: return "rs_twotower: bad argument (null pointer, dim not a multiple of 4, unknown dtype, ...)";
    case RS_ERR_UNSUPPORTED: return "rs_twotower: unsupported shape for this kernel";
    case RS_ERR_WORKSPACE: return "rs_twotower: workspace too small";
    default: return cudaGetErrorString((cudaError_t)code);
  }
}

extern "C" int rs_gather_rows(const void* table, int table_dtype, int64_t rows, int64_t dim, const int64_t* ids,
                              int64_t n, int64_t clamp_max, void* out, int out_dtype, int* oob_flag, void* stream) {
  if (n == 0) return RS_OK;
  if (!table || !ids || !out || !dim_ok(dim, table_dtype, out_dtype)) return RS_ERR_BAD_ARG;
  const RowMap m = row_map(dim);
  const int grid = grid_for_warps((n + m.rpw - 1) / m.rpw, 8, 8);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(table_dtype, TD, DISPATCH_DT(out_dtype, OD, (gather_rows_kernel<TD, OD><<<grid, 256, 0, st>>>(
      table, rows, dim, ids, n, clamp_max, out, oob_flag))));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_scatter_add_rows(const void* d_out, int d_out_dtype, const int64_t* ids, int64_t n, int64_t dim,
                                   int64_t rows, int64_t padding_idx, int64_t clamp_max, float scale, float* d_table,
                                   int* oob_flag, void* stream) {
  if (n == 0) return RS_OK;
  if (!d_out || !ids || !d_table || !dim_ok(dim, d_out_dtype, 0)) return RS_ERR_BAD_ARG;
  const RowMap m = row_map(dim);
  const int grid = grid_for_warps((n + m.rpw - 1) / m.rpw, 8, 8);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(d_out_dtype, GD, (scatter_add_rows_kernel<GD><<<grid, 256, 0, st>>>(
      d_out, ids, n, dim, rows, padding_idx, clamp_max, scale, d_table, oob_flag)));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_seq_front_fwd(const void* base, int base_dtype, const int64_t* const* ids,
                                const float* const* tables, const int64_t* table_rows, int n_tables,
                                const float* gates, const float* pos_table, int64_t L, int64_t P, int64_t dim,
                                void* out, int out_dtype, int* oob_flag, void* stream) {
  if (P == 0) return RS_OK;
  if (!out || n_tables < 0 || n_tables > RS_MAX_TABLES || (n_tables && !gates) || !dim_ok(dim, 0, 0) || L <= 0)
    return RS_ERR_BAD_ARG;
  SeqFrontParams prm;
  prm.n_tables = n_tables;
  for (int t = 0; t < RS_MAX_TABLES; ++t) {
    prm.ids[t] = t < n_tables ? ids[t] : nullptr;
    prm.tables[t] = t < n_tables ? tables[t] : nullptr;
    prm.rows[t] = t < n_tables ? table_rows[t] : 0;
    if (t < n_tables && (!prm.ids[t] || !prm.tables[t])) return RS_ERR_BAD_ARG;
  }
  const RowMap m = row_map(dim);
  const int grid = grid_for_warps((P + m.rpw - 1) / m.rpw, 8, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (!base) base_dtype = RS_F32;
  if (dim == 128 && n_tables >= 1 && n_tables <= 3 && L < (1ll << 30)) {
    const int grid128 = grid_for_warps((P + 31) / 32, 8, 3);
#define LAUNCH_SF(NT)                                                                                        \
  DISPATCH_DT(base_dtype, BD, DISPATCH_DT(out_dtype, OD, (seq_front_fwd128_kernel<BD, OD, NT><<<grid128, 256, 0, st>>>( \
      base, prm, gates, pos_table, (int)L, P, out, oob_flag))))
    if (n_tables == 1) { LAUNCH_SF(1); } else if (n_tables == 2) { LAUNCH_SF(2); } else { LAUNCH_SF(3); }
    RS_LAUNCH_CHECK();
    return RS_OK;
  }
  DISPATCH_DT(base_dtype, BD, DISPATCH_DT(out_dtype, OD, (seq_front_fwd_kernel<BD, OD><<<grid, 256, 0, st>>>(
      base, prm, gates, pos_table, L, P, dim, out, oob_flag))));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_normalized_rows_fwd(const float* table, int64_t rows, int64_t dim, const int64_t* ids, int64_t n,
                                      float eps, void* out, int out_dtype, float* inv_norm, int* oob_flag,
                                      void* stream) {
  if (n == 0) return RS_OK;
  if (!table || !ids || !out || !dim_ok(dim, 0, 0)) return RS_ERR_BAD_ARG;
  const RowMap m = row_map(dim);
  const int grid = grid_for_warps((n + m.rpw - 1) / m.rpw, 8, 8);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(out_dtype, OD, (normalized_rows_fwd_kernel<OD><<<grid, 256, 0, st>>>(table, rows, dim, ids, n, eps, out,
                                                                                   inv_norm, oob_flag)));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_normalized_rows_bwd(const void* d_out, int d_out_dtype, const float* table, int64_t rows,
                                      int64_t dim, const int64_t* ids, int64_t n, float eps, float* d_table,
                                      void* stream) {
  if (n == 0) return RS_OK;
  if (!d_out || !table || !ids || !d_table || !dim_ok(dim, 0, 0)) return RS_ERR_BAD_ARG;
  const RowMap m = row_map(dim);
  const int grid = grid_for_warps((n + m.rpw - 1) / m.rpw, 8, 8);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(d_out_dtype, GD, (normalized_rows_bwd_kernel<GD><<<grid, 256, 0, st>>>(d_out, table, rows, dim, ids, n,
                                                                                     eps, d_table)));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

template <int OD>
static int launch_ln_front(const float* table, int64_t rows, int64_t dim, const int64_t* ids, int64_t n,
                           const float* addB, int64_t periodB, const float* addC, const float* ln_w,
                           const float* ln_b, float eps, float p, uint64_t seed, void* out, float* mean, float* rstd,
                           int* oob, cudaStream_t st) {
  const RowMap m = row_map(dim);
  const int grid = grid_for_warps((n + m.rpw - 1) / m.rpw, 8, 8);
  const int need = (m.vecs + m.lpr - 1) / m.lpr;
  if (need <= 1) ln_front_kernel<OD, 1><<<grid, 256, 0, st>>>(table, rows, dim, ids, n, addB, periodB, addC, ln_w, ln_b, eps, p, seed, out, mean, rstd, oob);
  else if (need <= 2) ln_front_kernel<OD, 2><<<grid, 256, 0, st>>>(table, rows, dim, ids, n, addB, periodB, addC, ln_w, ln_b, eps, p, seed, out, mean, rstd, oob);
  else if (need <= 8) ln_front_kernel<OD, 8><<<grid, 256, 0, st>>>(table, rows, dim, ids, n, addB, periodB, addC, ln_w, ln_b, eps, p, seed, out, mean, rstd, oob);
  else return RS_ERR_UNSUPPORTED;
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_std_front_fwd(const float* table, int64_t rows, int64_t dim, const int64_t* ids, int64_t n,
                                const float* field_emb, int64_t n_fields, const float* ln_w, const float* ln_b,
                                float eps, void* out, int out_dtype, float* mean, float* rstd, int* oob_flag,
                                void* stream) {
  if (n == 0) return RS_OK;
  if (!table || !ids || !out || !ln_w || !ln_b || !dim_ok(dim, 0, 0)) return RS_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(out_dtype, OD, return launch_ln_front<OD>(table, rows, dim, ids, n, field_emb, n_fields, nullptr, ln_w,
                                                        ln_b, eps, 0.f, 0, out, mean, rstd, oob_flag, st));
  return RS_OK;
}

extern "C" int rs_bert_embed_fwd(const float* word, int64_t vocab, const float* pos, const float* type0,
                                 const float* ln_w, const float* ln_b, float eps, const int64_t* ids, int64_t n_seq,
                                 int64_t T, int64_t dim, float dropout_p, uint64_t seed, void* out, int out_dtype,
                                 int* oob_flag, void* stream) {
  if (n_seq == 0 || T == 0) return RS_OK;
  if (!word || !pos || !type0 || !ids || !out || !ln_w || !ln_b || !dim_ok(dim, 0, 0) || dropout_p < 0.f ||
      dropout_p >= 1.f)
    return RS_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(out_dtype, OD, return launch_ln_front<OD>(word, vocab, dim, ids, n_seq * T, pos, T, type0, ln_w, ln_b,
                                                        eps, dropout_p, seed, out, nullptr, nullptr, oob_flag, st));
  return RS_OK;
}

extern "C" int rs_masked_mean_fwd(const void* feats, int feats_dtype, const int64_t* mask, int64_t n_seq, int64_t T,
                                  int64_t dim, float* out, void* stream) {
  if (n_seq == 0) return RS_OK;
  if (!feats || !mask || !out || !dim_ok(dim, 0, 0)) return RS_ERR_BAD_ARG;
  const RowMap m = row_map(dim);
  const int grid = grid_for_warps((n_seq + m.rpw - 1) / m.rpw, 8, 8);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(feats_dtype, FD, (masked_mean_fwd_kernel<FD><<<grid, 256, 0, st>>>(feats, mask, n_seq, T, dim, out)));
  RS_LAUNCH_CHECK();
  return RS_OK;
}
extern "C" int rs_masked_mean_bwd(const float* d_out, const int64_t* mask, int64_t n_seq, int64_t T, int64_t dim,
                                  void* d_feats, int d_feats_dtype, void* stream) {
  if (n_seq == 0) return RS_OK;
  if (!d_out || !mask || !d_feats || !dim_ok(dim, 0, 0)) return RS_ERR_BAD_ARG;
  const RowMap m = row_map(dim);
  const int grid = grid_for_warps((n_seq + m.rpw - 1) / m.rpw, 8, 8);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(d_feats_dtype, FD, (masked_mean_bwd_kernel<FD><<<grid, 256, 0, st>>>(d_out, mask, n_seq, T, dim, d_feats)));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_static_front_fwd(const int64_t* const* ids, const float* const* tables, const int64_t* table_rows,
                                   const float* cont, const float* cont_w, const float* cont_b, const float* gates,
                                   int64_t B, float* out, int* oob_flag, void* stream) {
  if (B == 0) return RS_OK;
  if (!ids || !tables || !table_rows || !cont || !cont_w || !cont_b || !gates || !out) return RS_ERR_BAD_ARG;
  StaticParams prm;
  for (int i = 0; i < 9; ++i) { prm.ids[i] = ids[i]; prm.tables[i] = tables[i]; prm.rows[i] = table_rows[i]; }
  int64_t blocks = (B * 100 + 255) / 256;
  if (blocks > RS_NUM_SMS * 8) blocks = RS_NUM_SMS * 8;
  static_front_fwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(prm, cont, cont_w, cont_b, gates, B, out,
                                                                         oob_flag);
  RS_LAUNCH_CHECK();
  return RS_OK;
}

static int static_bwd_grid(int64_t B) {
  int64_t g = (B + RS_STATIC_CHUNK - 1) / RS_STATIC_CHUNK;
  return (int)(g < RS_NUM_SMS ? (g < 1 ? 1 : g) : RS_NUM_SMS);
}
extern "C" size_t rs_static_front_bwd_workspace_bytes(int64_t B) {
  return (size_t)static_bwd_grid(B) * RS_STATIC_ACC * sizeof(float);
}
extern "C" int rs_static_front_bwd(const float* d_out, const int64_t* const* ids, const float* const* tables,
                                   const int64_t* table_rows, const float* cont, const float* cont_w,
                                   const float* cont_b, const float* gates, int64_t B, int64_t padding_idx,
                                   float* const* d_tables, float* d_gates, float* d_cont_w, float* d_cont_b,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  if (!d_out || !ids || !tables || !d_tables || !d_gates || !d_cont_w || !d_cont_b || !workspace) return RS_ERR_BAD_ARG;
  if (workspace_bytes < rs_static_front_bwd_workspace_bytes(B)) return RS_ERR_WORKSPACE;
  StaticParams prm;
  StaticGradPtrs gp;
  for (int i = 0; i < 9; ++i) {
    prm.ids[i] = ids[i]; prm.tables[i] = tables[i]; prm.rows[i] = table_rows[i]; gp.t[i] = d_tables[i];
  }
  const int grid = static_bwd_grid(B);
  cudaStream_t st = (cudaStream_t)stream;
  static_front_bwd_kernel<<<grid, 896, 0, st>>>(d_out, prm, cont, cont_w, cont_b, gates, B, padding_idx,
                                                (float*)workspace);
  RS_LAUNCH_CHECK();
  static_front_bwd_finalize<<<(RS_STATIC_ACC + 127) / 128, 128, 0, st>>>((const float*)workspace, grid, gp, d_gates,
                                                                         d_cont_w, d_cont_b);
  RS_LAUNCH_CHECK();
  return RS_OK;
}
