// Sparse backward of the embedding fronts: scatter-add of gradient rows into dense fp32 tables.
//   * rs_seq_front_bwd      one sequential pass over dX: positional column sums in registers, small
//                           tables in per-warp private shared-memory copies (no atomics), wide tables
//                           by 128-bit vector atomics (mode 1) or left to the sorted path (mode 0)
//   * rs_sort_ids           stable LSD radix sort of (id, position), 9 bits per pass
//   * rs_segment_reduce_rows  deterministic segmented sum over the sorted order (tile partials + fix-up)
#include "common.cuh"
#include "../../include/rs_twotower.h"

namespace rs {

// =============================================================================================
// seq_front backward: dense pass
// =============================================================================================
#define BWD_WARPS 8
#define BWD_BC 64          // batch rows per (l, chunk) work item
#define BWD_UNROLL 4

struct SeqBwdParams {
  const int64_t* ids[RS_MAX_TABLES];
  const float* tables[RS_MAX_TABLES];
  float* d_tables[RS_MAX_TABLES];
  int64_t rows[RS_MAX_TABLES];
  int mode[RS_MAX_TABLES];
  int small_off[RS_MAX_TABLES];   // row offset of a mode-2 table inside the private copy
  int n_tables;
  int small_rows;                 // total rows of all mode-2 tables
};

template <int GD, int NV>
__global__ void __launch_bounds__(BWD_WARPS * 32) seq_front_bwd_kernel(
    const void* __restrict__ dx, SeqBwdParams prm, const float* __restrict__ gates, int64_t L, int64_t B,
    int64_t dim, int64_t padding_idx, int nchunks, float* __restrict__ pos_part, float* __restrict__ small_part,
    float* __restrict__ gate_part) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int vecs = (int)(dim >> 2);
  const int small_elems = prm.small_rows * (int)dim;
  float* priv = smem + (size_t)wid * small_elems;            // this warp's private copy of the small tables
  for (int i = lane; i < small_elems; i += 32) priv[i] = 0.f;
  __syncwarp();

  float g[RS_MAX_TABLES], dotacc[RS_MAX_TABLES];
#pragma unroll
  for (int t = 0; t < RS_MAX_TABLES; ++t) {
    g[t] = (t < prm.n_tables) ? __ldg(gates + t) : 0.f;
    dotacc[t] = 0.f;
  }
  const int64_t n_items = L * nchunks;
  const int64_t gw = (int64_t)blockIdx.x * BWD_WARPS + wid, nw = (int64_t)gridDim.x * BWD_WARPS;
  for (int64_t item = gw; item < n_items; item += nw) {
    const int64_t l = item % L;
    const int c = (int)(item / L);
    const int64_t b0 = (int64_t)c * BWD_BC;
    const int64_t b1 = (b0 + BWD_BC < B) ? b0 + BWD_BC : B;
    float4 pacc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) pacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t bb = b0; bb < b1; bb += BWD_UNROLL) {
      float4 d[BWD_UNROLL][NV];
      int64_t id[BWD_UNROLL][RS_MAX_TABLES];
#pragma unroll
      for (int u = 0; u < BWD_UNROLL; ++u) {
        const int64_t p = (bb + u) * L + l;
        const bool live = (bb + u) < b1;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = lane + 32 * i;
          d[u][i] = (live && v < vecs) ? load4<GD>(dx, p * dim + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int t = 0; t < RS_MAX_TABLES; ++t) {
          id[u][t] = -1;
          if (live && t < prm.n_tables && prm.mode[t] != 0) {
            const int64_t x = __ldg(prm.ids[t] + p);
            id[u][t] = (x >= 0 && x < prm.rows[t]) ? x : -1;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < BWD_UNROLL; ++u) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = lane + 32 * i;
          if (v >= vecs) continue;
          pacc[i].x += d[u][i].x; pacc[i].y += d[u][i].y; pacc[i].z += d[u][i].z; pacc[i].w += d[u][i].w;
#pragma unroll
          for (int t = 0; t < RS_MAX_TABLES; ++t) {
            const int64_t x = id[u][t];
            if (x < 0) continue;
            const float4 e = ldg_f4(prm.tables[t] + x * dim + 4 * v);
            dotacc[t] += dot4(e, d[u][i]);
            if (x == padding_idx) continue;
            if (prm.mode[t] == 1) {
              red_add_f4(prm.d_tables[t] + x * dim + 4 * v,
                         make_float4(g[t] * d[u][i].x, g[t] * d[u][i].y, g[t] * d[u][i].z, g[t] * d[u][i].w));
            } else {
              float4* q = reinterpret_cast<float4*>(priv + ((size_t)(prm.small_off[t] + x)) * dim + 4 * v);
              float4 a = *q;
              *q = fma4(a, d[u][i], g[t]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < vecs) *reinterpret_cast<float4*>(pos_part + ((size_t)c * L + l) * dim + 4 * v) = pacc[i];
    }
  }
  // ---- CTA-level combine, fixed order over warps -> deterministic
  __shared__ float s_dot[BWD_WARPS][RS_MAX_TABLES];
#pragma unroll
  for (int t = 0; t < RS_MAX_TABLES; ++t) {
    const float s = warp_sum(dotacc[t]);
    if (lane == 0) s_dot[wid][t] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < small_elems; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < BWD_WARPS; ++w) s += smem[(size_t)w * small_elems + i];
    small_part[(size_t)blockIdx.x * small_elems + i] = s;
  }
  if (threadIdx.x < RS_MAX_TABLES) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < BWD_WARPS; ++w) s += s_dot[w][threadIdx.x];
    gate_part[(size_t)blockIdx.x * RS_MAX_TABLES + threadIdx.x] = s;
  }
}

// dim == 128, L <= 64 fast path.  A CTA streams whole batch rows (L consecutive positions = one contiguous
// slab of dX); warp w owns the time steps l = w, w+8, ... so the positional sums stay in registers, and it
// keeps up to 8 independent row loads in flight per batch row.
#define BWD_LPW 8          // time steps per warp (L <= 8 * BWD_LPW)
template <int GD>
__global__ void __launch_bounds__(BWD_WARPS * 32, 2) seq_front_bwd128_kernel(
    const void* __restrict__ dx, SeqBwdParams prm, const float* __restrict__ gates, int L, int64_t B,
    int64_t padding_idx, float* __restrict__ pos_part, float* __restrict__ small_part,
    float* __restrict__ gate_part) {
  constexpr int D = 128;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int small_elems = prm.small_rows * D;
  float* priv = smem + (size_t)wid * small_elems;
  for (int i = lane; i < small_elems; i += 32) priv[i] = 0.f;
  __syncwarp();
  float g[RS_MAX_TABLES], dotacc[RS_MAX_TABLES];
#pragma unroll
  for (int t = 0; t < RS_MAX_TABLES; ++t) {
    g[t] = (t < prm.n_tables) ? __ldg(gates + t) : 0.f;
    dotacc[t] = 0.f;
  }
  float4 pacc[BWD_LPW];
#pragma unroll
  for (int i = 0; i < BWD_LPW; ++i) pacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    float4 d[BWD_LPW];
#pragma unroll
    for (int i = 0; i < BWD_LPW; ++i) {
      const int l = wid + BWD_WARPS * i;
      d[i] = (l < L) ? load4<GD>(dx, (b * L + l) * D + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int t = 0; t < RS_MAX_TABLES; ++t) {
      if (t >= prm.n_tables || prm.mode[t] == 0) continue;       // warp-uniform
      int id[BWD_LPW];
      float4 e[BWD_LPW];
#pragma unroll
      for (int i = 0; i < BWD_LPW; ++i) {
        const int l = wid + BWD_WARPS * i;
        id[i] = -1;
        if (l < L) {
          const int64_t x = __ldg(prm.ids[t] + b * L + l);
          if (x >= 0 && x < prm.rows[t]) id[i] = (int)x;
        }
        if (id[i] >= 0) e[i] = ldg_f4(prm.tables[t] + (int64_t)id[i] * D + 4 * lane);
      }
#pragma unroll
      for (int i = 0; i < BWD_LPW; ++i) {
        if (id[i] < 0) continue;
        dotacc[t] += dot4(e[i], d[i]);
        if (id[i] == padding_idx) continue;
        if (prm.mode[t] == 1) {
          red_add_f4(prm.d_tables[t] + (int64_t)id[i] * D + 4 * lane,
                     make_float4(g[t] * d[i].x, g[t] * d[i].y, g[t] * d[i].z, g[t] * d[i].w));
        } else {
          float4* q = reinterpret_cast<float4*>(priv + ((size_t)(prm.small_off[t] + id[i])) * D + 4 * lane);
          *q = fma4(*q, d[i], g[t]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < BWD_LPW; ++i) {
      pacc[i].x += d[i].x; pacc[i].y += d[i].y; pacc[i].z += d[i].z; pacc[i].w += d[i].w;
    }
  }
#pragma unroll
  for (int i = 0; i < BWD_LPW; ++i) {
    const int l = wid + BWD_WARPS * i;
    if (l < L) *reinterpret_cast<float4*>(pos_part + ((size_t)blockIdx.x * L + l) * D + 4 * lane) = pacc[i];
  }
  __shared__ float s_dot[BWD_WARPS][RS_MAX_TABLES];
#pragma unroll
  for (int t = 0; t < RS_MAX_TABLES; ++t) {
    const float s = warp_sum(dotacc[t]);
    if (lane == 0) s_dot[wid][t] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < small_elems; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < BWD_WARPS; ++w) s += smem[(size_t)w * small_elems + i];
    small_part[(size_t)blockIdx.x * small_elems + i] = s;
  }
  if (threadIdx.x < RS_MAX_TABLES) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < BWD_WARPS; ++w) s += s_dot[w][threadIdx.x];
    gate_part[(size_t)blockIdx.x * RS_MAX_TABLES + threadIdx.x] = s;
  }
}

__global__ void seq_front_bwd_finalize(SeqBwdParams prm, int64_t L, int64_t dim, int nchunks, int nctas,
                                       const float* __restrict__ pos_part, const float* __restrict__ small_part,
                                       const float* __restrict__ gate_part, float* __restrict__ d_pos,
                                       float* __restrict__ d_gates) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_pos = L * dim;
  const int64_t n_small = (int64_t)prm.small_rows * dim;
  if (tid < n_pos) {
    if (d_pos) {
      float s = 0.f;
      for (int c = 0; c < nchunks; ++c) s += pos_part[(size_t)c * n_pos + tid];
      d_pos[tid] = s;
    }
  } else if (tid < n_pos + n_small) {
    const int64_t i = tid - n_pos;
    float s = 0.f;
    for (int c = 0; c < nctas; ++c) s += small_part[(size_t)c * n_small + i];
    const int64_t row = i / dim;
    for (int t = 0; t < prm.n_tables; ++t)
      if (prm.mode[t] == 2 && row >= prm.small_off[t] && row < prm.small_off[t] + prm.rows[t])
        prm.d_tables[t][(row - prm.small_off[t]) * dim + (i % dim)] = s;
  } else if (tid < n_pos + n_small + RS_MAX_TABLES) {
    const int t = (int)(tid - n_pos - n_small);
    if (t < prm.n_tables) {
      float s = 0.f;
      if (prm.mode[t] != 0)
        for (int c = 0; c < nctas; ++c) s += gate_part[(size_t)c * RS_MAX_TABLES + t];
      d_gates[t] = s;
    }
  }
}

// =============================================================================================
// stable LSD radix sort of (id, position); ids < 2^31.  9-bit digits: the 105,543-row item table (17-bit keys) sorts
// in two passes.  Per pass: per-tile digit histogram -> one CTA per digit scans its row of the [digit][tile] table
// (the single-CTA scan of round 1 was 45 us of every 60 us pass, profiles/r02d) -> stable scatter, which adds the
// digit bases from a 512-entry block scan of the digit totals.
// =============================================================================================
#define SORT_THREADS 256
#define SORT_WARPS 8
#define SORT_ITEMS 8                       // rounds of 32 per warp
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)
#define SORT_BITS 9
#define SORT_DIGITS (1 << SORT_BITS)
#define SORT_DPT (SORT_DIGITS / SORT_THREADS)   // digits per thread

// first pass reads the int64 ids (clamp + range check); later passes read int32 keys
__device__ __forceinline__ int sort_key_first(const int64_t* ids, int64_t i, int64_t rows, int64_t clamp_max,
                                              int* oob) {
  int64_t x = __ldg(ids + i);
  if (clamp_max >= 0 && x > clamp_max) x = clamp_max;
  if (x < 0 || x >= rows) { if (oob && x != -1) *oob = 1; x = rows; }   // out-of-range ids sort last (key == rows); -1 = null id
  return (int)x;
}

template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(const int64_t* __restrict__ ids,
                                                                 const int* __restrict__ keys_in, int64_t n,
                                                                 int64_t rows, int64_t clamp_max, int shift,
                                                                 int ntiles, unsigned* __restrict__ hist,
                                                                 int* __restrict__ oob) {
  __shared__ unsigned h[SORT_DIGITS];
#pragma unroll
  for (int q = 0; q < SORT_DPT; ++q) h[threadIdx.x + q * SORT_THREADS] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
  for (int k = 0; k < SORT_ITEMS; ++k) {
    const int64_t i = base + k * SORT_THREADS + threadIdx.x;
    if (i < n) {
      const int key = FIRST ? sort_key_first(ids, i, rows, clamp_max, oob) : keys_in[i];
      atomicAdd(&h[(key >> shift) & (SORT_DIGITS - 1)], 1u);          // integer counts: order-independent
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < SORT_DPT; ++q) {
    const int d = threadIdx.x + q * SORT_THREADS;
    hist[(size_t)d * ntiles + blockIdx.x] = h[d];
  }
}

// CTA d: exclusive scan of hist[d][0..ntiles) in place, total -> totals[d]
__global__ void __launch_bounds__(SORT_THREADS) sort_scan_kernel(unsigned* __restrict__ hist, int ntiles,
                                                                 unsigned* __restrict__ totals) {
  __shared__ unsigned s_warp[SORT_WARPS];
  __shared__ unsigned s_carry;
  unsigned* row = hist + (size_t)blockIdx.x * ntiles;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int c0 = 0; c0 < ntiles; c0 += SORT_THREADS) {
    const int i = c0 + threadIdx.x;
    const unsigned v = i < ntiles ? row[i] : 0u;
    unsigned x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    unsigned before = s_carry;
#pragma unroll
    for (int ww = 0; ww < SORT_WARPS; ++ww) if (ww < w) before += s_warp[ww];
    if (i < ntiles) row[i] = before + x - v;
    __syncthreads();
    if (threadIdx.x == SORT_THREADS - 1) s_carry = before + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = s_carry;
}

template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(const int64_t* __restrict__ ids,
                                                                    const int* __restrict__ keys_in,
                                                                    const int* __restrict__ vals_in, int64_t n,
                                                                    int64_t rows, int64_t clamp_max, int shift,
                                                                    int ntiles, const unsigned* __restrict__ offs,
                                                                    const unsigned* __restrict__ totals,
                                                                    int* __restrict__ keys_out,
                                                                    int* __restrict__ vals_out) {
  __shared__ unsigned wcount[SORT_WARPS][SORT_DIGITS];     // phase A: per-warp digit counts; phase B: running bases
  __shared__ unsigned dbase[SORT_DIGITS];                   // exclusive scan of the digit totals
  __shared__ unsigned s_warp[SORT_WARPS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < SORT_WARPS * SORT_DIGITS; i += SORT_THREADS) (&wcount[0][0])[i] = 0;
  {
    // thread t owns digits [t*DPT, (t+1)*DPT): local sums, block scan, write back
    unsigned loc[SORT_DPT], sum = 0;
#pragma unroll
    for (int q = 0; q < SORT_DPT; ++q) { loc[q] = totals[threadIdx.x * SORT_DPT + q]; sum += loc[q]; }
    unsigned x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    unsigned run = x - sum;
#pragma unroll
    for (int ww = 0; ww < SORT_WARPS; ++ww) if (ww < w) run += s_warp[ww];
#pragma unroll
    for (int q = 0; q < SORT_DPT; ++q) { dbase[threadIdx.x * SORT_DPT + q] = run; run += loc[q]; }
  }
  __syncthreads();
  // warp w owns the contiguous sub-range [base + w*256, base + (w+1)*256) of the tile, walked in order
  const int64_t wbase = (int64_t)blockIdx.x * SORT_TILE + (int64_t)w * (32 * SORT_ITEMS);
  int key[SORT_ITEMS], val[SORT_ITEMS];
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; ++k) {
    const int64_t i = wbase + k * 32 + lane;
    if (i < n) {
      key[k] = FIRST ? sort_key_first(ids, i, rows, clamp_max, nullptr) : keys_in[i];
      val[k] = FIRST ? (int)i : vals_in[i];
    } else { key[k] = -1; val[k] = 0; }
  }
  // phase A: count
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; ++k) {
    const bool live = key[k] >= 0;
    const int dg = live ? ((key[k] >> shift) & (SORT_DIGITS - 1)) : SORT_DIGITS + lane;   // dead lanes: unique pseudo-digit
    const unsigned peers = __match_any_sync(0xffffffffu, dg);
    if (live && (peers & ((1u << lane) - 1)) == 0) wcount[w][dg] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // exclusive prefix over warps per digit + global offset of (digit, tile)
#pragma unroll
  for (int q = 0; q < SORT_DPT; ++q) {
    const int dg = threadIdx.x + q * SORT_THREADS;
    unsigned run = dbase[dg] + offs[(size_t)dg * ntiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < SORT_WARPS; ++ww) { unsigned c = wcount[ww][dg]; wcount[ww][dg] = run; run += c; }
  }
  __syncthreads();
  // phase B: stable scatter
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; ++k) {
    const bool live = key[k] >= 0;
    const int dg = live ? ((key[k] >> shift) & (SORT_DIGITS - 1)) : SORT_DIGITS + lane;
    const unsigned peers = __match_any_sync(0xffffffffu, dg);
    const int rank = __popc(peers & ((1u << lane) - 1));
    unsigned dst = 0;
    if (live) dst = wcount[w][dg] + rank;
    __syncwarp();
    if (live && rank == 0) wcount[w][dg] += __popc(peers);
    __syncwarp();
    if (live) { keys_out[dst] = key[k]; vals_out[dst] = val[k]; }
  }
}

// =============================================================================================
// segmented reduction over the sorted order
// =============================================================================================
#define SEG_TILE 32
#define SEG_WARPS 8

template <int GD, int NV>
__global__ void __launch_bounds__(SEG_WARPS * 32) segment_tile_kernel(
    const void* __restrict__ d_out, const int* __restrict__ skeys, const int* __restrict__ spos, int64_t n,
    int64_t dim, int64_t rows, int64_t padding_idx, const float* __restrict__ scale_dev,
    const float* __restrict__ dot_table, float* __restrict__ d_table, float* __restrict__ partL,
    float* __restrict__ partR, float* __restrict__ dot_part) {
  const int lane = threadIdx.x & 31;
  const int vecs = (int)(dim >> 2);
  const int64_t gw = (int64_t)blockIdx.x * SEG_WARPS + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * SEG_WARPS;
  const int64_t ntiles = (n + SEG_TILE - 1) / SEG_TILE;
  const float scale = scale_dev ? __ldg(scale_dev) : 1.0f;
  float dot_local = 0.f;
  for (int64_t tile = gw; tile < ntiles; tile += nw) {
    const int64_t t0 = tile * SEG_TILE;
    const int cnt = (int)((n - t0) < SEG_TILE ? (n - t0) : SEG_TILE);
    const int my_key = lane < cnt ? skeys[t0 + lane] : -1;
    const int my_pos = lane < cnt ? spos[t0 + lane] : 0;
    const int prev_key = t0 > 0 ? skeys[t0 - 1] : -2;
    const int next_key = (t0 + SEG_TILE < n) ? skeys[t0 + SEG_TILE] : -3;
    float4 acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = __shfl_sync(0xffffffffu, my_key, 0);
    int seg_a = 0;
    // flush segment [seg_a, j) with key `cur`
    auto flush = [&](int j) {
      const bool left_done = (seg_a > 0) || (prev_key != cur);
      const bool right_done = (j < cnt) || (next_key != cur);
      const bool valid = cur >= 0 && cur < rows;
      if (dot_table && valid) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = lane + 32 * i;
          if (v < vecs) dot_local += dot4(ldg_f4(dot_table + (int64_t)cur * dim + 4 * v), acc[i]);
        }
      }
      if (valid && cur != padding_idx) {
        float* dst;
        float sc = 1.0f;
        if (left_done && right_done) { dst = d_table + (int64_t)cur * dim; sc = scale; }
        else if (!left_done) dst = partL + tile * dim;
        else dst = partR + tile * dim;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = lane + 32 * i;
          if (v < vecs)
            *reinterpret_cast<float4*>(dst + 4 * v) =
                make_float4(acc[i].x * sc, acc[i].y * sc, acc[i].z * sc, acc[i].w * sc);
        }
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // rows are fetched SEG_G at a time (independent loads in flight), then folded in sorted order
    constexpr int SEG_G = NV == 1 ? 8 : (NV == 2 ? 4 : 1);
    for (int j0 = 0; j0 < cnt; j0 += SEG_G) {
      float4 gbuf[SEG_G][NV];
#pragma unroll
      for (int u = 0; u < SEG_G; ++u) {
        const int p = __shfl_sync(0xffffffffu, my_pos, (j0 + u) & 31);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = lane + 32 * i;
          gbuf[u][i] = (j0 + u < cnt && v < vecs) ? load4<GD>(d_out, (int64_t)p * dim + 4 * v)
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < SEG_G; ++u) {
        const int j = j0 + u;
        const int k = __shfl_sync(0xffffffffu, my_key, j & 31);
        if (j < cnt) {
          if (k != cur) { flush(j); cur = k; seg_a = j; }
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            acc[i].x += gbuf[u][i].x; acc[i].y += gbuf[u][i].y; acc[i].z += gbuf[u][i].z; acc[i].w += gbuf[u][i].w;
          }
        }
      }
    }
    flush(cnt);
  }
  if (dot_part) {
    dot_local = warp_sum(dot_local);
    if (lane == 0) dot_part[gw] = dot_local;
  }
}

// Second level of the fix-up.  A Zipf head item spans hundreds of tiles; one warp adding their partial rows one after
// the other is a serial chain of L2 latencies (ncu r02b: segment_fixup 7-13 % issue, 0.3 % DRAM).  Groups of 32
// consecutive tiles that lie entirely INSIDE one segment are pre-reduced here, one warp per group, so that the chain
// walks 32 tiles per step.  Fixed summation order -> deterministic.
template <int NV>
__global__ void __launch_bounds__(SEG_WARPS * 32) segment_group_kernel(const int* __restrict__ skeys, int64_t n,
                                                                       int64_t dim, const float* __restrict__ partL,
                                                                       float* __restrict__ partG) {
  const int lane = threadIdx.x & 31;
  const int vecs = (int)(dim >> 2);
  const int64_t gw = (int64_t)blockIdx.x * SEG_WARPS + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * SEG_WARPS;
  const int64_t ntiles = (n + SEG_TILE - 1) / SEG_TILE;
  const int64_t ngroups = ntiles / 32;                       // only complete groups can be interior
  constexpr int FX_G = NV == 1 ? 8 : (NV == 2 ? 4 : 1);
  for (int64_t g = gw; g < ngroups; g += nw) {
    const int64_t e0 = g * 32 * SEG_TILE, e1 = e0 + 32 * SEG_TILE;
    if (e0 == 0 || e1 >= n) continue;
    const int k = skeys[e0];
    if (skeys[e0 - 1] != k || skeys[e1 - 1] != k || skeys[e1] != k) continue;
    float4 acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = 0; j0 < 32; j0 += FX_G) {
      float4 gb[FX_G][NV];
#pragma unroll
      for (int q = 0; q < FX_G; ++q)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = lane + 32 * i;
          gb[q][i] = v < vecs ? *reinterpret_cast<const float4*>(partL + (g * 32 + j0 + q) * dim + 4 * v)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int q = 0; q < FX_G; ++q)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          acc[i].x += gb[q][i].x; acc[i].y += gb[q][i].y; acc[i].z += gb[q][i].z; acc[i].w += gb[q][i].w;
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < vecs) *reinterpret_cast<float4*>(partG + g * dim + 4 * v) = acc[i];
    }
  }
}

// one warp per tile whose last segment continues into the following tiles and STARTS in this tile
template <int NV>
__global__ void __launch_bounds__(SEG_WARPS * 32) segment_fixup_kernel(
    const int* __restrict__ skeys, int64_t n, int64_t dim, int64_t rows, int64_t padding_idx,
    const float* __restrict__ scale_dev, float* __restrict__ d_table, const float* __restrict__ partL,
    const float* __restrict__ partR, const float* __restrict__ partG) {
  const int lane = threadIdx.x & 31;
  const int vecs = (int)(dim >> 2);
  const int64_t gw = (int64_t)blockIdx.x * SEG_WARPS + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * SEG_WARPS;
  const int64_t ntiles = (n + SEG_TILE - 1) / SEG_TILE;
  const float scale = scale_dev ? __ldg(scale_dev) : 1.0f;
  for (int64_t tile = gw; tile + 1 < ntiles; tile += nw) {
    const int64_t t0 = tile * SEG_TILE, t1 = t0 + SEG_TILE;      // t1 < n because tile+1 < ntiles
    const int k = skeys[t1 - 1];
    if (skeys[t1] != k) continue;                                 // last segment ends here
    if (skeys[t0] == k && t0 > 0 && skeys[t0 - 1] == k) continue; // whole tile belongs to an earlier head
    if (k < 0 || k >= rows || k == padding_idx) continue;
    float4 acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      acc[i] = v < vecs ? *reinterpret_cast<const float4*>(partR + tile * dim + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // walk the chain: whole groups of 32 tiles that lie inside the segment come pre-reduced from segment_group_kernel;
    // up to the next group boundary lane j tests whether the segment runs through tile u0+j, a ballot finds where it
    // ends, and the partial rows are summed in tile order with FX_G independent loads in flight
    constexpr int FX_G = NV == 1 ? 8 : (NV == 2 ? 4 : 1);
    bool more = true;
    int64_t u0 = tile + 1;
    while (more && u0 < ntiles) {
      if ((u0 & 31) == 0) {
        const int64_t e1 = (u0 + 32) * SEG_TILE;             // (skeys[u0 * SEG_TILE - 1] == skeys[u0 * SEG_TILE] == k: inside the chain)
        if (e1 < n && skeys[e1 - 1] == k && skeys[e1] == k) {
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int v = lane + 32 * i;
            if (v < vecs) {
              const float4 gq = *reinterpret_cast<const float4*>(partG + (u0 >> 5) * dim + 4 * v);
              acc[i].x += gq.x; acc[i].y += gq.y; acc[i].z += gq.z; acc[i].w += gq.w;
            }
          }
          u0 += 32;
          continue;
        }
      }
      const int nb = 32 - (int)(u0 & 31);                     // tiles up to the next group boundary
      const int64_t u = u0 + lane;
      bool through = false;                                   // segment covers tile u entirely AND continues past it
      if (lane < nb && u < ntiles) {
        const int64_t u1 = (u + 1) * SEG_TILE;
        through = (u1 < n) && skeys[u1 - 1] == k && skeys[u1] == k;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, through);
      const unsigned full = nb == 32 ? 0xffffffffu : ((1u << nb) - 1u);
      const int run = (bal == full) ? nb : __ffs(~bal) - 1;   // tiles u0..u0+run-1 pass through; u0+run ends the segment
      int last = run < nb ? run : nb - 1;                     // last tile of this batch that contributes a partL piece
      if (run < nb) more = false;
      if (u0 + last >= ntiles) last = (int)(ntiles - 1 - u0);
      for (int j0 = 0; j0 <= last; j0 += FX_G) {
        float4 gb[FX_G][NV];
#pragma unroll
        for (int q = 0; q < FX_G; ++q)
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int v = lane + 32 * i;
            gb[q][i] = (j0 + q <= last && v < vecs)
                           ? *reinterpret_cast<const float4*>(partL + (u0 + j0 + q) * dim + 4 * v)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
        for (int q = 0; q < FX_G; ++q)
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            acc[i].x += gb[q][i].x; acc[i].y += gb[q][i].y; acc[i].z += gb[q][i].z; acc[i].w += gb[q][i].w;
          }
      }
      u0 += nb;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < vecs)
        *reinterpret_cast<float4*>(d_table + (int64_t)k * dim + 4 * v) =
            make_float4(acc[i].x * scale, acc[i].y * scale, acc[i].z * scale, acc[i].w * scale);
    }
  }
}

__global__ void __launch_bounds__(1024) dot_finalize_kernel(const float* __restrict__ dot_part, int nparts,
                                                            float* __restrict__ dot_out) {
  // fixed association order (strided partial sums, then a warp tree): deterministic
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += dot_part[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) *dot_out += s;
  }
}

}  // namespace rs

// =============================================================================================
// C ABI
// =============================================================================================
using namespace rs;

#define DISPATCH_DT(dt, NAME, ...)                                      \
  switch (dt) {                                                         \
    case RS_F32: { constexpr int NAME = RS_F32; __VA_ARGS__; break; }   \
    case RS_F16: { constexpr int NAME = RS_F16; __VA_ARGS__; break; }   \
    case RS_BF16: { constexpr int NAME = RS_BF16; __VA_ARGS__; break; } \
    default: return RS_ERR_BAD_ARG;                                     \
  }

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct SeqBwdPlan {
  int nchunks, grid, small_rows, fast;
  size_t pos_bytes, small_bytes, gate_bytes, smem_bytes;
};
static int make_seq_bwd_plan(int64_t P, int64_t L, int64_t dim, int n_tables, const int64_t* table_rows,
                             const int* big_mode, SeqBwdPlan* pl) {
  if (L <= 0 || P % L != 0) return RS_ERR_BAD_ARG;
  const int64_t B = P / L;
  pl->nchunks = (int)((B + BWD_BC - 1) / BWD_BC);
  int64_t items = L * pl->nchunks;
  int64_t grid = (items + BWD_WARPS - 1) / BWD_WARPS;
  if (grid > 2 * RS_NUM_SMS) grid = 2 * RS_NUM_SMS;
  if (grid < 1) grid = 1;
  pl->grid = (int)grid;
  pl->fast = (dim == 128 && L <= BWD_WARPS * BWD_LPW) ? 1 : 0;
  if (pl->fast) {                      // one CTA streams whole batch rows; one positional partial per CTA
    pl->grid = (int)(B < 2 * RS_NUM_SMS ? (B < 1 ? 1 : B) : 2 * RS_NUM_SMS);
    pl->nchunks = pl->grid;
  }
  pl->small_rows = 0;
  for (int t = 0; t < n_tables; ++t)
    if (big_mode[t] == 2) pl->small_rows += (int)table_rows[t];
  pl->smem_bytes = (size_t)BWD_WARPS * pl->small_rows * dim * sizeof(float);
  if (pl->smem_bytes > 96 * 1024) return RS_ERR_UNSUPPORTED;     // 2 CTAs/SM
  pl->pos_bytes = align256((size_t)pl->nchunks * L * dim * sizeof(float));
  pl->small_bytes = align256((size_t)pl->grid * pl->small_rows * dim * sizeof(float));
  pl->gate_bytes = align256((size_t)pl->grid * RS_MAX_TABLES * sizeof(float));
  return RS_OK;
}

extern "C" size_t rs_seq_front_bwd_workspace_bytes(int64_t P, int64_t L, int64_t dim, int n_tables,
                                                   const int64_t* table_rows, const int* big_mode) {
  SeqBwdPlan pl;
  if (make_seq_bwd_plan(P, L, dim, n_tables, table_rows, big_mode, &pl) != RS_OK) return 0;
  return pl.pos_bytes + pl.small_bytes + pl.gate_bytes + 256;
}

extern "C" int rs_seq_front_bwd(const void* dx, int dx_dtype, const int64_t* const* ids, const float* const* tables,
                                const int64_t* table_rows, const int* big_mode, int n_tables, const float* gates,
                                int64_t L, int64_t P, int64_t dim, int64_t padding_idx, float* const* d_tables,
                                float* d_gates, float* d_pos, void* workspace, size_t workspace_bytes, void* stream) {
  if (P == 0) return RS_OK;
  if (!dx || n_tables < 0 || n_tables > RS_MAX_TABLES || dim <= 0 || (dim & 3) || !workspace) return RS_ERR_BAD_ARG;
  if (n_tables && (!gates || !d_gates)) return RS_ERR_BAD_ARG;
  SeqBwdPlan pl;
  int rc = make_seq_bwd_plan(P, L, dim, n_tables, table_rows, big_mode, &pl);
  if (rc != RS_OK) return rc;
  if (workspace_bytes < pl.pos_bytes + pl.small_bytes + pl.gate_bytes) return RS_ERR_WORKSPACE;
  SeqBwdParams prm;
  prm.n_tables = n_tables;
  prm.small_rows = pl.small_rows;
  int off = 0;
  for (int t = 0; t < RS_MAX_TABLES; ++t) {
    const bool on = t < n_tables;
    prm.ids[t] = on ? ids[t] : nullptr;
    prm.tables[t] = on ? tables[t] : nullptr;
    prm.d_tables[t] = on ? d_tables[t] : nullptr;
    prm.rows[t] = on ? table_rows[t] : 0;
    prm.mode[t] = on ? big_mode[t] : 0;
    prm.small_off[t] = 0;
    if (on && big_mode[t] == 2) { prm.small_off[t] = off; off += (int)table_rows[t]; }
    if (on && big_mode[t] != 0 && (!prm.ids[t] || !prm.tables[t] || !prm.d_tables[t])) return RS_ERR_BAD_ARG;
  }
  char* ws = (char*)workspace;
  float* pos_part = (float*)ws;
  float* small_part = (float*)(ws + pl.pos_bytes);
  float* gate_part = (float*)(ws + pl.pos_bytes + pl.small_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t B = P / L;
  const int need = (int)(((dim >> 2) + 31) / 32);
  if (need > 8) return RS_ERR_UNSUPPORTED;
#define LAUNCH_BWD(GD, NV)                                                                                     \
  do {                                                                                                         \
    if (pl.smem_bytes > 32 * 1024)                                                                             \
      cudaFuncSetAttribute(seq_front_bwd_kernel<GD, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                           (int)pl.smem_bytes);                                                                \
    seq_front_bwd_kernel<GD, NV><<<pl.grid, BWD_WARPS * 32, pl.smem_bytes, st>>>(                              \
        dx, prm, gates, L, B, dim, padding_idx, pl.nchunks, pos_part, small_part, gate_part);                  \
  } while (0)
#define LAUNCH_BWD128(GD)                                                                                       \
  do {                                                                                                         \
    if (pl.smem_bytes > 32 * 1024)                                                                             \
      cudaFuncSetAttribute(seq_front_bwd128_kernel<GD>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                           (int)pl.smem_bytes);                                                                \
    seq_front_bwd128_kernel<GD><<<pl.grid, BWD_WARPS * 32, pl.smem_bytes, st>>>(                               \
        dx, prm, gates, (int)L, B, padding_idx, pos_part, small_part, gate_part);                              \
  } while (0)
  if (pl.fast) {
    DISPATCH_DT(dx_dtype, GD, LAUNCH_BWD128(GD));
  } else {
    DISPATCH_DT(dx_dtype, GD, if (need <= 1) LAUNCH_BWD(GD, 1); else if (need <= 2) LAUNCH_BWD(GD, 2); else LAUNCH_BWD(GD, 8));
  }
  RS_LAUNCH_CHECK();
  const int64_t fin = L * dim + (int64_t)pl.small_rows * dim + RS_MAX_TABLES;
  seq_front_bwd_finalize<<<(int)((fin + 255) / 256), 256, 0, st>>>(prm, L, dim, pl.nchunks, pl.grid, pos_part,
                                                                   small_part, gate_part, d_pos, d_gates);
  RS_LAUNCH_CHECK();
  return RS_OK;
}

// ---------------------------------------------------------------------------------------------
static inline int sort_ntiles(int64_t n) { return (int)((n + SORT_TILE - 1) / SORT_TILE); }

extern "C" size_t rs_sort_ids_workspace_bytes(int64_t n) {
  return 4 * align256((size_t)n * sizeof(int)) + align256((size_t)SORT_DIGITS * (sort_ntiles(n) + 1) * sizeof(unsigned)) + 256;
}

extern "C" int rs_sort_ids(const int64_t* ids, int64_t n, int64_t rows, int64_t clamp_max, int32_t* sorted_ids,
                           int32_t* sorted_pos, void* workspace, size_t workspace_bytes, int* oob_flag,
                           void* stream) {
  if (n == 0) return RS_OK;
  if (!ids || !sorted_ids || !sorted_pos || !workspace || rows <= 0 || rows >= (1ll << 31) - 1 || n >= (1ll << 31))
    return RS_ERR_BAD_ARG;
  if (workspace_bytes < rs_sort_ids_workspace_bytes(n)) return RS_ERR_WORKSPACE;
  int bits = 1;
  while ((1ll << bits) <= rows) ++bits;           // keys lie in [0, rows] (rows == out-of-range marker)
  const int passes = (bits + SORT_BITS - 1) / SORT_BITS;
  const int ntiles = sort_ntiles(n);
  char* ws = (char*)workspace;
  const size_t nb = align256((size_t)n * sizeof(int));
  int* kbuf[2] = {(int*)ws, (int*)(ws + nb)};
  int* vbuf[2] = {(int*)(ws + 2 * nb), (int*)(ws + 3 * nb)};
  unsigned* hist = (unsigned*)(ws + 4 * nb);
  unsigned* totals = hist + (size_t)SORT_DIGITS * ntiles;
  cudaStream_t st = (cudaStream_t)stream;
  const int* kin = nullptr;
  const int* vin = nullptr;
  for (int p = 0; p < passes; ++p) {
    const int shift = SORT_BITS * p;
    int* kout = (p == passes - 1) ? sorted_ids : kbuf[p & 1];
    int* vout = (p == passes - 1) ? sorted_pos : vbuf[p & 1];
    if (p == 0) sort_hist_kernel<true><<<ntiles, SORT_THREADS, 0, st>>>(ids, nullptr, n, rows, clamp_max, shift, ntiles, hist, oob_flag);
    else sort_hist_kernel<false><<<ntiles, SORT_THREADS, 0, st>>>(nullptr, kin, n, rows, clamp_max, shift, ntiles, hist, nullptr);
    RS_LAUNCH_CHECK();
    sort_scan_kernel<<<SORT_DIGITS, SORT_THREADS, 0, st>>>(hist, ntiles, totals);
    RS_LAUNCH_CHECK();
    if (p == 0) sort_scatter_kernel<true><<<ntiles, SORT_THREADS, 0, st>>>(ids, nullptr, nullptr, n, rows, clamp_max, shift, ntiles, hist, totals, kout, vout);
    else sort_scatter_kernel<false><<<ntiles, SORT_THREADS, 0, st>>>(nullptr, kin, vin, n, rows, clamp_max, shift, ntiles, hist, totals, kout, vout);
    RS_LAUNCH_CHECK();
    kin = kout;
    vin = vout;
  }
  return RS_OK;
}

// ---------------------------------------------------------------------------------------------
static inline int seg_grid(int64_t n) {
  const int64_t ntiles = (n + SEG_TILE - 1) / SEG_TILE;
  int64_t g = (ntiles + SEG_WARPS - 1) / SEG_WARPS;
  const int64_t cap = (int64_t)RS_NUM_SMS * 8;
  return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}
extern "C" size_t rs_segment_reduce_workspace_bytes(int64_t n, int64_t dim) {
  const size_t ntiles = (size_t)((n + SEG_TILE - 1) / SEG_TILE);
  return 2 * align256(ntiles * dim * sizeof(float)) + align256((size_t)RS_NUM_SMS * 8 * SEG_WARPS * sizeof(float)) +
         align256((ntiles / 32 + 1) * dim * sizeof(float)) + 256;
}

extern "C" int rs_segment_reduce_rows(const void* d_out, int d_out_dtype, const int32_t* sorted_ids,
                                      const int32_t* sorted_pos, int64_t n, int64_t dim, int64_t rows,
                                      int64_t padding_idx, const float* scale_dev, const float* dot_table, float* d_table,
                                      float* dot_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (n == 0) return RS_OK;
  if (!d_out || !sorted_ids || !sorted_pos || !d_table || !workspace || dim <= 0 || (dim & 3)) return RS_ERR_BAD_ARG;
  if (workspace_bytes < rs_segment_reduce_workspace_bytes(n, dim)) return RS_ERR_WORKSPACE;
  if (dot_table && !dot_out) return RS_ERR_BAD_ARG;
  const size_t ntiles = (size_t)((n + SEG_TILE - 1) / SEG_TILE);
  const size_t pb = align256(ntiles * dim * sizeof(float));
  char* ws = (char*)workspace;
  float* partL = (float*)ws;
  float* partR = (float*)(ws + pb);
  float* dot_part = (float*)(ws + 2 * pb);
  float* partG = (float*)(ws + 2 * pb + align256((size_t)RS_NUM_SMS * 8 * SEG_WARPS * sizeof(float)));
  const int grid = seg_grid(n);
  if (rows <= 0) return RS_ERR_BAD_ARG;        // keys == rows mark out-of-range ids (rs_sort_ids) and are skipped
  cudaStream_t st = (cudaStream_t)stream;
  const int need = (int)(((dim >> 2) + 31) / 32);
  if (need > 8) return RS_ERR_UNSUPPORTED;
  const int64_t ngroups = (int64_t)ntiles / 32;
  const int ggrid = (int)(ngroups < SEG_WARPS ? 1 : (ngroups / SEG_WARPS < (int64_t)RS_NUM_SMS * 8 ? ngroups / SEG_WARPS + 1 : (int64_t)RS_NUM_SMS * 8));
#define LAUNCH_SEG(GD, NV)                                                                                    \
  do {                                                                                                        \
    segment_tile_kernel<GD, NV><<<grid, SEG_WARPS * 32, 0, st>>>(d_out, sorted_ids, sorted_pos, n, dim, rows, \
                                                                 padding_idx, scale_dev, dot_table, d_table,  \
                                                                 partL, partR, dot_table ? dot_part : nullptr); \
    segment_group_kernel<NV><<<ggrid, SEG_WARPS * 32, 0, st>>>(sorted_ids, n, dim, partL, partG);             \
    segment_fixup_kernel<NV><<<grid, SEG_WARPS * 32, 0, st>>>(sorted_ids, n, dim, rows, padding_idx, scale_dev, \
                                                              d_table, partL, partR, partG);                  \
  } while (0)
  DISPATCH_DT(d_out_dtype, GD, if (need <= 1) LAUNCH_SEG(GD, 1); else if (need <= 2) LAUNCH_SEG(GD, 2); else LAUNCH_SEG(GD, 8));
  RS_LAUNCH_CHECK_N(3);
  if (dot_table) {
    dot_finalize_kernel<<<1, 1024, 0, st>>>(dot_part, grid * SEG_WARPS, dot_out);
    RS_LAUNCH_CHECK();
  }
  return RS_OK;
}
