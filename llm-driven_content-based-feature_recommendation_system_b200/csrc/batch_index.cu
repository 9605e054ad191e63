// N1: on-device batch assembly for the train step.
//
// The reference collates a batch on the host (SASRecDataset.__getitem__ + default collate,
// tower_code/v1_refine_usertower.py:204-306) and derives everything the step needs from it inside the step with
// boolean indexing (`output_1[valid_mask]`, `target_ids[valid_mask]`, tower_code/v1_usertower_train.py:794-804:
// nonzero + host synchronisation) -- the step's shapes are data dependent.  Here the collated [B, L] tensors go to the
// device as they are and ONE stream-ordered call derives every index the packed step consumes, into arrays of STATIC
// (bucketed) capacity: the true counts live in device scalars, rows / columns beyond them are inert padding (weight 0,
// count 0).  Static shapes are what lets one captured CUDA graph serve every batch of a shape bucket.
//
//   packed token order  : the valid time steps in batch-major order (t = cu[b] + rank within the user)
//   U1 grid [grid_cap]  : [T valid tokens | E "last = padding" positions | padding]
//   encoder tokens      : [view-1 valid T | view-2 valid T | view-1 extras E | view-2 extras E | padding] = 2*grid_cap,
//                         everything behind 2T is the attention kernel's zero tail (one pseudo-sequence)
//   columns             : the distinct target items in ascending id order with their multiplicities (histogram over
//                         the catalogue + scan: integer atomics, deterministic)
#include "common.cuh"
#include "../../include/rs_twotower.h"

namespace rs {

#define BI_THREADS 256
#define BI_SCAN_THREADS 1024

struct BiWs {
  int* lens;      // [B]
  int* ex;        // [B]   1: the position len-1 (counted from the left) is padding -> carried as an extra token
  int* cu;        // [B+1] exclusive scan of lens
  int* eoff;      // [B]   exclusive scan of ex
  int* cnt;       // [n_item_rows] occurrences of every item among the valid targets
  int* colidx;    // [n_item_rows] column of every present item
  int* cmeta;     // [4] meta of the column compaction (largest list, overflow, distinct targets)
  void* oc_ws;    // workspace of rs_owner_compact
};

__device__ __forceinline__ void user_masks(const uint8_t* pad, int64_t b, int64_t L, int lane, unsigned& m0, unsigned& m1) {
  const int l0 = lane, l1 = lane + 32;
  const bool v0 = l0 < L && pad[b * L + l0] == 0;
  const bool v1 = l1 < L && pad[b * L + l1] == 0;
  m0 = __ballot_sync(0xffffffffu, v0);
  m1 = __ballot_sync(0xffffffffu, v1);
}

// warp per user: sequence length, "extra" flag, histogram of the valid targets
__global__ void __launch_bounds__(BI_THREADS) bi_user_kernel(const uint8_t* __restrict__ pad,
                                                              const int64_t* __restrict__ target_ids, int64_t B, int64_t L,
                                                              int64_t n_item_rows, BiWs ws, int* __restrict__ meta) {
  const int lane = threadIdx.x & 31;
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  unsigned m0, m1;
  user_masks(pad, b, L, lane, m0, m1);
  const int len = __popc(m0) + __popc(m1);
  if (lane == 0) {
    ws.lens[b] = len;
    const int q = len > 0 ? len - 1 : 0;
    const bool q_valid = q < 32 ? ((m0 >> q) & 1u) : ((m1 >> (q - 32)) & 1u);
    ws.ex[b] = q_valid ? 0 : 1;
    if (len == 0) atomicOr(meta + 3, 2);                   // an empty sequence: not representable (the host checks too)
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int l = lane + 32 * h;
    if (((h ? m1 : m0) >> lane) & 1u) {
      const int64_t id = target_ids[b * L + l];
      if (id >= 0 && id < n_item_rows) atomicAdd(ws.cnt + id, 1);
      else atomicOr(meta + 3, 4);                          // target outside the catalogue
    }
  }
}

// block-wide exclusive scan helper: every thread owns `per` consecutive elements
__device__ __forceinline__ int block_excl_scan(int local_sum, int* total) {
  __shared__ int warp_sums[32];
  __shared__ int grand;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int v = local_sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  if (lane == 31) warp_sums[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += n;
    }
    warp_sums[lane] = w;                                   // inclusive over warps
    if (lane == 31) grand = w;
  }
  __syncthreads();
  const int before = (warp > 0 ? warp_sums[warp - 1] : 0) + (v - local_sum);
  *total = grand;
  __syncthreads();
  return before;
}

struct BiOut {
  int64_t *pk_item_ids, *pk_time_ids, *pk_pos_ids, *pk_index_2v, *fold_inv1, *fold_inv2;
  int32_t *cu_seqlens_2v, *row_cu;
  int64_t *select_2v, *users_2v, *main_tgt, *last_tgt;
  float* row_weight;
  int64_t* col_item_ids;
  float* col_counts;
  int64_t* pos_col;
  int32_t* meta;
};

// one CTA: scans over the users (sequence offsets, extras).  The column list (distinct targets, ascending id, with
// multiplicities) is the owner-major compaction of shard_route.cu with a single owner (rs_owner_compact, world = 1).
__global__ void __launch_bounds__(BI_SCAN_THREADS) bi_scan_kernel(int64_t B, int64_t n_item_rows, int64_t tok_cap,
                                                                   int64_t col_cap, int64_t grid_cap, BiWs ws, BiOut o,
                                                                   int counts_only) {
  const int tid = threadIdx.x;
  {
    const int per = (int)((B + BI_SCAN_THREADS - 1) / BI_SCAN_THREADS);
    const int64_t b0 = (int64_t)tid * per, b1 = b0 + per < B ? b0 + per : B;
    int sl = 0, se = 0;
    for (int64_t b = b0; b < b1; ++b) { sl += ws.lens[b]; se += ws.ex[b]; }
    int T, E;
    int pl = block_excl_scan(sl, &T);
    int pe = block_excl_scan(se, &E);
    if (tid == 0) {
      o.meta[0] = T; o.meta[1] = E;
      if (!counts_only && (T > tok_cap || (int64_t)T + E > grid_cap)) atomicOr(o.meta + 3, 1);
    }
    if (counts_only) return;
    for (int64_t b = b0; b < b1; ++b) {
      ws.cu[b] = pl; ws.eoff[b] = pe;
      o.row_cu[b] = pl;
      o.cu_seqlens_2v[b] = pl;
      o.cu_seqlens_2v[B + b] = T + pl;
      pl += ws.lens[b]; pe += ws.ex[b];
    }
    if (tid == 0) {
      ws.cu[B] = T;
      o.row_cu[B] = T;
      o.cu_seqlens_2v[2 * B] = 2 * T;
      o.cu_seqlens_2v[2 * B + 1] = (int)(2 * grid_cap);     // the zero tail: extras of both views + padding
    }
  }
}

// warp per user: the user's tokens into the packed arrays; then a grid-stride fill of the padding ranges
__global__ void __launch_bounds__(BI_THREADS) bi_fill_kernel(const uint8_t* __restrict__ pad,
                                                              const int64_t* __restrict__ item_ids,
                                                              const int64_t* __restrict__ time_ids,
                                                              const int64_t* __restrict__ target_ids, int64_t B, int64_t L,
                                                              int64_t n_item_rows, int64_t tok_cap, int64_t grid_cap,
                                                              BiWs ws, BiOut o) {
  const int lane = threadIdx.x & 31;
  const int64_t T = o.meta[0], E = o.meta[1];
  const float w = T > 0 ? 1.0f / (float)T : 0.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) {                 // the column compaction's counters into the index's meta
    o.meta[2] = ws.cmeta[2];
    if (ws.cmeta[1]) atomicOr(o.meta + 3, 1);
  }
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp0; b < B; b += nwarps) {
    unsigned m0, m1;
    user_masks(pad, b, L, lane, m0, m1);
    const int64_t c0 = ws.cu[b];
    const int len = ws.lens[b];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const unsigned m = h ? m1 : m0;
      if ((m >> lane) & 1u) {
        const int l = lane + 32 * h;
        const int rank = __popc(m & ((1u << lane) - 1u)) + (h ? __popc(m0) : 0);
        const int64_t t = c0 + rank, p = b * L + l;
        if (t < tok_cap && t < grid_cap) {
          const int64_t tg = target_ids[p];
          o.pk_item_ids[t] = item_ids[p];
          o.pk_time_ids[t] = time_ids[p];
          o.pk_pos_ids[t] = l + 1;
          o.main_tgt[t] = tg;
          o.pos_col[t] = (tg >= 0 && tg < n_item_rows) ? ws.colidx[tg] : 0;
          o.select_2v[t] = t;
          o.users_2v[t] = b;
          o.row_weight[t] = w;
          o.pk_index_2v[t] = t;
          o.fold_inv1[t] = t;
          if (T + t < 2 * grid_cap) { o.pk_index_2v[T + t] = t; o.fold_inv2[t] = T + t; }
        }
      }
    }
    if (lane == 0) {
      // the row DuoRec reads: position len-1 counted from the LEFT of the grid (v1_usertower_train.py:830)
      const int q = len > 0 ? len - 1 : 0;
      const int64_t pq = b * L + q;
      int64_t lp1, lp2;
      if (ws.ex[b]) {
        const int64_t e = ws.eoff[b], u = T + e;
        lp1 = 2 * T + e;
        lp2 = 2 * T + E + e;
        if (u < grid_cap && lp2 < 2 * grid_cap) {
          o.pk_item_ids[u] = item_ids[pq];
          o.pk_time_ids[u] = time_ids[pq];
          o.pk_pos_ids[u] = q + 1;
          o.pk_index_2v[lp1] = u;
          o.pk_index_2v[lp2] = u;
          o.fold_inv1[u] = lp1;
          o.fold_inv2[u] = lp2;
        } else {
          lp1 = 0; lp2 = 0;
        }
      } else {
        const int rank = q < 32 ? __popc(m0 & ((1u << q) - 1u)) : __popc(m0) + __popc(m1 & ((1u << (q - 32)) - 1u));
        lp1 = c0 + rank;
        lp2 = T + lp1;
        if (lp2 >= 2 * grid_cap) { lp1 = 0; lp2 = 0; }
      }
      o.select_2v[tok_cap + b] = lp1;
      o.select_2v[tok_cap + B + b] = lp2;
      o.users_2v[tok_cap + b] = b;
      o.users_2v[tok_cap + B + b] = B + b;
      o.last_tgt[b] = target_ids[pq];
    }
  }
  // ---- padding: main rows [T, tok_cap), U1 rows [T+E, grid_cap) and their two encoder slots each
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gn = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = T + gtid; t < tok_cap; t += gn) {
    // padding rows: encoder row t itself (a finite, real row: a view-2 token), profile row B -- the users column stays
    // ascending, which the head's sort-free backward relies on (towers._forward_packed, select_prefix)
    o.main_tgt[t] = 0; o.pos_col[t] = 0; o.select_2v[t] = t; o.users_2v[t] = B; o.row_weight[t] = 0.f;
  }
  const int64_t used = T + E, npad = grid_cap - used;
  for (int64_t k = gtid; k < npad; k += gn) {
    const int64_t u = used + k, s1 = 2 * used + k, s2 = 2 * used + npad + k;
    o.pk_item_ids[u] = 0; o.pk_time_ids[u] = 0; o.pk_pos_ids[u] = 0;
    o.fold_inv1[u] = s1; o.fold_inv2[u] = s2;
    o.pk_index_2v[s1] = u; o.pk_index_2v[s2] = u;
  }
}

// out[u,:] = x[i1[u],:] + x[i2[u],:]   (dim % 4 == 0): the two dropout views' gradients of one U1 row
template <int DT>
__global__ void __launch_bounds__(256) gather_add2_kernel(const void* __restrict__ x, const int64_t* __restrict__ i1,
                                                          const int64_t* __restrict__ i2, int64_t n, int64_t n_src,
                                                          int dim4, void* __restrict__ out) {
  const int64_t total = n * dim4;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = k / dim4;
    const int c = (int)(k - u * dim4);
    const int64_t a = __ldg(i1 + u), b = __ldg(i2 + u);
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
    if (a >= 0 && a < n_src) va = load4<DT>(x, (a * dim4 + c) * 4);
    if (b >= 0 && b < n_src) vb = load4<DT>(x, (b * dim4 + c) * 4);
    store4<DT>(out, k * 4, make_float4(va.x + vb.x, va.y + vb.y, va.z + vb.z, va.w + vb.w));
  }
}

__global__ void bi_merge_counts_kernel(int32_t* meta, const int32_t* cmeta) { meta[2] = cmeta[2]; }

}  // namespace rs

using namespace rs;

extern "C" size_t rs_owner_compact_workspace_bytes(int world, int64_t rows_per_owner);
extern "C" int rs_owner_compact(const int32_t* cnt, int world, int64_t rows_per_owner, int64_t n_ids, int64_t cap,
                                int64_t* out_rows, int64_t* out_ids, float* out_cnt, int32_t* slot_of, int32_t* meta,
                                void* workspace, size_t workspace_bytes, void* stream);

static inline size_t bi_al(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t rs_batch_index_workspace_bytes(int64_t B, int64_t L, int64_t n_item_rows) {
  (void)L;
  return 4 * bi_al((size_t)(B + 1) * 4) + 2 * bi_al((size_t)n_item_rows * 4) + 256 +
         bi_al(rs_owner_compact_workspace_bytes(1, n_item_rows));
}

static BiWs bi_carve(void* workspace, int64_t B, int64_t n_item_rows) {
  char* p = (char*)workspace;
  BiWs ws;
  ws.lens = (int*)p; p += bi_al((size_t)(B + 1) * 4);
  ws.ex = (int*)p; p += bi_al((size_t)(B + 1) * 4);
  ws.cu = (int*)p; p += bi_al((size_t)(B + 1) * 4);
  ws.eoff = (int*)p; p += bi_al((size_t)(B + 1) * 4);
  ws.cnt = (int*)p; p += bi_al((size_t)n_item_rows * 4);
  ws.colidx = (int*)p; p += bi_al((size_t)n_item_rows * 4);
  ws.cmeta = (int*)p; p += 256;
  ws.oc_ws = p;
  return ws;
}

extern "C" int rs_batch_index_counts(const uint8_t* padding_mask, const int64_t* target_ids, int64_t B, int64_t L,
                                     int64_t n_item_rows, int32_t* meta, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  if (!padding_mask || !target_ids || !meta || !workspace || B <= 0 || L <= 0 || L > 64 || n_item_rows <= 0)
    return RS_ERR_BAD_ARG;
  if (workspace_bytes < rs_batch_index_workspace_bytes(B, L, n_item_rows)) return RS_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  BiWs ws = bi_carve(workspace, B, n_item_rows);
  cudaError_t e = cudaMemsetAsync(ws.cnt, 0, (size_t)n_item_rows * 4, st);
  if (e != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(meta, 0, 8 * sizeof(int32_t), st)) != cudaSuccess) return (int)e;
  bi_user_kernel<<<(int)((B * 32 + BI_THREADS - 1) / BI_THREADS), BI_THREADS, 0, st>>>(padding_mask, target_ids, B, L,
                                                                                      n_item_rows, ws, meta);
  RS_LAUNCH_CHECK();
  BiOut o = {};
  o.meta = meta;
  bi_scan_kernel<<<1, BI_SCAN_THREADS, 0, st>>>(B, n_item_rows, 0, 0, 0, ws, o, 1);
  RS_LAUNCH_CHECK();
  // distinct valid targets = present bins of the histogram (capacity = all: nothing is truncated, outputs unused)
  int rc = rs_owner_compact(ws.cnt, 1, n_item_rows, n_item_rows, n_item_rows, nullptr, (int64_t*)nullptr, nullptr, ws.colidx,
                            ws.cmeta, ws.oc_ws, rs_owner_compact_workspace_bytes(1, n_item_rows), stream);
  (void)rc;
  bi_merge_counts_kernel<<<1, 1, 0, st>>>(meta, ws.cmeta);
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_batch_index_build(const rs_batch_index* d, void* workspace, size_t workspace_bytes, void* stream) {
  if (!d || !workspace) return RS_ERR_BAD_ARG;
  if (!d->padding_mask || !d->item_ids || !d->time_ids || !d->target_ids || d->B <= 0 || d->L <= 0 || d->L > 64 ||
      d->n_item_rows <= 0 || d->tok_cap <= 0 || d->col_cap <= 0 || d->grid_cap < d->tok_cap || (d->grid_cap % 64) != 0)
    return RS_ERR_BAD_ARG;
  if (!d->pk_item_ids || !d->pk_time_ids || !d->pk_pos_ids || !d->pk_index_2v || !d->fold_inv1 || !d->fold_inv2 ||
      !d->cu_seqlens_2v || !d->row_cu || !d->select_2v || !d->users_2v || !d->main_tgt || !d->last_tgt || !d->row_weight ||
      !d->col_item_ids || !d->col_counts || !d->pos_col || !d->meta)
    return RS_ERR_BAD_ARG;
  if (2 * d->grid_cap > 0x7fffffffLL) return RS_ERR_UNSUPPORTED;
  if (workspace_bytes < rs_batch_index_workspace_bytes(d->B, d->L, d->n_item_rows)) return RS_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  BiWs ws = bi_carve(workspace, d->B, d->n_item_rows);
  BiOut o;
  o.pk_item_ids = d->pk_item_ids; o.pk_time_ids = d->pk_time_ids; o.pk_pos_ids = d->pk_pos_ids;
  o.pk_index_2v = d->pk_index_2v; o.fold_inv1 = d->fold_inv1; o.fold_inv2 = d->fold_inv2;
  o.cu_seqlens_2v = d->cu_seqlens_2v; o.row_cu = d->row_cu;
  o.select_2v = d->select_2v; o.users_2v = d->users_2v; o.main_tgt = d->main_tgt; o.last_tgt = d->last_tgt;
  o.row_weight = d->row_weight; o.col_item_ids = d->col_item_ids; o.col_counts = d->col_counts; o.pos_col = d->pos_col;
  o.meta = d->meta;
  cudaError_t e = cudaMemsetAsync(ws.cnt, 0, (size_t)d->n_item_rows * 4, st);
  if (e != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(d->meta, 0, 8 * sizeof(int32_t), st)) != cudaSuccess) return (int)e;
  const int ugrid = (int)((d->B * 32 + BI_THREADS - 1) / BI_THREADS);
  bi_user_kernel<<<ugrid, BI_THREADS, 0, st>>>(d->padding_mask, d->target_ids, d->B, d->L, d->n_item_rows, ws, d->meta);
  RS_LAUNCH_CHECK();
  bi_scan_kernel<<<1, BI_SCAN_THREADS, 0, st>>>(d->B, d->n_item_rows, d->tok_cap, d->col_cap, d->grid_cap, ws, o, 0);
  RS_LAUNCH_CHECK();
  // columns: the distinct valid targets in ascending id order with their multiplicities; colidx[id] = column
  int rc = rs_owner_compact(ws.cnt, 1, d->n_item_rows, d->n_item_rows, d->col_cap, nullptr, d->col_item_ids, d->col_counts,
                            ws.colidx, ws.cmeta, ws.oc_ws, rs_owner_compact_workspace_bytes(1, d->n_item_rows), stream);
  if (rc != RS_OK) return rc;
  const int fgrid = ugrid < RS_NUM_SMS * 8 ? ugrid : RS_NUM_SMS * 8;
  bi_fill_kernel<<<fgrid, BI_THREADS, 0, st>>>(d->padding_mask, d->item_ids, d->time_ids, d->target_ids, d->B, d->L,
                                               d->n_item_rows, d->tok_cap, d->grid_cap, ws, o);
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_gather_add2(const void* x, int dtype, const int64_t* i1, const int64_t* i2, int64_t n, int64_t n_src,
                              int64_t dim, void* out, void* stream) {
  if (!x || !i1 || !i2 || !out || n < 0 || dim <= 0 || (dim % 4) != 0) return RS_ERR_BAD_ARG;
  if (n == 0) return RS_OK;
  const int64_t total = n * (dim / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)RS_NUM_SMS * 16) blocks = (int64_t)RS_NUM_SMS * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RS_F32) gather_add2_kernel<RS_F32><<<(int)blocks, 256, 0, st>>>(x, i1, i2, n, n_src, (int)(dim / 4), out);
  else if (dtype == RS_BF16) gather_add2_kernel<RS_BF16><<<(int)blocks, 256, 0, st>>>(x, i1, i2, n, n_src, (int)(dim / 4), out);
  else if (dtype == RS_F16) gather_add2_kernel<RS_F16><<<(int)blocks, 256, 0, st>>>(x, i1, i2, n, n_src, (int)(dim / 4), out);
  else return RS_ERR_BAD_ARG;
  RS_LAUNCH_CHECK();
  return RS_OK;
}
