// Row-sharded tables: device-side routing of a batch's ids against `owner = id % world`.
//
// The reference is single-process; what shards is the wide item tables.  A rank needs every DISTINCT item
// row of its batch once (Zipf ids: ~20 k distinct among ~110 k tokens), so the exchange moves distinct ids only:
//   histogram of the batch's ids over the catalogue  ->  owner-major compaction (owner by owner, ascending local row)
//   -> request list [world, cap] (static capacity per owner, -1 padded: an equal-split all-to-all, no host-side split
//   sizes, graph-capturable) and, for every id, its slot in the arrival buffer.
// The same compaction over the ALL-REDUCED target histogram gives every rank the identical, owner-major list of the
// box-wide distinct targets (the negatives that span the box) without exchanging id lists.
// Integer atomics + scans only: deterministic.  No host synchronisation anywhere.
#include "common.cuh"
#include "../../include/rs_twotower.h"

namespace rs {

#define SR_MAX_WORLD 64

__global__ void __launch_bounds__(256) id_histogram_kernel(const int64_t* __restrict__ ids, int64_t n,
                                                           const int32_t* __restrict__ n_valid, int64_t n_bins,
                                                           int force_bin0, int32_t* __restrict__ cnt, int* __restrict__ oob) {
  const int64_t lim = n_valid ? min((int64_t)__ldg(n_valid), n) : n;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gtid == 0 && force_bin0) atomicAdd(cnt, 1);
  for (int64_t i = gtid; i < lim; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = __ldg(ids + i);
    if (id >= 0 && id < n_bins) atomicAdd(cnt + id, 1);
    else if (oob) *oob = 1;
  }
}

// ---- owner-major compaction in three small launches (tile counts -> scan of the tile counts -> write).  Bins are
// visited owner by owner (owner r holds ids r, r + world, ...), OC_TILE local rows per CTA, 4 consecutive local rows per
// thread: slot order = ascending local row within an owner.
#define OC_THREADS 256
#define OC_PER 4
#define OC_TILE (OC_THREADS * OC_PER)
#define OC_SCAN_THREADS 1024

__device__ __forceinline__ int oc_block_scan(int local_sum, int* total) {        // exclusive, OC_THREADS threads
  __shared__ int warp_sums[OC_THREADS / 32];
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
    int w = lane < OC_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < OC_THREADS / 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += n;
    }
    if (lane < OC_THREADS / 32) warp_sums[lane] = w;
    if (lane == OC_THREADS / 32 - 1) grand = w;
  }
  __syncthreads();
  const int before = (warp > 0 ? warp_sums[warp - 1] : 0) + (v - local_sum);
  *total = grand;
  __syncthreads();
  return before;
}

__device__ __forceinline__ int oc_present(const int32_t* __restrict__ cnt, int world, int64_t R, int64_t n_ids, int r,
                                          int64_t j, int* c_out) {
  int c = 0;
  if (j < R) {
    const int64_t id = j * world + r;
    if (id < n_ids) c = __ldg(cnt + id);
  }
  *c_out = c;
  return c > 0 ? 1 : 0;
}

__global__ void __launch_bounds__(OC_THREADS) oc_count_kernel(const int32_t* __restrict__ cnt, int world, int64_t R,
                                                              int64_t n_ids, int tiles_per_owner,
                                                              int32_t* __restrict__ tile_cnt) {
  const int r = blockIdx.x / tiles_per_owner, t = blockIdx.x % tiles_per_owner;
  const int64_t j0 = (int64_t)t * OC_TILE + (int64_t)threadIdx.x * OC_PER;
  int s = 0, c;
#pragma unroll
  for (int e = 0; e < OC_PER; ++e) s += oc_present(cnt, world, R, n_ids, r, j0 + e, &c);
  int total;
  (void)oc_block_scan(s, &total);
  if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}

// one CTA: exclusive scan of the tile counts; tile_off[b] = slots before tile b WITHIN its owner; owner_cnt[r]
__global__ void __launch_bounds__(OC_SCAN_THREADS) oc_scan_kernel(const int32_t* __restrict__ tile_cnt, int n_tiles,
                                                                  int tiles_per_owner, int world, int64_t cap,
                                                                  int32_t* __restrict__ tile_off,
                                                                  int32_t* __restrict__ owner_cnt,
                                                                  int32_t* __restrict__ meta) {
  __shared__ int warp_sums[32];
  __shared__ int seg[SR_MAX_WORLD + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n_tiles + OC_SCAN_THREADS - 1) / OC_SCAN_THREADS;
  const int b0 = tid * per, b1 = min(b0 + per, n_tiles);
  int s = 0;
  for (int b = b0; b < b1; ++b) s += tile_cnt[b];
  int v = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  if (lane == 31) warp_sums[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += n;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  int run = (warp > 0 ? warp_sums[warp - 1] : 0) + (v - s);
  const int total = warp_sums[31];
  for (int b = b0; b < b1; ++b) {
    if (b % tiles_per_owner == 0) seg[b / tiles_per_owner] = run;
    run += tile_cnt[b];
  }
  if (tid == 0) seg[world] = total;
  __syncthreads();
  run = (warp > 0 ? warp_sums[warp - 1] : 0) + (v - s);
  for (int b = b0; b < b1; ++b) {
    tile_off[b] = run - seg[b / tiles_per_owner];
    run += tile_cnt[b];
  }
  if (tid < world) owner_cnt[tid] = seg[tid + 1] - seg[tid];
  if (tid == 0) {
    int maxc = 0;
    for (int r = 0; r < world; ++r) maxc = max(maxc, seg[r + 1] - seg[r]);
    meta[0] = maxc;
    meta[1] = maxc > cap ? 1 : 0;
    meta[2] = total;
  }
}

__global__ void __launch_bounds__(OC_THREADS) oc_write_kernel(const int32_t* __restrict__ cnt, int world, int64_t R,
                                                              int64_t n_ids, int tiles_per_owner, int64_t cap,
                                                              const int32_t* __restrict__ tile_off,
                                                              const int32_t* __restrict__ owner_cnt,
                                                              int64_t* __restrict__ out_rows, int64_t* __restrict__ out_ids,
                                                              float* __restrict__ out_cnt, int32_t* __restrict__ slot_of) {
  const int r = blockIdx.x / tiles_per_owner, t = blockIdx.x % tiles_per_owner;
  const int64_t j0 = (int64_t)t * OC_TILE + (int64_t)threadIdx.x * OC_PER;
  int c[OC_PER], s = 0;
#pragma unroll
  for (int e = 0; e < OC_PER; ++e) s += oc_present(cnt, world, R, n_ids, r, j0 + e, &c[e]);
  int total;
  int sl = oc_block_scan(s, &total) + tile_off[blockIdx.x];
#pragma unroll
  for (int e = 0; e < OC_PER; ++e) {
    const int64_t j = j0 + e, id = j * world + r;
    if (j >= R || id >= n_ids) continue;
    if (c[e] > 0) {
      if (sl < cap) {
        const int64_t o = (int64_t)r * cap + sl;
        if (out_rows) out_rows[o] = j;
        if (out_ids) out_ids[o] = id;
        if (out_cnt) out_cnt[o] = (float)c[e];
        slot_of[id] = (int32_t)o;
      } else {
        slot_of[id] = -1;
      }
      ++sl;
    } else {
      slot_of[id] = -1;
    }
  }
  // empty slots behind this owner's list, shared out over the owner's tiles
  const int64_t used = min((int64_t)owner_cnt[r], cap);
  for (int64_t k = used + (int64_t)t * OC_THREADS + threadIdx.x; k < cap; k += (int64_t)tiles_per_owner * OC_THREADS) {
    const int64_t o = (int64_t)r * cap + k;
    if (out_rows) out_rows[o] = -1;
    if (out_ids) out_ids[o] = 0;
    if (out_cnt) out_cnt[o] = 0.f;
  }
}

__global__ void __launch_bounds__(256) lookup_i32_kernel(const int32_t* __restrict__ table, int64_t n_table,
                                                         const int64_t* __restrict__ ids, int64_t n, int64_t fill,
                                                         int64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = __ldg(ids + i);
    int64_t v = fill;
    if (id >= 0 && id < n_table) { const int32_t t = __ldg(table + id); v = t >= 0 ? (int64_t)t : fill; }
    out[i] = v;
  }
}

}  // namespace rs

using namespace rs;

extern "C" int rs_id_histogram(const int64_t* ids, int64_t n, const int32_t* n_valid_dev, int64_t n_bins, int force_bin0,
                               int32_t* cnt, int* oob_flag, void* stream) {
  if (!ids || !cnt || n < 0 || n_bins <= 0) return RS_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(cnt, 0, (size_t)n_bins * sizeof(int32_t), st);
  if (e != cudaSuccess) return (int)e;
  int64_t blocks = (n + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > (int64_t)RS_NUM_SMS * 8) blocks = (int64_t)RS_NUM_SMS * 8;
  id_histogram_kernel<<<(int)blocks, 256, 0, st>>>(ids, n, n_valid_dev, n_bins, force_bin0, cnt, oob_flag);
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" size_t rs_owner_compact_workspace_bytes(int world, int64_t rows_per_owner) {
  const int64_t tpo = (rows_per_owner + OC_TILE - 1) / OC_TILE;
  return (size_t)(2 * world * tpo + SR_MAX_WORLD) * sizeof(int32_t) + 256;
}

extern "C" int rs_owner_compact(const int32_t* cnt, int world, int64_t rows_per_owner, int64_t n_ids, int64_t cap,
                                int64_t* out_rows, int64_t* out_ids, float* out_cnt, int32_t* slot_of, int32_t* meta,
                                void* workspace, size_t workspace_bytes, void* stream) {
  if (!cnt || !slot_of || !meta || !workspace || world <= 0 || rows_per_owner <= 0 || cap <= 0 || n_ids <= 0)
    return RS_ERR_BAD_ARG;
  if (world > SR_MAX_WORLD || (int64_t)world * cap > 0x7fffffffLL || n_ids > (int64_t)world * rows_per_owner) return RS_ERR_UNSUPPORTED;
  if (workspace_bytes < rs_owner_compact_workspace_bytes(world, rows_per_owner)) return RS_ERR_WORKSPACE;
  const int tpo = (int)((rows_per_owner + OC_TILE - 1) / OC_TILE);
  const int n_tiles = world * tpo;
  int32_t* tile_cnt = (int32_t*)workspace;
  int32_t* tile_off = tile_cnt + n_tiles;
  int32_t* owner_cnt = tile_off + n_tiles;
  cudaStream_t st = (cudaStream_t)stream;
  oc_count_kernel<<<n_tiles, OC_THREADS, 0, st>>>(cnt, world, rows_per_owner, n_ids, tpo, tile_cnt);
  RS_LAUNCH_CHECK();
  oc_scan_kernel<<<1, OC_SCAN_THREADS, 0, st>>>(tile_cnt, n_tiles, tpo, world, cap, tile_off, owner_cnt, meta);
  RS_LAUNCH_CHECK();
  oc_write_kernel<<<n_tiles, OC_THREADS, 0, st>>>(cnt, world, rows_per_owner, n_ids, tpo, cap, tile_off, owner_cnt,
                                                  out_rows, out_ids, out_cnt, slot_of);
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_lookup_i32(const int32_t* table, int64_t n_table, const int64_t* ids, int64_t n, int64_t fill,
                             int64_t* out, void* stream) {
  if (!table || !ids || !out || n < 0) return RS_ERR_BAD_ARG;
  if (n == 0) return RS_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)RS_NUM_SMS * 8) blocks = (int64_t)RS_NUM_SMS * 8;
  lookup_i32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(table, n_table, ids, n, fill, out);
  RS_LAUNCH_CHECK();
  return RS_OK;
}
