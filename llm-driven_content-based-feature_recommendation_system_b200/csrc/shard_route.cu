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

#define SR_SCAN_THREADS 1024
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

// block-wide exclusive scan over per-thread partial sums (same helper as batch_index.cu)
__device__ __forceinline__ int sr_block_scan(int local_sum, int* total) {
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
    warp_sums[lane] = w;
    if (lane == 31) grand = w;
  }
  __syncthreads();
  const int before = (warp > 0 ? warp_sums[warp - 1] : 0) + (v - local_sum);
  *total = grand;
  __syncthreads();
  return before;
}

// bins in owner-major order: k = r * R + j  <->  id = j * world + r
__global__ void __launch_bounds__(SR_SCAN_THREADS) owner_compact_kernel(const int32_t* __restrict__ cnt, int world,
                                                                        int64_t R, int64_t n_ids, int64_t cap,
                                                                        int64_t* __restrict__ out_rows,
                                                                        int64_t* __restrict__ out_ids,
                                                                        float* __restrict__ out_cnt,
                                                                        int32_t* __restrict__ slot_of,
                                                                        int32_t* __restrict__ meta) {
  __shared__ int seg[SR_MAX_WORLD + 1];
  const int tid = threadIdx.x;
  const int64_t nb = (int64_t)world * R;
  const int64_t per = (nb + SR_SCAN_THREADS - 1) / SR_SCAN_THREADS;
  const int64_t k0 = (int64_t)tid * per, k1 = k0 + per < nb ? k0 + per : nb;
  int s = 0;
  for (int64_t k = k0; k < k1; ++k) {
    const int64_t id = (k % R) * world + k / R;
    s += (id < n_ids && cnt[id] > 0) ? 1 : 0;
  }
  int total;
  const int start = sr_block_scan(s, &total);
  int run = start;
  for (int64_t k = k0; k < k1; ++k) {
    if (k % R == 0) seg[k / R] = run;
    const int64_t id = (k % R) * world + k / R;
    run += (id < n_ids && cnt[id] > 0) ? 1 : 0;
  }
  if (tid == 0) seg[world] = total;
  __syncthreads();
  int maxc = 0;
  for (int r = 0; r < world; ++r) maxc = max(maxc, seg[r + 1] - seg[r]);
  if (tid == 0) {
    meta[0] = maxc;
    meta[1] = maxc > cap ? 1 : 0;
    meta[2] = total;
  }
  run = start;
  for (int64_t k = k0; k < k1; ++k) {
    const int r = (int)(k / R);
    const int64_t j = k % R, id = j * world + r;
    if (id >= n_ids) continue;
    const int c = cnt[id];
    if (c > 0) {
      const int sl = run - seg[r];
      if (sl < cap) {
        const int64_t o = (int64_t)r * cap + sl;
        out_rows[o] = j;
        if (out_ids) out_ids[o] = id;
        if (out_cnt) out_cnt[o] = (float)c;
        slot_of[id] = (int32_t)o;
      } else {
        slot_of[id] = -1;
      }
      ++run;
    } else {
      slot_of[id] = -1;
    }
  }
  // empty slots behind every owner's list
  for (int r = 0; r < world; ++r) {
    const int c = min((int64_t)(seg[r + 1] - seg[r]), cap);
    for (int64_t sl = c + tid; sl < cap; sl += SR_SCAN_THREADS) {
      const int64_t o = (int64_t)r * cap + sl;
      out_rows[o] = -1;
      if (out_ids) out_ids[o] = 0;
      if (out_cnt) out_cnt[o] = 0.f;
    }
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

extern "C" int rs_owner_compact(const int32_t* cnt, int world, int64_t rows_per_owner, int64_t n_ids, int64_t cap,
                                int64_t* out_rows, int64_t* out_ids, float* out_cnt, int32_t* slot_of, int32_t* meta,
                                void* stream) {
  if (!cnt || !out_rows || !slot_of || !meta || world <= 0 || rows_per_owner <= 0 || cap <= 0 || n_ids <= 0) return RS_ERR_BAD_ARG;
  if (world > SR_MAX_WORLD || (int64_t)world * cap > 0x7fffffffLL || n_ids > (int64_t)world * rows_per_owner) return RS_ERR_UNSUPPORTED;
  owner_compact_kernel<<<1, SR_SCAN_THREADS, 0, (cudaStream_t)stream>>>(cnt, world, rows_per_owner, n_ids, cap, out_rows,
                                                                       out_ids, out_cnt, slot_of, meta);
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
