// R1: fused dot-product scoring + running top-k.  The [n_users, n_items] score matrix of the reference
// (matmul then torch.topk) is never written: each CTA keeps a block of user rows in shared memory,
// streams item tiles through a cp.async double buffer, accumulates fp32 scores in registers (the
// reference scores in fp32 with TF32 off -- exact-id parity needs fp32 products), and only scores
// that beat the row's current k-th best reach the shared-memory candidate lists.
// Order of the result: score descending, equal scores by ascending id.
#include "common.cuh"
#include "../../include/rs_twotower.h"
#include <float.h>

namespace rs {

#define TK_THREADS 256
#define TK_BN 128            // items per tile
#define TK_KT 32             // k-slice per pipeline stage
#define TK_LDB (TK_KT + 4)   // padded row stride (floats) of an item-tile row: conflict-free LDS.128
#define TK_MAX_SPLIT 32

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = pred ? 16 : 0;                    // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ bool better(float v, int id, float wv, int wid) {
  return (v > wv) || (v == wv && id < wid);
}

// BM user rows per CTA; TM = BM/16 rows per thread; thread (ty, tx): rows ty*TM.., items tx + 16*c
// EXCL (hard-negative mining, C4/C5): a column j is ignored for row i when key[i] == key[j] or when the
// item-item similarity <rowitems_i, items_j> exceeds `thr` (i != j); the second Gram product shares the item tile.
struct MineArgs {
  const float* row_items;      // [n_users, dim]: the item vector that belongs to each row
  const int64_t* key_row;      // [n_users]
  const int64_t* key_col;      // [n_items]
  float thr;
  int* avail;                  // [n_users] += number of non-ignored columns seen by this CTA
};

template <int BM, bool EXCL>
__global__ void __launch_bounds__(TK_THREADS) topk_kernel(const float* __restrict__ users, int64_t n_users,
                                                          const float* __restrict__ items, int64_t n_items, int dim,
                                                          int k, int mask0, int nsplit, int64_t* __restrict__ out_ids,
                                                          float* __restrict__ out_scores,
                                                          float* __restrict__ part_scores,
                                                          int* __restrict__ part_ids, MineArgs mine,
                                                          const int* __restrict__ run_if) {
  if (run_if && *run_if == 0) return;      // (uniform) conditional launch: the tensor-core path did not overflow
  constexpr int TM = BM / 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lda = dim + 4;
  float* As = reinterpret_cast<float*>(smem_raw);                   // [BM][lda] (+ [BM][lda] row items when EXCL)
  float* A2 = As + BM * lda;
  float* Bs = As + (EXCL ? 2 : 1) * BM * lda;                       // [2][TK_BN][TK_LDB]
  float* Sc = Bs + 2 * TK_BN * TK_LDB;                              // [BM][TK_BN] passing scores
  unsigned* Mk = reinterpret_cast<unsigned*>(Sc + BM * TK_BN);      // [BM][4] pass bitmask
  float* thr = reinterpret_cast<float*>(Mk + BM * 4);               // [BM] current k-th best (or -inf)
  int* cnt = reinterpret_cast<int*>(thr + BM);                      // [BM]
  int* wslot = cnt + BM;                                            // [BM] slot of the worst kept entry
  float* Ls = reinterpret_cast<float*>(wslot + BM);                 // [BM][k]
  int* Li = reinterpret_cast<int*>(Ls + (size_t)BM * k);            // [BM][k]

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t u0 = (int64_t)blockIdx.x * BM;
  // item range of this split, in whole tiles
  const int64_t tiles_total = (n_items + TK_BN - 1) / TK_BN;
  const int64_t tiles_per = (tiles_total + nsplit - 1) / nsplit;
  const int64_t tile_lo = (int64_t)blockIdx.y * tiles_per;
  const int64_t tile_hi = min(tile_lo + tiles_per, tiles_total);

  for (int i = tid; i < BM; i += TK_THREADS) { thr[i] = -INFINITY; cnt[i] = 0; wslot[i] = 0; }
  for (int i = tid; i < BM * 4; i += TK_THREADS) Mk[i] = 0u;
  // users block -> smem (zero rows past n_users)
  const int vpr = dim >> 2;
  for (int i = tid; i < BM * vpr; i += TK_THREADS) {
    const int r = i / vpr, v = i % vpr;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u0 + r < n_users) x = ldg_f4(users + (u0 + r) * dim + 4 * v);
    *reinterpret_cast<float4*>(As + r * lda + 4 * v) = x;
    if (EXCL) {
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      if (u0 + r < n_users) y = ldg_f4(mine.row_items + (u0 + r) * dim + 4 * v);
      *reinterpret_cast<float4*>(A2 + r * lda + 4 * v) = y;
    }
  }
  int64_t my_key[TM];
  int navail[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    navail[i] = 0;
    my_key[i] = (EXCL && u0 + ty * TM + i < n_users) ? __ldg(mine.key_row + u0 + ty * TM + i) : -1;
  }
  const int ksteps = dim / TK_KT;
  const int64_t nstages = (tile_hi - tile_lo) * ksteps;

  auto issue = [&](int64_t stage) {
    const int64_t tile = tile_lo + stage / ksteps;
    const int kt = (int)(stage % ksteps);
    float* dst = Bs + (stage & 1) * TK_BN * TK_LDB;
    // 128 rows x 8 chunks of 16 B
    for (int q = tid; q < TK_BN * (TK_KT / 4); q += TK_THREADS) {
      const int r = q >> 3, c = q & 7;
      const int64_t item = tile * TK_BN + r;
      const bool ok = item < n_items;
      cp_async16(dst + r * TK_LDB + 4 * c, items + (ok ? item : 0) * dim + kt * TK_KT + 4 * c, ok);
    }
    cp_async_commit();
  };

  float acc[TM][8];
  float acg[EXCL ? TM : 1][8];
  if (nstages > 0) issue(0);
  for (int64_t stage = 0; stage < nstages; ++stage) {
    const int kt = (int)(stage % ksteps);
    if (kt == 0) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) { acc[i][c] = 0.f; if (EXCL) acg[i][c] = 0.f; }
    }
    if (stage + 1 < nstages) { issue(stage + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const float* bs = Bs + (stage & 1) * TK_BN * TK_LDB;
#pragma unroll
    for (int k4 = 0; k4 < TK_KT; k4 += 4) {
      float4 a[TM], b[8];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(As + (ty * TM + i) * lda + kt * TK_KT + k4);
#pragma unroll
      for (int c = 0; c < 8; ++c) b[c] = *reinterpret_cast<const float4*>(bs + (tx + 16 * c) * TK_LDB + k4);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          acc[i][c] = fmaf(a[i].x, b[c].x, acc[i][c]);
          acc[i][c] = fmaf(a[i].y, b[c].y, acc[i][c]);
          acc[i][c] = fmaf(a[i].z, b[c].z, acc[i][c]);
          acc[i][c] = fmaf(a[i].w, b[c].w, acc[i][c]);
        }
      if (EXCL) {
        float4 a2[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) a2[i] = *reinterpret_cast<const float4*>(A2 + (ty * TM + i) * lda + kt * TK_KT + k4);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            acg[i][c] = fmaf(a2[i].x, b[c].x, acg[i][c]);
            acg[i][c] = fmaf(a2[i].y, b[c].y, acg[i][c]);
            acg[i][c] = fmaf(a2[i].z, b[c].z, acg[i][c]);
            acg[i][c] = fmaf(a2[i].w, b[c].w, acg[i][c]);
          }
      }
    }
    if (kt == ksteps - 1) {
      // ---- threshold test in registers; only passing scores touch shared memory
      const int64_t tile = tile_lo + stage / ksteps;
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int r = ty * TM + i;
        const float t = thr[r];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int col = tx + 16 * c;
          const int64_t item = tile * TK_BN + col;
          bool live = item < n_items && !(mask0 && item == 0);
          if (EXCL && live) {
            const bool self = (item == u0 + r);
            live = !(my_key[i] == __ldg(mine.key_col + item)) && !(acg[i][c] > mine.thr && !self) && (u0 + r < n_users);
            navail[i] += live ? 1 : 0;
          }
          if (live && acc[i][c] >= t) {
            Sc[r * TK_BN + col] = acc[i][c];
            atomicOr(&Mk[r * 4 + (col >> 5)], 1u << (col & 31));
          }
        }
      }
      __syncthreads();
      // ---- merge: warp w owns rows w, w+8, ...
      for (int r = wid; r < BM; r += TK_THREADS / 32) {
        unsigned m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = Mk[r * 4 + j];
        if ((m[0] | m[1] | m[2] | m[3]) == 0u) continue;
        float* ls = Ls + (size_t)r * k;
        int* li = Li + (size_t)r * k;
        int n = cnt[r];
        int ws = wslot[r];
        float wv = n == k ? ls[ws] : -INFINITY;
        int wi = n == k ? li[ws] : 0x7fffffff;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          unsigned bits = m[j];
          while (bits) {
            const int col = j * 32 + __ffs(bits) - 1;
            bits &= bits - 1;
            const float v = Sc[r * TK_BN + col];
            const int id = (int)(tile * TK_BN + col);
            if (n < k) {
              if (lane == 0) { ls[n] = v; li[n] = id; }
              ++n;
              __syncwarp();
              if (n < k) continue;
            } else {
              if (!better(v, id, wv, wi)) continue;
              if (lane == 0) { ls[ws] = v; li[ws] = id; }
              __syncwarp();
            }
            // (re)compute the worst kept entry: warp arg-min under (score asc, id desc)
            float bv = INFINITY; int bi = -1, bs_ = 0;
            for (int e = lane; e < k; e += 32) {
              const float ev = ls[e]; const int ei = li[e];
              if (better(bv, bi, ev, ei)) { bv = ev; bi = ei; bs_ = e; }     // (ev,ei) is worse than (bv,bi)
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
              const int os = __shfl_xor_sync(0xffffffffu, bs_, o);
              if (better(bv, bi, ov, oi)) { bv = ov; bi = oi; bs_ = os; }
            }
            wv = bv; wi = bi; ws = bs_;
          }
        }
        if (lane == 0) {
          cnt[r] = n; wslot[r] = ws;
          thr[r] = n == k ? wv : -INFINITY;
#pragma unroll
          for (int j = 0; j < 4; ++j) Mk[r * 4 + j] = 0u;
        }
      }
    }
    __syncthreads();
  }

  if (EXCL) {
#pragma unroll
    for (int i = 0; i < TM; ++i)
      if (navail[i] && u0 + ty * TM + i < n_users) atomicAdd(mine.avail + u0 + ty * TM + i, navail[i]);
  }
  // ---- emit: rank sort of each row's list (score desc, id asc)
  for (int r = wid; r < BM; r += TK_THREADS / 32) {
    const int64_t u = u0 + r;
    if (u >= n_users) continue;
    const float* ls = Ls + (size_t)r * k;
    const int* li = Li + (size_t)r * k;
    const int n = cnt[r];
    for (int e = lane; e < k; e += 32) {
      if (e < n) {
        const float v = ls[e]; const int id = li[e];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += better(ls[j], li[j], v, id) ? 1 : 0;
        if (nsplit == 1) { out_ids[u * k + rank] = id; out_scores[u * k + rank] = v; }
        else {
          part_scores[(u * nsplit + blockIdx.y) * k + rank] = v;
          part_ids[(u * nsplit + blockIdx.y) * k + rank] = id;
        }
      } else {      // fewer than k candidates in this split: pad
        if (nsplit == 1) { out_ids[u * k + e] = -1; out_scores[u * k + e] = -INFINITY; }
        else { part_scores[(u * nsplit + blockIdx.y) * k + e] = -INFINITY; part_ids[(u * nsplit + blockIdx.y) * k + e] = 0x7fffffff; }
      }
    }
  }
}

// k-way merge of nsplit sorted partial lists, one warp per user row (lane = list)
__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ part_scores,
                                                         const int* __restrict__ part_ids, int64_t n_users, int k,
                                                         int nsplit, int64_t* __restrict__ out_ids,
                                                         float* __restrict__ out_scores,
                                                         const int* __restrict__ run_if) {
  if (run_if && *run_if == 0) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t u = warp; u < n_users; u += nwarps) {
    const float* ps = part_scores + (u * nsplit + lane) * k;
    const int* pi = part_ids + (u * nsplit + lane) * k;
    int head = 0;
    float hv = -INFINITY; int hi = 0x7fffffff;
    if (lane < nsplit) { hv = ps[0]; hi = pi[0]; }
    for (int o = 0; o < k; ++o) {
      float bv = hv; int bi = hi; int bl = lane;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, s);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
        const int ol = __shfl_xor_sync(0xffffffffu, bl, s);
        if (better(ov, oi, bv, bi) || (ov == bv && oi == bi && ol < bl)) { bv = ov; bi = oi; bl = ol; }
      }
      if (lane == 0) { out_scores[u * k + o] = bv; out_ids[u * k + o] = (bi == 0x7fffffff) ? -1 : bi; }
      if (lane == bl) {
        ++head;
        if (head < k) { hv = ps[head]; hi = pi[head]; } else { hv = -INFINITY; hi = 0x7fffffff; }
      }
    }
  }
}

}  // namespace rs

using namespace rs;

struct TopkPlan { int bm, nsplit; size_t smem; int64_t blocks; };

static size_t topk_smem(int bm, int dim, int k, bool excl = false) {
  size_t f = (size_t)(excl ? 2 : 1) * bm * (dim + 4) + 2 * TK_BN * TK_LDB + (size_t)bm * TK_BN;   // As(,A2), Bs, Sc
  size_t bytes = f * 4 + (size_t)bm * 4 * 4 + (size_t)bm * 4 * 3 + (size_t)bm * k * 8;
  return bytes;
}
static int make_topk_plan(int64_t n_users, int64_t n_items, int64_t dim, int64_t k, TopkPlan* pl, bool excl = false) {
  if (dim <= 0 || dim % TK_KT != 0 || dim > 512 || k <= 0 || k > 1024 || k > n_items) return RS_ERR_UNSUPPORTED;
  const int cand[3] = {64, 32, 16};
  pl->bm = 0;
  for (int i = excl ? 1 : 0; i < 3; ++i) {          // the mining variant carries two accumulator sets: BM <= 32
    if (topk_smem(cand[i], (int)dim, (int)k, excl) <= 200 * 1024) { pl->bm = cand[i]; break; }
  }
  if (!pl->bm) return RS_ERR_UNSUPPORTED;
  // small user batches: prefer more CTAs
  while (pl->bm > 16 && (n_users + pl->bm - 1) / pl->bm < RS_NUM_SMS) pl->bm >>= 1;
  pl->smem = topk_smem(pl->bm, (int)dim, (int)k, excl);
  pl->blocks = (n_users + pl->bm - 1) / pl->bm;
  const int64_t tiles_total = (n_items + TK_BN - 1) / TK_BN;
  int64_t ns = (2 * RS_NUM_SMS + pl->blocks - 1) / pl->blocks;
  const int64_t max_by_k = tiles_total / ((k + TK_BN - 1) / TK_BN + 1);   // each split sees well over k items
  if (ns > max_by_k) ns = max_by_k;
  if (ns > TK_MAX_SPLIT) ns = TK_MAX_SPLIT;
  if (ns < 1) ns = 1;
  // drop empty trailing splits
  const int64_t tiles_per = (tiles_total + ns - 1) / ns;
  ns = (tiles_total + tiles_per - 1) / tiles_per;
  pl->nsplit = (int)ns;
  return RS_OK;
}

extern "C" size_t rs_topk_workspace_bytes(int64_t n_users, int64_t n_items, int64_t dim, int64_t k) {
  TopkPlan pl;
  if (make_topk_plan(n_users, n_items, dim, k, &pl) != RS_OK || pl.nsplit == 1) return 256;
  return (size_t)n_users * pl.nsplit * k * 8 + 256;
}

// `run_if` (device int, may be NULL): when given, every kernel of this call returns at once unless *run_if != 0 -- the
// exact path behind the tensor-core candidate path (topk_tc.cu), taken only when that path flagged an overflow
int rs_retrieve_topk_cond(const float* users, int64_t n_users, const float* items, int64_t n_items,
                          int64_t dim, int64_t k, int mask_index0, int64_t* out_ids, float* out_scores,
                          void* workspace, size_t workspace_bytes, const int* run_if, void* stream) {
  if (n_users == 0) return RS_OK;
  if (!users || !items || !out_ids || !out_scores) return RS_ERR_BAD_ARG;
  TopkPlan pl;
  int rc = make_topk_plan(n_users, n_items, dim, k, &pl);
  if (rc != RS_OK) return rc;
  float* ps = nullptr;
  int* pi = nullptr;
  if (pl.nsplit > 1) {
    const size_t need = (size_t)n_users * pl.nsplit * k * 8;
    if (!workspace || workspace_bytes < need) return RS_ERR_WORKSPACE;
    ps = (float*)workspace;
    pi = (int*)((char*)workspace + (size_t)n_users * pl.nsplit * k * 4);
  }
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)pl.blocks, (unsigned)pl.nsplit);
  MineArgs none = {};
#define LAUNCH_TK(BM)                                                                                       \
  do {                                                                                                      \
    cudaError_t e = cudaFuncSetAttribute(topk_kernel<BM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)pl.smem);                                                     \
    if (e != cudaSuccess) return (int)e;                                                                    \
    topk_kernel<BM, false><<<grid, TK_THREADS, pl.smem, st>>>(users, n_users, items, n_items, (int)dim, (int)k, \
                                                       mask_index0, pl.nsplit, out_ids, out_scores, ps, pi, none, run_if); \
  } while (0)
  if (pl.bm == 64) LAUNCH_TK(64); else if (pl.bm == 32) LAUNCH_TK(32); else LAUNCH_TK(16);
  RS_LAUNCH_CHECK();
  if (pl.nsplit > 1) {
    const int g = grid_for_warps(n_users, 8, 8);
    topk_merge_kernel<<<g, 256, 0, st>>>(ps, pi, n_users, (int)k, pl.nsplit, out_ids, out_scores, run_if);
    RS_LAUNCH_CHECK();
  }
  return RS_OK;
}

extern "C" int rs_retrieve_topk(const float* users, int64_t n_users, const float* items, int64_t n_items,
                                int64_t dim, int64_t k, int mask_index0, int64_t* out_ids, float* out_scores,
                                void* workspace, size_t workspace_bytes, void* stream) {
  return rs_retrieve_topk_cond(users, n_users, items, n_items, dim, k, mask_index0, out_ids, out_scores, workspace,
                               workspace_bytes, nullptr, stream);
}

// ---------------------------------------------------------------------------------------------
// C4 / C5 hard-negative mining
extern "C" size_t rs_mine_workspace_bytes(int64_t n, int64_t dim, int64_t k) {
  TopkPlan pl;
  if (make_topk_plan(n, n, dim, k, &pl, true) != RS_OK || pl.nsplit == 1) return 256;
  return (size_t)n * pl.nsplit * k * 8 + 256;
}

extern "C" int rs_mine_hard_negatives(const float* u, const float* v, const int64_t* key, int64_t n, int64_t dim,
                                      int64_t k, float hnm_threshold, int64_t* out_ids, float* out_scores,
                                      int32_t* avail, void* workspace, size_t workspace_bytes, void* stream) {
  if (n == 0) return RS_OK;
  if (!u || !v || !key || !out_ids || !out_scores || !avail) return RS_ERR_BAD_ARG;
  TopkPlan pl;
  int rc = make_topk_plan(n, n, dim, k, &pl, true);
  if (rc != RS_OK) return rc;
  float* ps = nullptr;
  int* pi = nullptr;
  if (pl.nsplit > 1) {
    const size_t need = (size_t)n * pl.nsplit * k * 8;
    if (!workspace || workspace_bytes < need) return RS_ERR_WORKSPACE;
    ps = (float*)workspace;
    pi = (int*)((char*)workspace + (size_t)n * pl.nsplit * k * 4);
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(avail, 0, (size_t)n * sizeof(int32_t), st);
  if (e != cudaSuccess) return (int)e;
  MineArgs ma = {v, key, key, hnm_threshold, avail};
  dim3 grid((unsigned)pl.blocks, (unsigned)pl.nsplit);
#define LAUNCH_MINE(BM)                                                                                     \
  do {                                                                                                      \
    e = cudaFuncSetAttribute(topk_kernel<BM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem); \
    if (e != cudaSuccess) return (int)e;                                                                    \
    topk_kernel<BM, true><<<grid, TK_THREADS, pl.smem, st>>>(u, n, v, n, (int)dim, (int)k, 0, pl.nsplit,    \
                                                             out_ids, out_scores, ps, pi, ma, nullptr);    \
  } while (0)
  if (pl.bm == 32) LAUNCH_MINE(32); else LAUNCH_MINE(16);
  RS_LAUNCH_CHECK();
  if (pl.nsplit > 1) {
    const int g = grid_for_warps(n, 8, 8);
    topk_merge_kernel<<<g, 256, 0, st>>>(ps, pi, n, (int)k, pl.nsplit, out_ids, out_scores, nullptr);
    RS_LAUNCH_CHECK();
  }
  return RS_OK;
}

// ---------------------------------------------------------------------------------------------
// sparse logits: out[i,c] = scale * <a_i, b_idx[i,c]> - bias[idx]   (-inf for idx < 0 or equal keys)
namespace rs {
template <int DT>
__global__ void __launch_bounds__(256) sparse_logits_fwd_kernel(const void* __restrict__ a, const void* __restrict__ b,
                                                                const int64_t* __restrict__ idx, int64_t n, int64_t m,
                                                                int kk, int dim, float scale,
                                                                const float* __restrict__ bias,
                                                                const int64_t* __restrict__ key_row,
                                                                const int64_t* __restrict__ key_col,
                                                                float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int vecs = dim >> 2;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = warp; i < n; i += nwarps) {
    float4 av[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) av[q] = (lane + 32 * q < vecs) ? load4<DT>(a, i * dim + 4 * (lane + 32 * q)) : make_float4(0, 0, 0, 0);
    const int64_t kr = key_row ? __ldg(key_row + i) : -1;
    for (int c = 0; c < kk; ++c) {
      const int64_t j = __ldg(idx + i * kk + c);
      float r = -INFINITY;
      if (j >= 0 && j < m && !(key_row && kr == __ldg(key_col + j))) {
        float d = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (lane + 32 * q < vecs) d += dot4(av[q], load4<DT>(b, j * dim + 4 * (lane + 32 * q)));
        d = warp_sum(d);
        r = d * scale - (bias ? __ldg(bias + j) : 0.f);
      }
      if (lane == 0) out[i * kk + c] = r;
    }
  }
}

// d_a[i] += sum_c g[i,c]*scale*b[idx];  d_b[idx] += g[i,c]*scale*a[i]   (fp32 outputs, vector atomics on d_b)
template <int DT>
__global__ void __launch_bounds__(256) sparse_logits_bwd_kernel(const void* __restrict__ a, const void* __restrict__ b,
                                                                const int64_t* __restrict__ idx, int64_t n, int64_t m,
                                                                int kk, int dim, float scale,
                                                                const int64_t* __restrict__ key_row,
                                                                const int64_t* __restrict__ key_col,
                                                                const float* __restrict__ g, float* __restrict__ d_a,
                                                                float* __restrict__ d_b) {
  const int lane = threadIdx.x & 31;
  const int vecs = dim >> 2;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = warp; i < n; i += nwarps) {
    float4 av[4], acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      av[q] = (lane + 32 * q < vecs) ? load4<DT>(a, i * dim + 4 * (lane + 32 * q)) : make_float4(0, 0, 0, 0);
      acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int64_t kr = key_row ? __ldg(key_row + i) : -1;
    for (int c = 0; c < kk; ++c) {
      const int64_t j = __ldg(idx + i * kk + c);
      if (j < 0 || j >= m || (key_row && kr == __ldg(key_col + j))) continue;
      const float w = __ldg(g + i * kk + c) * scale;
      if (w == 0.f) continue;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int v = lane + 32 * q;
        if (v < vecs) {
          const float4 bv = load4<DT>(b, j * dim + 4 * v);
          acc[q] = fma4(acc[q], bv, w);
          red_add_f4(d_b + j * dim + 4 * v, make_float4(w * av[q].x, w * av[q].y, w * av[q].z, w * av[q].w));
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int v = lane + 32 * q;
      if (v < vecs) *reinterpret_cast<float4*>(d_a + i * dim + 4 * v) = acc[q];
    }
  }
}
}  // namespace rs

#define SL_DISPATCH(dt, NAME, ...)                                      \
  switch (dt) {                                                         \
    case RS_F32: { constexpr int NAME = RS_F32; __VA_ARGS__; break; }   \
    case RS_F16: { constexpr int NAME = RS_F16; __VA_ARGS__; break; }   \
    case RS_BF16: { constexpr int NAME = RS_BF16; __VA_ARGS__; break; } \
    default: return RS_ERR_BAD_ARG;                                     \
  }

extern "C" int rs_sparse_logits_fwd(const void* a, const void* b, int ab_dtype, const int64_t* idx, int64_t n,
                                    int64_t m, int64_t k, int64_t dim, float scale, const float* bias,
                                    const int64_t* key_row, const int64_t* key_col, float* out, void* stream) {
  if (n == 0 || k == 0) return RS_OK;
  if (!a || !b || !idx || !out || dim <= 0 || (dim & 3) || dim > 512 || ((key_row != nullptr) != (key_col != nullptr)))
    return RS_ERR_BAD_ARG;
  const int grid = grid_for_warps(n, 8, 8);
  cudaStream_t st = (cudaStream_t)stream;
  SL_DISPATCH(ab_dtype, DT, (rs::sparse_logits_fwd_kernel<DT><<<grid, 256, 0, st>>>(a, b, idx, n, m, (int)k, (int)dim, scale,
                                                                                 bias, key_row, key_col, out)));
  RS_LAUNCH_CHECK();
  return RS_OK;
}

extern "C" int rs_sparse_logits_bwd(const void* a, const void* b, int ab_dtype, const int64_t* idx, int64_t n,
                                    int64_t m, int64_t k, int64_t dim, float scale, const int64_t* key_row,
                                    const int64_t* key_col, const float* g, float* d_a, float* d_b, void* stream) {
  if (n == 0) return RS_OK;
  if (!a || !b || !idx || !g || !d_a || !d_b || dim <= 0 || (dim & 3) || dim > 512) return RS_ERR_BAD_ARG;
  const int grid = grid_for_warps(n, 8, 8);
  cudaStream_t st = (cudaStream_t)stream;
  SL_DISPATCH(ab_dtype, DT, (rs::sparse_logits_bwd_kernel<DT><<<grid, 256, 0, st>>>(a, b, idx, n, m, (int)k, (int)dim, scale,
                                                                                 key_row, key_col, g, d_a, d_b)));
  RS_LAUNCH_CHECK();
  return RS_OK;
}
