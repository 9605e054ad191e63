// R1 on the tensor cores: candidate pass (tcgen05, bf16 operands) + exact fp32 re-scoring.
//
// Reference: `scores = U @ I.T` in fp32 (TF32 off), `torch.topk(scores, k)` -- tower_code/v1_usertower_train.py:672-675,
// mined_inference.py:901-909, 1103-1108, 1536-1543.  The ids must be the fp32 ranking's, so a reduced-precision
// contraction can only be a FILTER:
//   1. operands rounded to bf16 (round-to-nearest: |a.b - a16.b16| <= 2^-8 (1 + 2^-9) |a| |b| by Cauchy-Schwarz);
//   2. candidate pass: S16 = U16 @ I16^T tile by tile on tcgen05 (fp32 accumulation in TMEM); every thread of the
//      epilogue owns one user row and keeps the K best per-chunk maxima it has seen -- their k-th largest is a lower
//      bound `thr` of the row's final k-th best approximate score -- and appends every (score, item) with
//      score > thr - 2 eps_row to the row's candidate list (eps_row = the bound above for this row's norm and the largest
//      item norm).  Any item of the exact top-k has approx >= exact - eps >= (k-th best exact) - eps >= (k-th best approx)
//      - 2 eps >= thr - 2 eps: it is in the list;
//   3. refine: per row, final threshold from the lists' K-best arrays, the surviving candidates are re-scored with an
//      exact fp32 dot product and the top k by (score desc, id asc) are written out.
// If a list overflows its capacity (adversarial score distributions), a device flag makes the exact fp32 kernel of
// topk.cu -- launched behind this path, returning at once otherwise -- produce the result instead: no host decision.
// The [n_users, n_items] score matrix is never written; candidates are a few hundred entries per row.
#include "tcgen05.cuh"
#include "../../include/rs_twotower.h"

namespace rs {

#define TT_BM 128
#define TT_BN 128
#define TT_K 128
#define TT_STAGES 4
#define TT_NWG 3
#define TT_THREADS (64 + 128 * TT_NWG)
#define TT_TILE_BYTES (TT_BN * TT_K * 2)
#define TT_CAP 256                       // entries per (split, warpgroup, row) candidate list
#define TT_KEEP 256                      // candidates that may survive the final threshold per row

struct TtShared {
  uint64_t full[TT_STAGES], empty[TT_STAGES];
  uint64_t a_full[2], a_empty[2];
  uint64_t tmem_full[TT_NWG], tmem_empty[TT_NWG];
  uint32_t tmem_base;
  uint32_t pad[3];
  float thr[TT_BM];                      // best known lower bound of each row's k-th best (shared by the warpgroups)
};

struct TtParams {
  int64_t n_users, n_items;
  int k, mask0;
  int tiles_per_item, nsplit, row_blocks, col_tiles;
  uint32_t idesc;
  const float* unorm;                    // [n_users]
  const float* imax;                     // device scalar: largest item norm
  float2* cand;                          // [nsplit][TT_NWG][n_users][TT_CAP] (score, id as int bits)
  int* cand_cnt;                         // [nsplit][TT_NWG][n_users]
  float* tops;                           // [nsplit][TT_NWG][n_users][KT] per-list best chunk maxima (descending)
  int* overflow;                         // device flag
};

// x fp32 [n, 128] -> bf16 copy + row norms (+ atomic max of the norms, as int bits: norms are >= 0)
__global__ void __launch_bounds__(256) tt_prep_kernel(const float* __restrict__ x, int64_t n, uint16_t* __restrict__ x16,
                                                      float* __restrict__ norm, int* __restrict__ max_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float local_max = 0.f;
  for (int64_t r = warp; r < n; r += nwarps) {
    const float4 v = ld_stream_f4(x + r * TT_K + 4 * lane);
    uint2 w;
    w.x = pack_bf16(v.x, v.y);
    w.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(x16 + r * TT_K + 4 * lane) = w;
    const float s = sqrtf(warp_sum(dot4(v, v)));
    if (lane == 0 && norm) norm[r] = s;
    local_max = fmaxf(local_max, s);
  }
  if (max_bits && lane == 0) atomicMax(max_bits, __float_as_int(local_max));
}

__device__ __forceinline__ void tt_item_coords(const TtParams& p, int item, int& rb, int& sp, int& lo, int& hi) {
  rb = item / p.nsplit;
  sp = item % p.nsplit;
  lo = sp * p.tiles_per_item;
  hi = min(lo + p.tiles_per_item, p.col_tiles);
}

template <int KT>
__global__ void __launch_bounds__(TT_THREADS, 1)
tt_cand_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TtParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                                        // 2 x 32 KB user blocks
  uint8_t* sB = base + 2 * TT_TILE_BYTES;                    // TT_STAGES x 32 KB item tiles
  TtShared& sh = *reinterpret_cast<TtShared*>(base + (2 + TT_STAGES) * TT_TILE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int i = 0; i < TT_STAGES; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sh.a_full[i], 1); mbar_init(&sh.a_empty[i], 1); }
    for (int i = 0; i < TT_NWG; ++i) { mbar_init(&sh.tmem_full[i], 1); mbar_init(&sh.tmem_empty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&sh.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_base;
  const int n_items_w = p.row_blocks * p.nsplit;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, item_n = 0;
      for (int item = blockIdx.x; item < n_items_w; item += gridDim.x, ++item_n) {
        int rb, sp, lo, hi;
        tt_item_coords(p, item, rb, sp, lo, hi);
        const uint32_t ab = item_n & 1;
        mbar_wait(&sh.a_empty[ab], ((item_n >> 1) & 1) ^ 1);
        mbar_expect_tx(&sh.a_full[ab], TT_TILE_BYTES);
        tma_load_2d(&mapA, &sh.a_full[ab], sA + ab * TT_TILE_BYTES, 0, rb * TT_BM);
        tma_load_2d(&mapA, &sh.a_full[ab], sA + ab * TT_TILE_BYTES + TC_BOX_BYTES, 64, rb * TT_BM);
        for (int ct = lo; ct < hi; ++ct, ++it) {
          const uint32_t s = it % TT_STAGES, ph = (it / TT_STAGES) & 1;
          mbar_wait(&sh.empty[s], ph ^ 1);
          mbar_expect_tx(&sh.full[s], TT_TILE_BYTES);
          tma_load_2d(&mapB, &sh.full[s], sB + s * TT_TILE_BYTES, 0, ct * TT_BN);
          tma_load_2d(&mapB, &sh.full[s], sB + s * TT_TILE_BYTES + TC_BOX_BYTES, 64, ct * TT_BN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, item_n = 0;
      for (int item = blockIdx.x; item < n_items_w; item += gridDim.x, ++item_n) {
        int rb, sp, lo, hi;
        tt_item_coords(p, item, rb, sp, lo, hi);
        const uint32_t ab = item_n & 1;
        mbar_wait(&sh.a_full[ab], (item_n >> 1) & 1);
        const uint64_t adesc = desc_kmajor(smem_u32(sA + ab * TT_TILE_BYTES));
        for (int ct = lo; ct < hi; ++ct, ++it) {
          const uint32_t s = it % TT_STAGES, ph = (it / TT_STAGES) & 1, g = it % TT_NWG, ng = it / TT_NWG;
          mbar_wait(&sh.tmem_empty[g], (ng & 1) ^ 1);
          mbar_wait(&sh.full[s], ph);
          tc_fence_after();
          const uint64_t bdesc = desc_kmajor(smem_u32(sB + s * TT_TILE_BYTES));
#pragma unroll
          for (int kk = 0; kk < TT_K / 16; ++kk) {
            const uint64_t off = (uint64_t)(((kk >> 2) * TC_BOX_BYTES + (kk & 3) * 32) >> 4);
            umma_f16(tmem_base + g * TT_BN, adesc + off, bdesc + off, p.idesc, kk > 0 ? 1u : 0u);
          }
          umma_commit(&sh.tmem_full[g]);
          umma_commit(&sh.empty[s]);
        }
        umma_commit(&sh.a_empty[ab]);
      }
    }
  } else {
    const int wg = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int rloc = quarter * 32 + lane;
    const float imax = __ldg(p.imax);
    uint32_t it = 0, nuse = 0;
    for (int item = blockIdx.x; item < n_items_w; item += gridDim.x) {
      int rb, sp, lo, hi;
      tt_item_coords(p, item, rb, sp, lo, hi);
      const int64_t row = (int64_t)rb * TT_BM + rloc;
      const bool row_ok = row < p.n_users;
      // |approx - exact| <= 2^-8 (1 + 2^-9) |u| |i|max = 0.003914 |u| |i|max for every item of this row (see the header);
      // 0.00394 leaves 6e-3 relative for the fp32 accumulation noise of the 128-term sums
      const float eps = row_ok ? (0.00394f * __ldg(p.unorm + row) * imax) : 0.f;
      const float margin = 2.f * eps;
      float top[KT];
#pragma unroll
      for (int q = 0; q < KT; ++q) top[q] = -INFINITY;
      int cnt = 0;
      bool first = true;
      const int64_t list = ((int64_t)sp * TT_NWG + wg) * p.n_users + row;
      float2* my = p.cand + list * TT_CAP;
      // all warpgroups are done with the previous item's thresholds before they are reset
      named_bar_sync(1, 128 * TT_NWG);
      if (wg == 0) sh.thr[rloc] = -INFINITY;
      named_bar_sync(1, 128 * TT_NWG);
      volatile float* shthr = sh.thr + rloc;
      for (int ct = lo; ct < hi; ++ct, ++it) {
        if ((int)(it % TT_NWG) != wg) continue;
        const int64_t c0 = (int64_t)ct * TT_BN;
        const bool edge = (c0 + TT_BN > p.n_items) || (p.mask0 && c0 == 0);
        mbar_wait(&sh.tmem_full[wg], nuse & 1);
        tc_fence_after();
        const uint32_t tt = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(wg * TT_BN);
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t r[32];
          tmem_ld32(tt + (uint32_t)(ch * 32), r);
          tmem_ld_wait();
          if (edge) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int64_t col = c0 + ch * 32 + j;
              if (col >= p.n_items || (p.mask0 && col == 0)) r[j] = 0xff800000u;      // -inf
            }
          }
          float mx = __uint_as_float(r[0]);
#pragma unroll
          for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
          // threshold: k-th largest of a set of DISTINCT items' scores seen so far -- every element of the list's first
          // chunk (so that the bound is finite after 32 >= k elements instead of after k chunks), then one element
          // per chunk (its maximum) -- this list's, or another warpgroup's of the same row when that is tighter
          bool changed = false;
          if (first) {
            first = false;
            changed = true;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = __uint_as_float(r[j]);
              if (v > top[KT - 1]) {
                top[KT - 1] = v;
#pragma unroll
                for (int q = KT - 1; q > 0; --q) {
                  const float a = top[q - 1], b = top[q];
                  top[q - 1] = fmaxf(a, b);
                  top[q] = fminf(a, b);
                }
              }
            }
          } else if (mx > top[KT - 1]) {
            changed = true;
            top[KT - 1] = mx;
#pragma unroll
            for (int q = KT - 1; q > 0; --q) {
              const float a = top[q - 1], b = top[q];
              top[q - 1] = fmaxf(a, b);
              top[q] = fminf(a, b);
            }
          }
          if (changed) {
            float mine = -INFINITY;
#pragma unroll
            for (int q = 0; q < KT; ++q) if (q == p.k - 1) mine = top[q];
            if (mine > *shthr) *shthr = mine;            // benign race: any stored value is a valid lower bound
          }
          const float lo_thr = *shthr - margin;
          if (mx > lo_thr && row_ok) {
            const int colbase = (int)(c0 + ch * 32);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = __uint_as_float(r[j]);
              if (v > lo_thr) {
                if (cnt < TT_CAP) my[cnt] = make_float2(v, __int_as_float(colbase + j));
                ++cnt;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&sh.tmem_empty[wg]);
        ++nuse;
      }
      if (row_ok) {
        p.cand_cnt[list] = min(cnt, TT_CAP);
        if (cnt > TT_CAP) atomicExch(p.overflow, 1);
#pragma unroll
        for (int q = 0; q < KT; ++q) p.tops[list * KT + q] = top[q];
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

__device__ __forceinline__ bool tt_better(float v, int id, float wv, int wid) { return (v > wv) || (v == wv && id < wid); }

// one warp per user row: final threshold, exact fp32 re-scoring of the survivors, top k by (score desc, id asc)
template <int KT>
__global__ void __launch_bounds__(256) tt_refine_kernel(const float* __restrict__ users, const float* __restrict__ items,
                                                        TtParams p, int n_lists, int64_t* __restrict__ out_ids,
                                                        float* __restrict__ out_scores) {
  __shared__ float s_sc[8][TT_KEEP];
  __shared__ int s_id[8][TT_KEEP];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * 8 + w, nwarps = (int64_t)gridDim.x * 8;
  const float imax = __ldg(p.imax);
  for (int64_t row = warp0; row < p.n_users; row += nwarps) {
    // ---- k-th largest of the union of the lists' best chunk maxima = lower bound of the k-th best approximate score
    const int nt = n_lists * KT;                    // <= 32 lists x 32
    float thr = -INFINITY;
    {
      // every lane holds the values v[q] = tops[(q*32 + lane)]; rank of a value = number of strictly greater values
      // (ties broken by position): the value of rank k-1 is the threshold
      for (int a0 = 0; a0 < nt; a0 += 32) {
        const int a = a0 + lane;
        float va = -INFINITY;
        if (a < nt) va = p.tops[(((int64_t)(a / KT)) * p.n_users + row) * KT + (a % KT)];
        int rank = 0;
        for (int b0 = 0; b0 < nt; b0 += 32) {
          const int bidx = b0 + lane;
          float vb = -INFINITY;
          if (bidx < nt) vb = p.tops[(((int64_t)(bidx / KT)) * p.n_users + row) * KT + (bidx % KT)];
          for (int l = 0; l < 32; ++l) {
            const float o = __shfl_sync(0xffffffffu, vb, l);
            const int oi = b0 + l;
            rank += (oi < nt && (o > va || (o == va && oi < a))) ? 1 : 0;
          }
        }
        const unsigned hit = __ballot_sync(0xffffffffu, a < nt && rank == p.k - 1);
        if (hit) thr = __shfl_sync(0xffffffffu, va, __ffs(hit) - 1);
      }
    }
    const float eps = 0.00394f * __ldg(p.unorm + row) * imax;
    const float keep_thr = thr - 2.f * eps;
    // ---- survivors -> shared memory (compacted with ballots)
    int n_keep = 0;
    for (int li = 0; li < n_lists; ++li) {
      const int64_t list = (int64_t)li * p.n_users + row;
      const int c = p.cand_cnt[list];
      const float2* src = p.cand + list * TT_CAP;
      for (int e0 = 0; e0 < c; e0 += 32) {
        const int e = e0 + lane;
        float2 v = make_float2(-INFINITY, 0.f);
        if (e < c) v = src[e];
        const bool keep = e < c && v.x >= keep_thr;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int pos = n_keep + __popc(bal & ((1u << lane) - 1u));
        if (keep && pos < TT_KEEP) { s_sc[w][pos] = v.x; s_id[w][pos] = __float_as_int(v.y); }
        n_keep += __popc(bal);
      }
    }
    if (n_keep > TT_KEEP) { if (lane == 0) atomicExch(p.overflow, 1); n_keep = TT_KEEP; }
    __syncwarp();
    // ---- exact scores: fp32 dot products, the warp shares the user row
    const float4 u = __ldg(reinterpret_cast<const float4*>(users + row * TT_K) + lane);
    for (int e = 0; e < n_keep; ++e) {
      const int id = s_id[w][e];
      const float4 x = __ldg(reinterpret_cast<const float4*>(items + (int64_t)id * TT_K) + lane);
      const float d = warp_sum(dot4(u, x));
      if (lane == 0) s_sc[w][e] = d;
    }
    __syncwarp();
    // ---- k rounds of arg-max over the survivors
    for (int r = 0; r < p.k; ++r) {
      float bv = -INFINITY;
      int bi = 0x7fffffff, be = -1;
      for (int e = lane; e < n_keep; e += 32) {
        const float v = s_sc[w][e];
        const int id = s_id[w][e];
        if (id >= 0 && tt_better(v, id, bv, bi)) { bv = v; bi = id; be = e; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const int oe = __shfl_xor_sync(0xffffffffu, be, o);
        if (oe >= 0 && (be < 0 || tt_better(ov, oi, bv, bi))) { bv = ov; bi = oi; be = oe; }
      }
      if (lane == 0) {
        out_ids[row * p.k + r] = be >= 0 ? (int64_t)bi : -1;
        out_scores[row * p.k + r] = be >= 0 ? bv : -INFINITY;
        if (be >= 0) s_id[w][be] = -1;                 // taken
        else atomicExch(p.overflow, 1);                // fewer than k survivors: cannot happen unless a list overflowed
      }
      __syncwarp();
    }
  }
}

}  // namespace rs

using namespace rs;

// exact fp32 kernel of topk.cu with a device-side condition (runs only when *run_if != 0)
int rs_retrieve_topk_cond(const float* users, int64_t n_users, const float* items, int64_t n_items, int64_t dim,
                          int64_t k, int mask_index0, int64_t* out_ids, float* out_scores, void* workspace,
                          size_t workspace_bytes, const int* run_if, void* stream);

static inline size_t tt_al(size_t x) { return (x + 255) & ~(size_t)255; }

struct TtPlan { int row_blocks, col_tiles, tiles_per_item, nsplit, grid, kt; };
static int tt_plan(int64_t n_users, int64_t n_items, int64_t k, TtPlan* pl) {
  if (k <= 0 || k > 32 || k > n_items) return RS_ERR_UNSUPPORTED;
  pl->kt = k <= 16 ? 16 : 32;
  pl->row_blocks = (int)((n_users + TT_BM - 1) / TT_BM);
  pl->col_tiles = (int)((n_items + TT_BN - 1) / TT_BN);
  // column splits: fill the SMs when there are few row blocks; every split must still see well over k items
  int ns = 1;
  if (pl->row_blocks < RS_NUM_SMS) ns = (RS_NUM_SMS + pl->row_blocks - 1) / pl->row_blocks;
  const int max_by_k = pl->col_tiles / 8 > 0 ? pl->col_tiles / 8 : 1;
  if (ns > max_by_k) ns = max_by_k;
  if (ns > 10) ns = 10;                               // TT_NWG * nsplit lists per row <= 32 (refine kernel)
  pl->tiles_per_item = (pl->col_tiles + ns - 1) / ns;
  pl->nsplit = (pl->col_tiles + pl->tiles_per_item - 1) / pl->tiles_per_item;
  const int64_t items_w = (int64_t)pl->row_blocks * pl->nsplit;
  pl->grid = (int)(items_w < RS_NUM_SMS ? items_w : RS_NUM_SMS);
  return RS_OK;
}

extern "C" size_t rs_topk_workspace_bytes(int64_t n_users, int64_t n_items, int64_t dim, int64_t k);

extern "C" size_t rs_retrieve_topk_tc_workspace_bytes(int64_t n_users, int64_t n_items, int64_t dim, int64_t k) {
  TtPlan pl;
  if (dim != TT_K || tt_plan(n_users, n_items, k, &pl) != RS_OK) return 256;
  const size_t lists = (size_t)pl.nsplit * TT_NWG * n_users;
  return tt_al((size_t)n_users * TT_K * 2) + tt_al((size_t)n_items * TT_K * 2) + tt_al((size_t)n_users * 4) + 256 +
         tt_al(lists * TT_CAP * sizeof(float2)) + tt_al(lists * 4) + tt_al(lists * pl.kt * 4) +
         tt_al(rs_topk_workspace_bytes(n_users, n_items, dim, k)) + 1024;
}

extern "C" int rs_retrieve_topk_tc(const float* users, int64_t n_users, const float* items, int64_t n_items, int64_t dim,
                                   int64_t k, int mask_index0, int64_t* out_ids, float* out_scores, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (n_users == 0) return RS_OK;
  if (!users || !items || !out_ids || !out_scores || !workspace) return RS_ERR_BAD_ARG;
  if (dim != TT_K) return RS_ERR_UNSUPPORTED;
  TtPlan pl;
  int rc = tt_plan(n_users, n_items, k, &pl);
  if (rc != RS_OK) return rc;
  if (workspace_bytes < rs_retrieve_topk_tc_workspace_bytes(n_users, n_items, dim, k)) return RS_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)workspace;
  uint16_t* u16 = (uint16_t*)w; w += tt_al((size_t)n_users * TT_K * 2);
  uint16_t* i16 = (uint16_t*)w; w += tt_al((size_t)n_items * TT_K * 2);
  float* unorm = (float*)w; w += tt_al((size_t)n_users * 4);
  int* scal = (int*)w; w += 256;                         // [0] max item norm (float bits), [1] overflow flag
  const size_t lists = (size_t)pl.nsplit * TT_NWG * n_users;
  float2* cand = (float2*)w; w += tt_al(lists * TT_CAP * sizeof(float2));
  int* cand_cnt = (int*)w; w += tt_al(lists * 4);
  float* tops = (float*)w; w += tt_al(lists * pl.kt * 4);
  void* fb_ws = w;
  const size_t fb_bytes = rs_topk_workspace_bytes(n_users, n_items, dim, k);
  cudaError_t e = cudaMemsetAsync(scal, 0, 256, st);
  if (e != cudaSuccess) return (int)e;
  tt_prep_kernel<<<grid_for_warps(n_users, 8, 8), 256, 0, st>>>(users, n_users, u16, unorm, nullptr);
  RS_LAUNCH_CHECK();
  tt_prep_kernel<<<grid_for_warps(n_items, 8, 8), 256, 0, st>>>(items, n_items, i16, nullptr, scal);
  RS_LAUNCH_CHECK();
  CUtensorMap mapA, mapB;
  if ((rc = make_map(&mapA, u16, n_users, RS_BF16)) != RS_OK) return rc;
  if ((rc = make_map(&mapB, i16, n_items, RS_BF16)) != RS_OK) return rc;
  TtParams p = {};
  p.n_users = n_users; p.n_items = n_items; p.k = (int)k; p.mask0 = mask_index0;
  p.tiles_per_item = pl.tiles_per_item; p.nsplit = pl.nsplit; p.row_blocks = pl.row_blocks; p.col_tiles = pl.col_tiles;
  // kind::f16 instruction descriptor: bf16 operands (both K-major), fp32 accumulate, M = N = 128
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TT_BN >> 3) << 17) | ((uint32_t)(TT_BM >> 4) << 24);
  p.unorm = unorm; p.imax = (const float*)scal; p.cand = cand; p.cand_cnt = cand_cnt; p.tops = tops; p.overflow = scal + 1;
  const size_t smem = (size_t)(2 + TT_STAGES) * TT_TILE_BYTES + sizeof(TtShared) + 1024;
  const int n_lists = pl.nsplit * TT_NWG;
  const int rgrid = grid_for_warps(n_users, 8, 8);
#define LAUNCH_TT(KT)                                                                                          \
  do {                                                                                                         \
    e = cudaFuncSetAttribute(tt_cand_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    if (e != cudaSuccess) return (int)e;                                                                       \
    tt_cand_kernel<KT><<<pl.grid, TT_THREADS, smem, st>>>(mapA, mapB, p);                                      \
    RS_LAUNCH_CHECK();                                                                                         \
    tt_refine_kernel<KT><<<rgrid, 256, 0, st>>>(users, items, p, n_lists, out_ids, out_scores);                \
    RS_LAUNCH_CHECK();                                                                                         \
  } while (0)
  if (pl.kt == 16) LAUNCH_TT(16); else LAUNCH_TT(32);
  // exact fallback, device-conditional on the overflow flag
  return rs_retrieve_topk_cond(users, n_users, items, n_items, dim, k, mask_index0, out_ids, out_scores, fb_ws, fb_bytes,
                               scal + 1, stream);
}
