// N4: ensemble merge of two retrieval lists.
//
// Reference: tower_code/mined_inference.py, evaluate_weighted_score_ensemble :1110-1189 and evaluate_rrf_ensemble
// :1337-1411.  Per user the two models' top-M lists are concatenated (2M candidates, an item both models rank appears
// twice), both models re-score every candidate, the scores are min-max normalised per user (or turned into
// reciprocal ranks), and for every blend weight alpha:  final = alpha * n1 + (1 - alpha) * n2, torch.topk(max_k + 20),
// back to global ids, then -- on the HOST, per user, in Python -- np.unique(return_index) to drop the duplicates while
// keeping the ranking order (:1182-1183, :1406-1407).
//
// Here one CTA per user does all of it in shared memory for every alpha: normalise, blend, sort, cut, de-duplicate.
// A candidate list holds an item at most twice (once per model's top-M).  The twin of every candidate is found once
// per user (one sort by id); after the blend's sort an entry is dropped iff its twin is ranked before it -- exactly
// what np.unique(return_index) + sort of the first occurrences keeps.  fp32 arithmetic mirrors torch's op by op
// (separately rounded multiply / add / divide), so with the same re-scored inputs the blended scores are bit-identical
// to the reference's and the ranking can only differ inside exact ties between DIFFERENT items (torch.topk leaves
// their order unspecified; here: candidate position ascending, which is also how a stable descending sort ranks the
// two equal-score copies of an item in the RRF variant).
#include "common.cuh"
#include "../../include/rs_twotower.h"

namespace rs {

#define ENS_THREADS 256
#define ENS_MAX_P 2048
#define ENS_MAX_ALPHAS 32

struct EnsAlphas { float a[ENS_MAX_ALPHAS]; float b[ENS_MAX_ALPHAS]; };   // alpha, (1 - alpha) (rounded from double)

// ascending order of the returned key == descending order of the score
__device__ __forceinline__ uint32_t desc_key(float f) {
  uint32_t u = __float_as_uint(f);
  u ^= (u & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u;      // ascending-orderable
  return ~u;
}

__device__ __forceinline__ void bitonic_sort(unsigned long long* keys, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n_pow2 >> 1); t += ENS_THREADS) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));      // lower index of the pair
        const int p = i | j;
        const bool up = (i & k) == 0;
        const unsigned long long a = keys[i], b = keys[p];
        if ((a > b) == up) { keys[i] = b; keys[p] = a; }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ float block_reduce_minmax(float v, bool want_max, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = want_max ? fmaxf(v, w) : fminf(v, w);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int k = 1; k < ENS_THREADS / 32; ++k) r = want_max ? fmaxf(r, red[k]) : fminf(r, red[k]);
  return r;
}

// mode 0: min-max weighted sum; mode 1: weighted reciprocal-rank fusion
__global__ void __launch_bounds__(ENS_THREADS) ensemble_merge_kernel(
    const int64_t* __restrict__ cand_ids, const float* __restrict__ s1, const float* __restrict__ s2, int64_t P,
    int p_pow2, int mode, float k_rrf, EnsAlphas al, int n_alpha, int64_t k_sel, int64_t n_users,
    int64_t* __restrict__ out_ids, int32_t* __restrict__ out_cnt, float* __restrict__ out_n1, float* __restrict__ out_n2) {
  extern __shared__ unsigned long long ens_smem[];
  unsigned long long* keys = ens_smem;                               // [p_pow2]
  float* n1 = reinterpret_cast<float*>(keys + p_pow2);               // [P]  normalised score / reciprocal rank, model 1
  float* n2 = n1 + P;                                                // [P]
  uint32_t* ids = reinterpret_cast<uint32_t*>(n2 + P);               // [P]  global ids (low 32 bits)
  int* twin = reinterpret_cast<int*>(ids + P);                       // [P]  position of the other copy of the item, -1
  int* rank_of = twin + P;                                           // [P]  rank of every candidate in the current blend
  __shared__ float red[ENS_THREADS / 32];
  __shared__ int wsum[ENS_THREADS / 32];
  const int64_t u = blockIdx.x;
  const int64_t base = u * P;
  for (int i = threadIdx.x; i < P; i += ENS_THREADS) { ids[i] = (uint32_t)cand_ids[base + i]; twin[i] = -1; }
  __syncthreads();
  // twins: sort (id, position); equal neighbours are the two copies of one item
  for (int i = threadIdx.x; i < p_pow2; i += ENS_THREADS)
    keys[i] = i < P ? (((unsigned long long)ids[i] << 32) | (unsigned)i) : ~0ull;
  __syncthreads();
  bitonic_sort(keys, p_pow2);
  for (int j = threadIdx.x; j + 1 < P; j += ENS_THREADS) {
    const unsigned long long a = keys[j], b = keys[j + 1];
    if ((a >> 32) == (b >> 32)) {
      const int pa = (int)(a & 0xFFFFFFFFu), pb = (int)(b & 0xFFFFFFFFu);
      twin[pa] = pb;
      twin[pb] = pa;
    }
  }
  __syncthreads();

  if (mode == 0) {
    float lo1 = INFINITY, hi1 = -INFINITY, lo2 = INFINITY, hi2 = -INFINITY;
    for (int i = threadIdx.x; i < P; i += ENS_THREADS) {
      const float a = s1[base + i], b = s2[base + i];
      lo1 = fminf(lo1, a); hi1 = fmaxf(hi1, a); lo2 = fminf(lo2, b); hi2 = fmaxf(hi2, b);
    }
    lo1 = block_reduce_minmax(lo1, false, red); hi1 = block_reduce_minmax(hi1, true, red);
    lo2 = block_reduce_minmax(lo2, false, red); hi2 = block_reduce_minmax(hi2, true, red);
    // (tensor - min) / (max - min + 1e-9), each op rounded to fp32 as torch does (:1139-1142)
    const float d1 = __fadd_rn(__fsub_rn(hi1, lo1), 1e-9f), d2 = __fadd_rn(__fsub_rn(hi2, lo2), 1e-9f);
    for (int i = threadIdx.x; i < P; i += ENS_THREADS) {
      n1[i] = __fdiv_rn(__fsub_rn(s1[base + i], lo1), d1);
      n2[i] = __fdiv_rn(__fsub_rn(s2[base + i], lo2), d2);
    }
  } else {
    // rank of every candidate under each model (descending score; ties by position), then 1 / (k_rrf + rank + 1)
    for (int m = 0; m < 2; ++m) {
      const float* s = m ? s2 : s1;
      float* dst = m ? n2 : n1;
      for (int i = threadIdx.x; i < p_pow2; i += ENS_THREADS)
        keys[i] = i < P ? (((unsigned long long)desc_key(s[base + i]) << 32) | (unsigned)i) : ~0ull;
      __syncthreads();
      bitonic_sort(keys, p_pow2);
      for (int r = threadIdx.x; r < P; r += ENS_THREADS) {
        const int pos = (int)(keys[r] & 0xFFFFFFFFu);
        dst[pos] = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(k_rrf, (float)r), 1.0f));
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (out_n1) for (int i = threadIdx.x; i < P; i += ENS_THREADS) { out_n1[base + i] = n1[i]; out_n2[base + i] = n2[i]; }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int ai = 0; ai < n_alpha; ++ai) {
    const float a = al.a[ai], b = al.b[ai];
    for (int i = threadIdx.x; i < p_pow2; i += ENS_THREADS) {
      unsigned long long k = ~0ull;
      if (i < P) {
        const float f = __fadd_rn(__fmul_rn(a, n1[i]), __fmul_rn(b, n2[i]));       // alpha * n1 + (1 - alpha) * n2
        k = ((unsigned long long)desc_key(f) << 32) | (unsigned)i;
      }
      keys[i] = k;
    }
    __syncthreads();
    bitonic_sort(keys, p_pow2);
    for (int r = threadIdx.x; r < P; r += ENS_THREADS) rank_of[(int)(keys[r] & 0xFFFFFFFFu)] = r;
    __syncthreads();
    // first k_sel of the ranking; an entry whose twin is ranked before it is a duplicate: dropped, order kept
    int64_t* dst = out_ids + ((int64_t)ai * n_users + u) * k_sel;
    int running = 0;
    for (int c0 = 0; c0 < k_sel; c0 += ENS_THREADS) {
      const int i = c0 + threadIdx.x;
      int keep = 0;
      uint32_t id = 0;
      if (i < k_sel && i < P) {
        const int pos = (int)(keys[i] & 0xFFFFFFFFu);
        const int t = twin[pos];
        id = ids[pos];
        keep = !(t >= 0 && rank_of[t] < i);
      }
      // block-wide exclusive scan of `keep`
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      const int in_warp = __popc(bal & ((1u << lane) - 1u));
      if (lane == 0) wsum[warp] = __popc(bal);
      __syncthreads();
      int before = 0, total = 0;
      for (int w = 0; w < ENS_THREADS / 32; ++w) { if (w < warp) before += wsum[w]; total += wsum[w]; }
      if (keep) dst[running + before + in_warp] = (int64_t)id;
      running += total;
      __syncthreads();
    }
    for (int i = running + threadIdx.x; i < k_sel; i += ENS_THREADS) dst[i] = -1;
    if (threadIdx.x == 0) out_cnt[(int64_t)ai * n_users + u] = running;
    __syncthreads();
  }
}

}  // namespace rs

using namespace rs;

extern "C" int rs_ensemble_merge(const int64_t* cand_ids, const float* s1, const float* s2, int64_t n_users, int64_t P,
                                 int mode, float k_rrf, const double* alphas, int n_alpha, int64_t k_sel,
                                 int64_t* out_ids, int32_t* out_cnt, float* out_n1, float* out_n2, void* stream) {
  if (!cand_ids || !s1 || !s2 || !alphas || !out_ids || !out_cnt || n_users < 0 || P <= 0 || k_sel <= 0) return RS_ERR_BAD_ARG;
  if ((out_n1 != nullptr) != (out_n2 != nullptr)) return RS_ERR_BAD_ARG;
  if (P > ENS_MAX_P || n_alpha <= 0 || n_alpha > ENS_MAX_ALPHAS || k_sel > P || (mode != 0 && mode != 1)) return RS_ERR_UNSUPPORTED;
  if (n_users == 0) return RS_OK;
  int p2 = 1;
  while (p2 < P) p2 <<= 1;
  EnsAlphas al;
  for (int i = 0; i < n_alpha; ++i) {
    al.a[i] = (float)alphas[i];                       // torch rounds the python scalar to the tensor's dtype
    al.b[i] = (float)(1.0 - alphas[i]);               // (1.0 - alpha) is computed in double by python, then rounded
  }
  const size_t smem = (size_t)p2 * 8 + (size_t)P * 20;
  cudaError_t e = cudaFuncSetAttribute(ensemble_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  ensemble_merge_kernel<<<(int)n_users, ENS_THREADS, smem, (cudaStream_t)stream>>>(
      cand_ids, s1, s2, P, p2, mode, k_rrf, al, n_alpha, k_sel, n_users, out_ids, out_cnt, out_n1, out_n2);
  RS_LAUNCH_CHECK();
  return RS_OK;
}
