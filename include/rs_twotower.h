/*
 * rs_twotower.h -- C ABI of the B200 (sm_100a) two-tower hot-path library
 * (librs_twotower.so).
 *
 * The reference (DotBlossom/LLM-driven_content-based-feature_recommendation_system)
 * is pure Python/PyTorch and has NO plugin registry, operator table or FFI
 * (SURVEY.md 8b): its boundary for this path is the nn.Module.forward / loss
 * function surface.  The host mirror of that surface lives in the Python
 * package next to this library; underneath it every arithmetic step crosses
 * THIS boundary.  Each entry point below names the reference lines (relative
 * to the reference checkout) whose ATen calls it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the comment says "host".  The caller owns every buffer; the library never
 *     allocates or frees device memory.  Scratch is passed in as `workspace`
 *     (size from the matching *_workspace_bytes()).
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered,
 *     no host synchronisation happens inside any call.
 *   - ids are int64 (the reference's torch.long).  dtype codes: RS_F32/F16/BF16.
 *   - return value: 0 on success, otherwise a cudaError_t value or one of the
 *     RS_ERR_* codes (>= 10000); rs_error_string() describes it.  The Python
 *     mirror raises RuntimeError, like the reference's PyTorch calls would.
 *   - `oob_flag` (int*, may be NULL): set to 1 by a kernel that meets an id
 *     outside [0, rows).  Such a row reads as zeros / is not accumulated.  The
 *     reference raises IndexError (CPU) or a device assert (CUDA) there; the
 *     host mirror turns the flag into IndexError when asked to check.
 *   - thread-safety: re-entrant; no global mutable state.
 */
#ifndef RS_TWOTOWER_H_
#define RS_TWOTOWER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_F32 0
#define RS_F16 1
#define RS_BF16 2

#define RS_ERR_BAD_ARG 10001
#define RS_ERR_UNSUPPORTED 10002
#define RS_ERR_WORKSPACE 10003

#define RS_MAX_TABLES 8

int rs_abi_version(void);
const char* rs_error_string(int code);
/* number of kernels this library has launched in this process (statistics for bench.py's gpu_launches) */
unsigned long long rs_launch_count(void);
void rs_count_launches(int n);

/* ------------------------------------------------------------------ gathers */

/* out[i,:] = table[min(ids[i], clamp_max) , :]   (clamp_max < 0: no clamp)
 * Replaces aten::embedding / index at: tower_code/v1_usertower_train.py:760
 * (pretrained_lookup[item_ids]), tower_code/v1_refine_usertower.py:833
 * (item_tower_emb[target_ids]), tower_code/mined_inference.py:670,687-688,705
 * (gnn_user_emb / item_content_emb / gnn_item_emb / channel_emb) and :695
 * (time_emb(seq_deltas.clamp(max=1000))).  dim % 4 == 0 (dim % 8 for 16-bit). */
int rs_gather_rows(const void* table, int table_dtype, int64_t rows, int64_t dim,
                   const int64_t* ids, int64_t n, int64_t clamp_max,
                   void* out, int out_dtype, int* oob_flag, void* stream);

/* d_table[ids[i],:] += scale * d_out[i,:]  for ids[i] != padding_idx  (padding_idx < 0: none).
 * fp32 vector atomics (red.global.add.v4.f32).  d_table is fp32 and is NOT cleared.
 * Replaces aten::embedding_dense_backward (autograd of the gathers above and of
 * item_tower.py:239,270 -> BERT word_embeddings [30522,768]). */
int rs_scatter_add_rows(const void* d_out, int d_out_dtype, const int64_t* ids, int64_t n, int64_t dim,
                        int64_t rows, int64_t padding_idx, int64_t clamp_max, float scale,
                        float* d_table, int* oob_flag, void* stream);

/* Deterministic variant: sort (id, position) by id with a stable LSD radix sort, then
 * segment-reduce in position order.  d_table must be zero-filled (rows not hit stay 0).
 * `sorted_*` may be NULL (then built inside `workspace`), or a cache filled by
 * rs_sort_ids() for the same `ids`. */
size_t rs_sort_ids_workspace_bytes(int64_t n);
int rs_sort_ids(const int64_t* ids, int64_t n, int64_t rows, int64_t clamp_max,
                int32_t* sorted_ids, int32_t* sorted_pos, void* workspace, size_t workspace_bytes,
                int* oob_flag, void* stream);
size_t rs_segment_reduce_workspace_bytes(int64_t n, int64_t dim);
int rs_segment_reduce_rows(const void* d_out, int d_out_dtype, const int32_t* sorted_ids,
                           const int32_t* sorted_pos, int64_t n, int64_t dim, int64_t rows,
                           int64_t padding_idx, const float* scale_dev /* device scalar or NULL (=1) */,
                           const float* dot_table /* optional [rows,dim]: also emit sum_i <dot_table[id_i], d_out[i]> */,
                           float* d_table, float* dot_out /* device scalar, accumulated, or NULL */,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------ U1: SASRec sequence front */

/* out[p,:] = base[p,:] + sum_t gates[t] * tables[t][ids[t][p],:] + pos_table[p % L,:]
 * in the reference's order of operations (product rounded, then added, left to
 * right, position row last) -> bit-exact in fp32.
 * Replaces tower_code/v1_refine_usertower.py:447-456 (6x aten::embedding + 7
 * elementwise kernels + the in-place adds).  `ids`, `tables`, `table_rows` are
 * HOST arrays of n_tables entries; gates is a device array [n_tables]
 * (sigmoid(seq_gate) * s_mask, :434-438).  base may be NULL, pos_table may be NULL.
 * The forward returns the STORED row for id 0 (padding rows are non-zero, :408-409). */
int rs_seq_front_fwd(const void* base, int base_dtype,
                     const int64_t* const* ids, const float* const* tables, const int64_t* table_rows,
                     int n_tables, const float* gates,
                     const float* pos_table, int64_t L, int64_t P, int64_t dim,
                     void* out, int out_dtype, int* oob_flag, void* stream);

/* U3: backward of the above (autograd of :447-456, i.e. 6x embedding_dense_backward
 * + the gate products, tower_code/v1_usertower_train.py:850).
 *   d_tables[t][id,:] += gates[t] * dx[p,:]   for id = ids[t][p] != padding_idx
 *   d_gates[t]         = sum_p < tables[t][ids[t][p],:], dx[p,:] >      (padding rows INCLUDED)
 *   d_pos[l,:]         = sum_b dx[b*L + l,:]
 * `big_mode[t]` (host): 0 = table is reduced by the caller with
 * rs_segment_reduce_rows (this call skips its rows and its gate dot), 1 = vector
 * atomics here, 2 = small table (rows <= 64): privatised per-warp copies, no
 * atomics, deterministic.  d_tables must be zero-filled by the caller for modes 1/2
 * (mode 2 overwrites).  d_gates / d_pos are overwritten. */
size_t rs_seq_front_bwd_workspace_bytes(int64_t P, int64_t L, int64_t dim, int n_tables,
                                        const int64_t* table_rows, const int* big_mode);
int rs_seq_front_bwd(const void* dx, int dx_dtype,
                     const int64_t* const* ids, const float* const* tables, const int64_t* table_rows,
                     const int* big_mode, int n_tables, const float* gates,
                     int64_t L, int64_t P, int64_t dim, int64_t padding_idx,
                     float* const* d_tables, float* d_gates, float* d_pos,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------- U2: SASRec static front */

/* out[b, :] = concat_i( gates[i] * tables[i][ids[i][b], :] , gates[9] * relu(cont[b,:] @ W^T + bias) )
 * 9 tiny tables (dims 16,16,16,16,4,4,4,4,4) + Linear(4->16): [B,100].
 * Replaces tower_code/v1_refine_usertower.py:472-491 (~25 launches). */
int rs_static_front_fwd(const int64_t* const* ids /*host[9]*/, const float* const* tables /*host[9]*/,
                        const int64_t* table_rows /*host[9]*/, const float* cont /*[B,4]*/,
                        const float* cont_w /*[16,4]*/, const float* cont_b /*[16]*/,
                        const float* gates /*dev[10]*/, int64_t B, float* out /*[B,100]*/,
                        int* oob_flag, void* stream);
/* backward: d_tables[i] (overwritten, padding row 0 excluded), d_gates[10], d_cont_w, d_cont_b. */
size_t rs_static_front_bwd_workspace_bytes(int64_t B);
int rs_static_front_bwd(const float* d_out /*[B,100]*/, const int64_t* const* ids, const float* const* tables,
                        const int64_t* table_rows, const float* cont, const float* cont_w, const float* cont_b,
                        const float* gates, int64_t B, int64_t padding_idx,
                        float* const* d_tables, float* d_gates, float* d_cont_w, float* d_cont_b,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------- U4: normalised item-matrix rows */

/* out[i,:] = table[ids[i],:] / max(||table[ids[i],:]||_2, eps);  inv_norm[i] = 1/max(||.||,eps)
 * Same values as F.normalize(table, p=2, dim=1)[ids] (tower_code/v1_usertower_train.py:810-811
 * + v1_refine_usertower.py:833) without normalising the other 100k rows. */
int rs_normalized_rows_fwd(const float* table, int64_t rows, int64_t dim, const int64_t* ids, int64_t n,
                           float eps, void* out, int out_dtype, float* inv_norm, int* oob_flag, void* stream);
/* d_table[id,:] += inv_norm * (g - v <v,g>)  with v the normalised row (rows at the eps
 * clamp: inv_norm * g).  The reference lets gradient flow into row 0 here (plain
 * indexing, not nn.Embedding) -- so does this. */
int rs_normalized_rows_bwd(const void* d_out, int d_out_dtype, const float* table, int64_t rows, int64_t dim,
                           const int64_t* ids, int64_t n, float eps, float* d_table, void* stream);

/* ------------------------------------------------------ I1/I2: item fronts */

/* I1: out[r,:] = LayerNorm(table[ids[r],:] + field_emb[r % n_fields,:]) -- item_tower.py:239-241.
 * mean/rstd [n] are saved for the backward. */
int rs_std_front_fwd(const float* table, int64_t rows, int64_t dim, const int64_t* ids, int64_t n,
                     const float* field_emb, int64_t n_fields, const float* ln_w, const float* ln_b, float eps,
                     void* out, int out_dtype, float* mean, float* rstd, int* oob_flag, void* stream);

/* I2: HF BertEmbeddings under no_grad: out[r,t,:] = LN(word[ids[r,t]] + type[0] + pos[t]) (eps 1e-12),
 * then dropout(p) with a counter-based RNG when p > 0 (item_tower.py:248-249; the module is in
 * train mode there, SURVEY.md 8c invariant 6). */
int rs_bert_embed_fwd(const float* word, int64_t vocab, const float* pos, const float* type0,
                      const float* ln_w, const float* ln_b, float eps, const int64_t* ids, int64_t n_seq,
                      int64_t T, int64_t dim, float dropout_p, uint64_t seed,
                      void* out, int out_dtype, int* oob_flag, void* stream);

/* I2 tail: out[r,:] = sum_t feats[r,t,:]*mask[r,t] / max(sum_t mask[r,t], 1e-9) -- item_tower.py:254-257 */
int rs_masked_mean_fwd(const void* feats, int feats_dtype, const int64_t* mask, int64_t n_seq, int64_t T,
                       int64_t dim, float* out, void* stream);
int rs_masked_mean_bwd(const float* d_out, const int64_t* mask, int64_t n_seq, int64_t T, int64_t dim,
                       void* d_feats, int d_feats_dtype, void* stream);

/* ----------------------------------- sequence encoder on packed valid tokens */

/* The reference runs nn.TransformerEncoder over the padded [B, L] grid with a causal + key-padding mask
 * (tower_code/v1_refine_usertower.py:343-352, :458-466).  Valid positions only ever attend to valid positions of
 * their own sequence and the rest of the layer is position-wise, so the encoder runs on the PACKED valid tokens
 * [T, 128]; cu_seqlens[n_seq+1] (int32) delimits the sequences.  The matmuls stay library GEMMs (nn.Linear in the
 * reference as well); these entry points are everything between them.  Dropout masks come from a counter-based hash
 * of (seed, element): nothing is stored, the backward re-derives them from the same seed.
 *
 * Causal attention for sequences of <= 64 tokens, head_dim == 32, straight from the packed in_proj output
 * qkv[T, 3, H, 32] (q | k | v, heads contiguous -- nn.MultiheadAttention's layout): out[T, H*32], lse[T, H].
 * Replaces the head transposes, the [B,H,L,L] merged-mask tensor, F.scaled_dot_product_attention and its dropout.
 * zero_tail: the last `zero_tail` sequences are queries at PADDED positions -- in the reference's padded grid every
 * key of such a query is masked and the attention output is 0 (fully-masked softmax row); they get out = 0 and no
 * gradient.  The reference reads such rows: `last_indices = valid_mask.sum(1) - 1` (v1_usertower_train.py:830)
 * addresses its LEFT-padded grid from the left, i.e. a padded position whenever a sequence fills less than half
 * of the window; the packed encoder carries those positions as extra one-token sequences to reproduce it.
 * one_row_from / one_rows: the sequences [one_row_from, n_seq - zero_tail) are only read at ONE token each (the second
 * dropout view in the last encoder layer feeds nothing but its DuoRec row, v1_usertower_train.py:789,830-842):
 * one_rows[k] (int64) is the packed row of that token for sequence one_row_from + k (NULL: the last token; a row
 * outside the sequence: none).  That row is computed by a matrix-vector kernel, the sequence's other rows of `out` are
 * zeros, and the backward builds d_qkv of all its tokens from that row's gradient alone.  one_row_from = -1: none.
 * Backward kernels (16-bit operands): with bias == NULL (the in_proj bias already inside qkv -- GEMM epilogue) every
 * (query tile, key tile) pair of a sequence is visited once (one-tile sequences in registers only, longer ones with dQ
 * summed in shared memory); with a bias pointer the two-phase kernel (dQ pass, then dK/dV pass) runs.  Same results. */
int rs_attn_varlen_fwd(const void* qkv, int dtype, const float* bias /*[3*H*32] in_proj bias added on load, or NULL*/,
                       const int32_t* cu_seqlens, int64_t n_seq, int64_t total_tokens,
                       int n_heads, int head_dim, int max_len, int64_t zero_tail, int64_t one_row_from,
                       const int64_t* one_rows, float scale, float dropout_p, uint64_t seed, void* out, float* lse,
                       void* stream);
int rs_attn_varlen_bwd(const void* qkv, const void* d_out, const void* out, int dtype, const float* bias,
                       const float* lse,
                       const int32_t* cu_seqlens, int64_t n_seq, int64_t total_tokens, int n_heads, int head_dim,
                       int max_len, int64_t zero_tail, int64_t one_row_from, const int64_t* one_rows, float scale,
                       float dropout_p, uint64_t seed, void* d_qkv, void* stream);
/* y[r,:] = dropout(LayerNorm(x[index ? index[r] : r, :])), dim == 128, fp32 statistics (mean/rstd[n_rows] saved).
 * `index` packs the valid rows of the padded grid on the way in (emb_ln, :458-459) and may repeat a row (one copy per
 * dropout view).  Backward: dx[r] = gradient w.r.t. the row that was normalised for output r ([n_rows,128], packed
 * like dy: the caller scatter-adds it by `index`), dw/db[128] reduced in a fixed order. */
int rs_ln_fwd(const void* x, int x_dtype, const int64_t* index, int64_t n_rows, int64_t dim, const float* w,
              const float* b, float eps, float dropout_p, uint64_t seed, void* y, int y_dtype, float* mean,
              float* rstd, void* stream);
/* y[r,:] = dropout(act(LayerNorm(x[r,:] + add[add_rows ? add_rows[r] : r, :] + lin_bias))), dim == 128; act: 0 identity,
 * 1 exact (erf) GELU.  The late-fusion head output_proj = Linear(256,128) -> LayerNorm -> GELU -> Linear
 * (tower_code/v1_refine_usertower.py:394-399, applied to cat([sequence output, user profile]) at :499-510) without the
 * concatenation: x = the sequence half of the first Linear (one GEMM on the rows), add = the profile half, fp32, one row
 * per user (`add`, `add_rows`, `lin_bias` nullable).  Also serves static_mlp's LayerNorm -> GELU -> Dropout (:384-389).
 * Backward: dx[n_rows,128] (x's dtype) = gradient of the LN input of every row; the caller sums it per profile row for
 * d add and over all rows for d lin_bias.  Workspace: rs_ln_bwd_workspace_bytes(n_rows). */
int rs_ln_act_fwd(const void* x, int x_dtype, const float* add, const int64_t* add_rows, const float* lin_bias,
                  int64_t n_rows, int64_t dim, const float* w, const float* b, float eps, int act, float dropout_p,
                  uint64_t seed, void* y, int y_dtype, float* mean, float* rstd, void* stream);
int rs_ln_act_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* add, const int64_t* add_rows,
                  const float* lin_bias, int64_t n_rows, int64_t dim, const float* w, const float* b, const float* mean,
                  const float* rstd, int act, float dropout_p, uint64_t seed, void* dx, float* dw, float* db,
                  void* workspace, size_t workspace_bytes, void* stream);
size_t rs_ln_bwd_workspace_bytes(int64_t n_rows);
int rs_ln_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const int64_t* index, int64_t n_rows,
              int64_t dim, const float* w, const float* mean, const float* rstd, float dropout_p, uint64_t seed,
              void* dx, const float* residual_grad /* nullable [n_rows,128] fp32, added to dx: the gradient that
              reaches the LayerNorm's input through the residual connection around it */,
              float* dw, float* db, void* workspace, size_t workspace_bytes, void* stream);
/* The biases of the four Linear layers of an encoder layer are folded into the kernel that consumes the GEMM output
 * (`bias` fp32 [n_cols] or NULL), so the GEMMs are plain matmuls and the bias gradients are column sums of tensors
 * these kernels produce (rs_colsum) instead of separate reductions behind each GEMM.
 * out = x + dropout(y + bias) (n elements = rows * n_cols); backward of the y branch: dy = mask(g) / keep */
int rs_dropout_add_fwd(const void* x, int x_dtype, const void* y, int y_dtype, const float* bias, int64_t n_cols,
                       int64_t n, float dropout_p, uint64_t seed, void* out, void* stream);
int rs_dropout_bwd(const void* g, int g_dtype, int64_t n, float dropout_p, uint64_t seed, void* dy, int dy_dtype,
                   void* stream);
/* out = dropout(gelu(z + bias)), exact erf GELU (activation="gelu"); backward dz = mask(g)/keep * gelu'(z + bias) */
int rs_gelu_dropout_fwd(const void* z, int dtype, const float* bias, int64_t n_cols, int64_t n, float dropout_p,
                        uint64_t seed, void* out, void* stream);
int rs_gelu_dropout_bwd(const void* z, const void* g, int dtype, const float* bias, int64_t n_cols, int64_t n,
                        float dropout_p, uint64_t seed, void* dz, void* stream);
/* The encoder's input x0 = dropout(LayerNorm_emb(e[index])) (tower_code/v1_refine_usertower.py:458-459) and the first
 * layer's h = LayerNorm_1(x0) in one pass over the packed rows (x0 fp32, h in the GEMM operand dtype; both LayerNorms'
 * statistics saved).  Backward, one warp per SOURCE row u of e (every u is read by exactly the two packed rows inv1[u],
 * inv2[u] -- the two dropout views, rs_batch_index_build): d e[u] = LN_emb'( mask(residual_grad + LN_1'(dh))[inv1[u]] +
 * the same at inv2[u] ), the four parameter gradients reduced in a fixed order.  dim == 128. */
int rs_emb_ln2_fwd(const void* x, int x_dtype, const int64_t* index, int64_t n_rows, int64_t dim, const float* w0,
                   const float* b0, float eps0, float dropout_p, uint64_t seed, const float* w1, const float* b1,
                   float eps1, float* x0, void* h, int h_dtype, float* mean0, float* rstd0, float* mean1, float* rstd1,
                   void* stream);
size_t rs_emb_ln2_bwd_workspace_bytes(int64_t n_src);
int rs_emb_ln2_bwd(const void* dh, int dh_dtype, const void* x, int x_dtype, const float* x0, const float* residual_grad,
                   const int64_t* inv1, const int64_t* inv2, int64_t n_src, int64_t dim, const float* w0, const float* w1,
                   const float* mean0, const float* rstd0, const float* mean1, const float* rstd1, float dropout_p,
                   uint64_t seed, void* dx, float* dw0, float* db0, float* dw1, float* db1, void* workspace,
                   size_t workspace_bytes, void* stream);
/* A pre-norm block ends with x1 = x + dropout(y + lin_bias) and the next one starts with h = LayerNorm(x1)
 * (nn.TransformerEncoderLayer, norm_first=True; tower_code/v1_refine_usertower.py:343-352): both in one pass (x, x1 fp32
 * residual stream; y the GEMM output; h in the next GEMM's operand dtype; mean/rstd saved), and one backward pass:
 * dx = residual_grad + LN'(dh) (fp32), dy = mask(dx)/keep (y's dtype), d gamma / d beta / d lin_bias reduced in a fixed
 * order.  dim == 128.  Workspace: rs_ln_bwd_dropout_workspace_bytes(n_rows). */
int rs_dropout_add_ln_fwd(const float* x, const void* y, int y_dtype, const float* lin_bias, int64_t n_rows, int64_t dim,
                          float dropout_p, uint64_t seed, const float* w, const float* b, float eps, float* x1, void* h,
                          int h_dtype, float* mean, float* rstd, void* stream);
size_t rs_ln_bwd_dropout_workspace_bytes(int64_t n_rows);
int rs_ln_bwd_dropout(const void* dh, int dh_dtype, const float* x1, const float* residual_grad /*nullable*/,
                      int64_t n_rows, int64_t dim, const float* w, const float* mean, const float* rstd, float dropout_p,
                      uint64_t seed, float* dx, void* dy, int dy_dtype, float* dw, float* db, float* d_lin_bias,
                      void* workspace, size_t workspace_bytes, void* stream);
/* The two backward kernels with the bias gradient folded in: d_bias[n_cols] = column sums of their output (fp32, fixed
 * order), no second pass over dy / dz.  n_cols / 4 must divide 256.  Workspace: rs_ew_colsum_workspace_bytes(n, n_cols).
 * For 16-bit activations the GELU uses erfc by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 << the 2^-9 rounding of the
 * result); fp32 activations use erff. */
size_t rs_ew_colsum_workspace_bytes(int64_t n, int64_t n_cols);
int rs_dropout_bwd_bias(const void* g, int g_dtype, int64_t n, int64_t n_cols, float dropout_p, uint64_t seed, void* dy,
                        int dy_dtype, float* d_bias, void* workspace, size_t workspace_bytes, void* stream);
int rs_gelu_dropout_bwd_bias(const void* z, const void* g, int dtype, const float* bias, int64_t n_cols, int64_t n,
                             float dropout_p, uint64_t seed, void* dz, float* d_bias, void* workspace,
                             size_t workspace_bytes, void* stream);
/* Advance the device-side dropout epoch that every kernel above mixes into its (by-value) seed.  Stream-ordered: call
 * it once at the start of a train step.  Inside a captured CUDA graph it is what makes every REPLAY draw new masks. */
int rs_rng_advance(void* stream);
/* y = x / max(||x||_2, eps) per 128-wide row == F.normalize(x, p=2, dim=-1) (v1_refine_usertower.py:510, the loop's
 * v1_usertower_train.py:807); inv_norm[n_rows] is saved for the backward (negative where the clamp was active).
 * bwd: dx = (g - y <g, y>) * inv_norm; y must be the fp32 output of the forward. */
int rs_l2_normalize_fwd(const void* x, int x_dtype, int64_t n_rows, int64_t dim, float eps, void* y, int y_dtype,
                        float* inv_norm, void* stream);
int rs_l2_normalize_bwd(const void* g, int g_dtype, const void* y, int y_dtype, const float* inv_norm, int64_t n_rows,
                        int64_t dim, void* dx, int dx_dtype, void* stream);
/* out[c] = sum_r x[r, c], fp32, fixed summation order (n_cols % 4 == 0, <= 1024) */
size_t rs_colsum_workspace_bytes(int64_t n_rows, int64_t n_cols);
int rs_colsum(const void* x, int dtype, int64_t n_rows, int64_t n_cols, float* out, void* workspace,
              size_t workspace_bytes, void* stream);

/* -------------------------------------------- C1..C5: fused in-batch softmax */

/* One pass over S = scale * (A @ B^T) - col_bias, [M,N], never written to memory:
 *   masked (i,j): j != i+diag_offset and (key_a_row[i]==key_a_col[j] or key_b_row[i]==key_b_col[j])
 *                 -> S = mask_value  (-inf in C2/C3/C4; -1e4 / -1e9 in C5)
 *   RS_CE_DIAG_MASK: the diagonal itself is masked (SupCon, v1_refine_usertower.py:613)
 *   RS_CE_DIAG_RAW:  the diagonal logit skips col_bias (positive recovery, mined_inference.py:774-775)
 *   RS_CE_SUPCON:    additionally pos_sum[i] = sum_j [key_a equal, key != 0, j != diag] S_ij and pos_cnt[i]
 *   RS_CE_NO_DIAG:   no column is a label (diag_offset ignored, diag[] not written): the columns are a set of
 *                    DISTINCT items with multiplicities folded into col_bias (bias_c - log m_c) and the row's own
 *                    item masked through key_a; the caller adds the positive back (see losses.logq_infonce_columns)
 * Keys are compared on their low 32 bits: ids must lie in [0, 2^32) (item ids and batch-row user ids do).
 * Outputs per row: lse[i] = logsumexp_j S_ij, diag[i] = S_{i,i+diag_offset}.
 * A,B are bf16 or fp16 [M,K] / [N,K] row-major, K == 128 (tcgen05 kind::f16, fp32 accumulate in TMEM).
 * Replaces mm + div + masks + masked_fill + cross_entropy at item_tower.py:1076-1082,
 * tower_code/v1_refine_usertower.py:836-861, :588-625, :794-815, mined_inference.py:738-789. */
#define RS_CE_DIAG_MASK 1
#define RS_CE_DIAG_RAW 2
#define RS_CE_SUPCON 4
#define RS_CE_NO_DIAG 8
typedef struct {
  const void* a;            /* [M,K] */
  const void* b;            /* [N,K] */
  int ab_dtype;             /* RS_BF16 or RS_F16 */
  int64_t M, N, K;
  float scale;              /* 1/temperature */
  const float* col_bias;    /* [N] or NULL: lambda*logq[tgt_j] */
  const int64_t* key_a_row; /* [M] or NULL */
  const int64_t* key_a_col; /* [N] */
  const int64_t* key_b_row; /* [M] or NULL */
  const int64_t* key_b_col; /* [N] */
  int64_t diag_offset;      /* label of row i is column i + diag_offset */
  float mask_value;
  int flags;
  float logit_bound;        /* > 0: the caller vouches that |<a_i, b_j>| <= logit_bound for all i, j (e.g. 1.0 for
                               unit-norm rows).  With a bounded exponent range the forward then sums 2^(s - C)
                               against a fixed offset C instead of tracking a running maximum, and the backward
                               factors per-row / per-column constants out of the exponent.  0: unknown (general path) */
} rs_ce_problem;
size_t rs_ce_workspace_bytes(const rs_ce_problem* p);
int rs_ce_fwd(const rs_ce_problem* p /*host*/, float* lse /*[M]*/, float* diag /*[M]*/,
              float* pos_sum /*[M] or NULL*/, float* pos_cnt /*[M] or NULL*/,
              void* workspace, size_t workspace_bytes, void* stream);
/* Backward.  With upstream per-row weights (the gradients of the scalar loss w.r.t. lse, diag, pos_sum):
 *   dS_ij = w_lse[i] * softmax_ij + w_diag[i] * [j == i+diag_offset] + w_pos[i] * [positive ij]
 *   dA = scale * dS @ B   [M,K],   dB = scale * dS^T @ A   [N,K]     (fp32, overwritten)
 * S is recomputed tile by tile on the tensor cores, dS goes TMEM -> registers -> bf16/fp16 tile in
 * shared memory -> second tcgen05.mma; neither S nor dS ever reaches HBM.  Masked entries get no
 * gradient (masked_fill semantics); with RS_CE_DIAG_RAW the diagonal's bias is not differentiated.
 * w_diag / w_pos may be NULL (= 0). */
int rs_ce_bwd(const rs_ce_problem* p /*host*/, const float* lse, const float* w_lse /*[M]*/,
              const float* w_diag /*[M] or NULL*/, const float* w_pos /*[M] or NULL*/, float* dA, float* dB,
              void* workspace, size_t workspace_bytes, void* stream);

/* Forward that also accumulates the row side of the backward ("flash" form of the same loss; same reference lines as
 * rs_ce_fwd / rs_ce_bwd).  With a bounded logit range (logit_bound > 0 and a moderate bias range, decided on the
 * device) every exponential e_ij = 2^(S_ij - C) is final when it is formed, and dS_ij = (w_i / l_i) e_ij is a per-row
 * scalar times e_ij: the pass that sums l_i also accumulates G_i = sum_j bf16(e_ij) B_j on the tensor cores (second
 * tcgen05.mma, P from TMEM), and the backward's dA is a row scaling of G.  Per loss the tensor cores then execute
 * S, P@B (here) and S, dS^T@A (rs_ce_bwd_from_grad, the dB side): 4 contractions instead of 5.
 *   g_parts : caller-owned [rs_ce_fwd_grad_bytes(p)] bytes, kept until the backward (split partials of G, fp32)
 *   g_info  : caller-owned float[8], kept until the backward: [0] = C (log2 units), [1] != 0 iff G is valid (else the
 *             backward falls back to its own row-side pass, stream-ordered, no host decision), [3] = scale2*bound
 * bf16 operands only; no RS_CE_SUPCON; masks must be -inf (a masked entry must carry no softmax mass).
 * lse / diag exactly as rs_ce_fwd. */
size_t rs_ce_fwd_grad_bytes(const rs_ce_problem* p);
/* The per-row tail of the in-batch loss over DISTINCT item columns (tower_code/v1_refine_usertower.py:826-861 regrouped by
 * item, see losses.logq_infonce_columns):  Z_i = e^{lse0_i} + e^{pos_i} - e^{own_i},  loss = sum_i w_i (log max(Z_i, 1e-30)
 * - pos_i), w_i = row_weight[i] (NULL: 1/n, the mean).  lse0 = log-sum-exp over the columns without the row's own item
 * (rs_ce_fwd), pos = the label's logit, own = log-sum-exp of the row's user's OTHER targets (NULL: none).  Also returns
 * the gradient coefficients c_lse0 / c_pos / c_own [n] (d loss / d input).  Fixed-order two-stage sum.  Rows with
 * w_i == 0 (the padding rows of a bucketed batch) contribute exactly 0 whatever their inputs hold. */
size_t rs_ce_row_combine_workspace_bytes(int64_t n);
int rs_ce_row_combine(const float* lse0, const float* pos, const float* own, const float* row_weight, int64_t n,
                      float* loss, float* c_lse0, float* c_pos, float* c_own, void* workspace, size_t workspace_bytes,
                      void* stream);

int rs_ce_fwd_grad(const rs_ce_problem* p /*host*/, float* lse /*[M]*/, float* diag /*[M]*/, float* g_parts,
                   float* g_info, void* workspace, size_t workspace_bytes, void* stream);
/* rs_ce_bwd with the row side taken from rs_ce_fwd_grad's outputs: dA = scale * (w_lse/l) * G + scale * w_diag * B[label]
 * (one pass over G), dB by the transposed tensor-core pass as in rs_ce_bwd.  dA or dB may be NULL. */
int rs_ce_bwd_from_grad(const rs_ce_problem* p /*host*/, const float* lse, const float* w_lse /*[M]*/,
                        const float* w_diag /*[M] or NULL*/, const float* g_parts, const float* g_info,
                        float* dA, float* dB, void* workspace, size_t workspace_bytes, void* stream);

/* Same-user block of the distinct-item softmax (losses.logq_infonce_columns; the same-user mask of
 * tower_code/v1_refine_usertower.py:848).  Rows are grouped by user (row_cu[n_users+1], <= 64 rows per user);
 * pos_col[i] is the column (distinct item) of row i's target.  Per user, over its own rows i, j:
 *   s_ij = scale*<u_i, cols[pos_col_j]> - col_bias[pos_col_j];  s_pos[i] = s_ii;
 *   own_lse[i] = log sum_{j: pos_col_j != pos_col_i} exp(s_ij)   (-inf if none)
 * Backward: d_u[n_rows,128] (overwritten), d_cols[n_cols,128] (ACCUMULATED: pass zeros), fp32, one row added per
 * (user, item) instead of one per (row, item) pair.  u / cols: fp32, fp16 or bf16 [., 128]. */
int rs_user_block_logits_fwd(const void* u, const void* cols, int dtype, const int64_t* pos_col, const int32_t* row_cu,
                             int64_t n_users, int64_t n_cols, int64_t dim, int max_len, float scale,
                             const float* col_bias, float* s_pos, float* own_lse, void* stream);
int rs_user_block_logits_bwd(const void* u, const void* cols, int dtype, const int64_t* pos_col, const int32_t* row_cu,
                             int64_t n_users, int64_t n_cols, int64_t dim, int max_len, float scale,
                             const float* col_bias, const float* own_lse, const float* g_pos, const float* g_own,
                             float* d_u, float* d_cols, void* stream);

/* ------------------------------------------- N1: on-device batch assembly for the train step */

/* The reference derives the step's index sets from the collated [B, L] batch inside the step with boolean indexing
 * (`output_1[valid_mask]`, `target_ids[valid_mask]`, tower_code/v1_usertower_train.py:794-804; the dataset side is
 * tower_code/v1_refine_usertower.py:204-306): nonzero + host synchronisation, data-dependent shapes.  This call builds
 * every index the packed step consumes on the device, stream-ordered, into arrays of STATIC capacity (shape buckets):
 * true counts in meta[], rows / columns beyond them are inert padding (row_weight 0, col_counts 0).
 *   packed order     t = (valid time steps, batch-major);  T = meta[0]
 *   U1 grid          [T valid | E positions `len-1` that are padding (v1_usertower_train.py:830 on a left-padded grid) | pad]
 *   encoder tokens   [view-1 valid | view-2 valid | view-1 extras | view-2 extras | pad]  (2 * grid_cap slots; everything
 *                    behind 2T is ONE zero-tail pseudo-sequence: cu_seqlens_2v has 2B + 2 entries, n_seq = 2B + 1)
 *   columns          distinct valid targets in ascending id order, col_counts = multiplicity;  U = meta[2]
 * meta[8] (int32): [0] T, [1] E, [2] U, [3] flags: 1 = a capacity was too small (outputs truncated, do not use),
 *                  2 = an empty sequence, 4 = a target id outside [0, n_item_rows).
 * padding_mask: [B, L] bytes, 1 = padding (torch.bool).  L <= 64.  grid_cap % 64 == 0, grid_cap >= tok_cap. */
typedef struct {
  const uint8_t* padding_mask;
  const int64_t* item_ids;       /* [B, L] */
  const int64_t* time_ids;       /* [B, L] */
  const int64_t* target_ids;     /* [B, L] */
  int64_t B, L, n_item_rows;
  int64_t tok_cap;               /* capacity of the main-loss row arrays   (>= T) */
  int64_t col_cap;               /* capacity of the column arrays          (>= U) */
  int64_t grid_cap;              /* capacity of the U1 token grid          (>= T + E) */
  int64_t* pk_item_ids;          /* [grid_cap] */
  int64_t* pk_time_ids;          /* [grid_cap] */
  int64_t* pk_pos_ids;           /* [grid_cap]  l + 1 (0 = padding slot) */
  int64_t* pk_index_2v;          /* [2*grid_cap] U1 row of every encoder token */
  int64_t* fold_inv1;            /* [grid_cap]  encoder slot of the view-1 copy of every U1 row */
  int64_t* fold_inv2;            /* [grid_cap]  ... of the view-2 copy */
  int32_t* cu_seqlens_2v;        /* [2B + 2] */
  int32_t* row_cu;               /* [B + 1]  rows of every user among the main-loss rows */
  int64_t* select_2v;            /* [tok_cap + 2B] encoder slots of: the main rows (slot t for row t, padding rows too) |
                                    DuoRec rows view 1 | view 2 */
  int64_t* users_2v;             /* [tok_cap + 2B] row of the (2B-row, two-view) profile matrix for each of those; ascending
                                    over the main rows (padding rows: B), then 0..B-1, B..2B-1 */
  int64_t* main_tgt;             /* [tok_cap] target item of every main row */
  int64_t* last_tgt;             /* [B] target at the DuoRec position */
  float* row_weight;             /* [tok_cap] 1/T for real rows, 0 for padding */
  int64_t* col_item_ids;         /* [col_cap] */
  float* col_counts;             /* [col_cap] */
  int64_t* pos_col;              /* [tok_cap] column of every main row's target */
  int32_t* meta;                 /* [8] */
} rs_batch_index;
size_t rs_batch_index_workspace_bytes(int64_t B, int64_t L, int64_t n_item_rows);
int rs_batch_index_build(const rs_batch_index* d /*host*/, void* workspace, size_t workspace_bytes, void* stream);
/* only the counts (meta[0..2] = T, E, U): what a loader needs to pick the shape bucket of a batch */
int rs_batch_index_counts(const uint8_t* padding_mask, const int64_t* target_ids, int64_t B, int64_t L,
                          int64_t n_item_rows, int32_t* meta, void* workspace, size_t workspace_bytes, void* stream);
/* out[u,:] = x[i1[u],:] + x[i2[u],:]  (x, out in `dtype`; an index outside [0, n_src) contributes 0): folds the
 * gradients of the two dropout views of a U1 row (fold_inv1 / fold_inv2) without a scatter. */
int rs_gather_add2(const void* x, int dtype, const int64_t* i1, const int64_t* i2, int64_t n, int64_t n_src, int64_t dim,
                   void* out, void* stream);

/* ------------------------------------------------------- R1: top-k retrieval */

/* ids/scores[b, 0..k) = top-k over items of <users[b,:], items[j,:]>, fp32, sorted by
 * (score desc, id asc); item 0 excluded when mask_index0.  The [b, n_items] score matrix
 * is never written.  Replaces matmul + topk at tower_code/v1_usertower_train.py:672-675,
 * mined_inference.py:901-909,1103-1108,1536-1543, temp_model/ranker_skelet.py:193-196.
 * k <= 1024. */
size_t rs_topk_workspace_bytes(int64_t n_users, int64_t n_items, int64_t dim, int64_t k);
int rs_retrieve_topk(const float* users, int64_t n_users, const float* items, int64_t n_items, int64_t dim,
                     int64_t k, int mask_index0, int64_t* out_ids, float* out_scores,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------- row-sharded tables: device-side routing */

/* The reference is single-process; these serve the N > 1 step's `owner = id % world` row sharding.
 * Conventions: an id of -1 is the NULL id everywhere in this library (rs_gather_rows / rs_normalized_rows_fwd read
 * zeros, rs_sort_ids / the scatters skip it, no oob flag).
 *
 * cnt[id] = number of i < min(n, *n_valid_dev) with ids[i] == id (cnt is cleared first; n_valid_dev may be NULL);
 * force_bin0: count id 0 once more (the padding id must always own slot 0 of rank 0's list). */
int rs_id_histogram(const int64_t* ids, int64_t n, const int32_t* n_valid_dev, int64_t n_bins, int force_bin0,
                    int32_t* cnt, int* oob_flag, void* stream);
/* Owner-major compaction of the present ids (cnt > 0) of a catalogue of n_ids ids sharded as owner = id % world,
 * local row = id / world (rows_per_owner = ceil(n_ids / world)): owner by owner, ascending local row, at most `cap`
 * per owner.  Slot o = owner * cap + s holds
 *   out_rows[o] = local row (-1: empty)    -- the request list of an equal-split all-to-all
 *   out_ids[o]  = global id (0: empty), out_cnt[o] = cnt as fp32 (0: empty)          (both optional)
 * and slot_of[id] = o for present ids, -1 otherwise ([n_ids] int32).
 * meta (int32[4]): [0] largest per-owner count, [1] 1 iff it exceeds cap (outputs truncated), [2] present ids. */
size_t rs_owner_compact_workspace_bytes(int world, int64_t rows_per_owner);
int rs_owner_compact(const int32_t* cnt, int world, int64_t rows_per_owner, int64_t n_ids, int64_t cap,
                     int64_t* out_rows /*nullable*/, int64_t* out_ids, float* out_cnt, int32_t* slot_of, int32_t* meta,
                     void* workspace, size_t workspace_bytes, void* stream);
/* out[i] = table[ids[i]] widened to int64; `fill` where the id is outside [0, n_table) or the entry is negative */
int rs_lookup_i32(const int32_t* table, int64_t n_table, const int64_t* ids, int64_t n, int64_t fill, int64_t* out,
                  void* stream);

/* ------------------------------------------------------- N4: ensemble merge of two retrieval lists */

/* tower_code/mined_inference.py:1110-1189 (min-max weighted sum) and :1337-1411 (weighted reciprocal-rank fusion).
 * cand_ids [n_users, P]: the two models' top-M lists side by side (P = 2M <= 2048; an item both models rank appears
 * twice); s1, s2 [n_users, P]: both models' re-scored candidates (fp32).  Per user and per blend weight alpha[a]:
 *   mode 0: n = (s - min) / (max - min + 1e-9) per model;   mode 1: n = 1 / (k_rrf + rank + 1), rank by descending s
 *   final = alpha * n1 + (1 - alpha) * n2;  the k_sel (= max_k + 20 in the reference) best in (final desc, id asc)
 *   order;  duplicates dropped, order kept (the reference does this per user with np.unique on the host, :1182-1183).
 * out_ids [n_alpha, n_users, k_sel] int64, -1 padded behind out_cnt [n_alpha, n_users] distinct ids.
 * out_n1 / out_n2 (optional, [n_users, P]): the normalised scores / reciprocal ranks.  alphas: HOST doubles.
 * ids are compared and returned on their low 32 bits. */
int rs_ensemble_merge(const int64_t* cand_ids, const float* s1, const float* s2, int64_t n_users, int64_t P,
                      int mode, float k_rrf, const double* alphas /*host*/, int n_alpha, int64_t k_sel,
                      int64_t* out_ids, int32_t* out_cnt, float* out_n1, float* out_n2, void* stream);

/* Same contract as rs_retrieve_topk (same reference lines), with the contraction on the tensor cores as a FILTER:
 * bf16 candidate pass on tcgen05 (every item whose approximate score is within twice the rounding bound
 * 2^-8 |u| |i|max of the row's k-th best survives) + exact fp32 re-scoring of the survivors + top-k by (score desc,
 * id asc): the ids are the fp32 ranking's.  If a candidate list overflows, a device flag routes the call through the
 * exact fp32 kernel (launched behind, returning at once otherwise).  dim == 128, k <= 32. */
size_t rs_retrieve_topk_tc_workspace_bytes(int64_t n_users, int64_t n_items, int64_t dim, int64_t k);
int rs_retrieve_topk_tc(const float* users, int64_t n_users, const float* items, int64_t n_items, int64_t dim,
                        int64_t k, int mask_index0, int64_t* out_ids, float* out_scores,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------- C4 / C5: hard-negative mining + sparse logits */

/* Mining (no gradient) of tower_code/v1_refine_usertower.py:775-791 (:643-668, :707-719): for each row i the
 * top-k columns j of cos_ij = <u_i, v_j> over the columns that are NOT ignored, where
 *   ignored(i,j) = key[i] == key[j]  ||  ( <v_i, v_j> > hnm_threshold && i != j ).
 * u, v: [n, dim] fp32, L2-normalised by the caller.  Fused: both Gram products share the item tile, the
 * [n,n] similarity / mask matrices are never written.  out_* [n,k] sorted (score desc, id asc), padded with
 * (-inf, -1) when a row has fewer than k candidates; avail[i] = number of non-ignored columns. */
size_t rs_mine_workspace_bytes(int64_t n, int64_t dim, int64_t k);
int rs_mine_hard_negatives(const float* u, const float* v, const int64_t* key, int64_t n, int64_t dim, int64_t k,
                           float hnm_threshold, int64_t* out_ids, float* out_scores, int32_t* avail,
                           void* workspace, size_t workspace_bytes, void* stream);

/* out[i,c] = scale * <a_i, b_idx[i,c]> - bias[idx[i,c]]   (-inf where idx < 0 or key_row[i] == key_col[idx]).
 * The gathered logits of torch.gather(logits, 1, top_k_indices) (:677-678, :732-734) without the [n,n] logits.
 * a [n,dim], b [m,dim] in ab_dtype (f32/f16/bf16); fp32 accumulation. */
int rs_sparse_logits_fwd(const void* a, const void* b, int ab_dtype, const int64_t* idx, int64_t n, int64_t m,
                         int64_t k, int64_t dim, float scale, const float* bias, const int64_t* key_row,
                         const int64_t* key_col, float* out, void* stream);
/* d_a [n,dim] fp32 overwritten; d_b [m,dim] fp32 accumulated (caller zero-fills). */
int rs_sparse_logits_bwd(const void* a, const void* b, int ab_dtype, const int64_t* idx, int64_t n, int64_t m,
                         int64_t k, int64_t dim, float scale, const int64_t* key_row, const int64_t* key_col,
                         const float* g, float* d_a, float* d_b, void* stream);

/* ------------------------------------------------------------- F1: FM / DeepFM */

/* No reference implementation exists: follows deepctr-torch 0.2.9 FM.
 * ids [B,F] local ids, offsets [F] (device) first row of each field in the concatenated
 * tables; emb [sum_vocab,k] fp32 (k in {4,8,16,32,64,128}), lin [sum_vocab] or NULL.
 *   fm[b]  = 0.5 * sum_d( (sum_f v)^2 - sum_f v^2 )  (+ sum_f lin[id] when lin != NULL)
 *   concat [B,F*k] (optional, may be NULL): the gathered rows, input of the deep MLP. */
int rs_fm_fwd(const int64_t* ids, const int64_t* offsets, int64_t B, int64_t F, const float* emb, int64_t k,
              const float* lin, int64_t total_rows, float* fm, void* concat, int concat_dtype,
              int* oob_flag, void* stream);
/* d_emb[row,:] += d_fm[b]*(S_b - v) + d_concat[b,f,:];  d_lin[row] += d_fm[b]  (vector atomics). */
int rs_fm_bwd(const int64_t* ids, const int64_t* offsets, int64_t B, int64_t F, const float* emb, int64_t k,
              const float* d_fm, const void* d_concat, int d_concat_dtype, int64_t total_rows,
              float* d_emb, float* d_lin, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RS_TWOTOWER_H_ */
