/* legommenders_b200 — C ABI of the B200-native Legommenders hot path.
 *
 * The reference (Jyonn/Legommenders) is pure Python over PyTorch and has no FFI; every entry point below
 * replaces the aten/library call(s) the reference makes at the cited file:line (paths relative to the
 * reference root).  The Python host in legommenders_b200/ binds these with ctypes (see INTEGRATION.md for
 * the stub a reference maintainer would add).
 *
 * Conventions: all pointers are DEVICE pointers owned by the caller; no allocation, no synchronisation;
 * work is enqueued on `stream`; ids/masks are int64 (the reference's wire format, loader/resampler.py:134);
 * floating point is fp32; row widths must be multiples of 4 floats (16-byte vector loads).  Every function
 * returns 0 on success or a negative LK_ERR_* code; lk_last_error() returns the message for the calling thread.
 */
#ifndef LEGOMMENDERS_B200_H_
#define LEGOMMENDERS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

const char* lk_version(void);
/* Id bounds: every gather / index / scoring kernel treats an id outside [0, table_rows) as an invalid position (zero row, zero score;
 * never dereferenced) and adds 1 per offending id to this DEVICE counter (null = do not count).  The reference raises IndexError
 * (aten::embedding / tensor indexing, model/legommender.py:153-157); the host mirror raises it when it next reads the counter. */
void lk_set_id_violation_counter(int32_t* device_counter);
const char* lk_last_error(void);
/* 1 when the library was compiled for sm_100a and the current device is compute capability 10.x */
int lk_device_ok(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long lk_launch_count(void);

/* per-call device timing inside the native step drivers: enable(1) clears and starts recording one CUDA-event pair per
 * sub-call; collect() synchronises the events and aggregates by entry-point name -> number of distinct names written
 * (names: cap strings of name_stride bytes; ms/flops/calls: cap entries).  flops = algorithmic 2*M*N*K of dense contractions. */
void lk_profile_enable(int on);
int lk_profile_collect(char* names, int name_stride, float* ms, double* flops, int* calls, int cap);

/* ---- (1) EmbeddingHub token gather — model/inputer/concat_inputer.py:105-113, simple_inputer.py:51-64,
 *      loader/embedding_hub.py:378-385 (aten::embedding + mask multiply + add) ------------------------- */
/* out[m,:] (+)= valid(m) ? table[ids[m],:] : 0, valid = mask ? mask[m]>0 : ids[m]>-1 */
int lk_gather_rows(const int64_t* ids, const int64_t* mask, const float* table, int64_t table_rows, float* out, int64_t M, int64_t E,
                   int accumulate, cudaStream_t stream);
/* gather straight into split-bf16 planes (A operand of the projection GEMM): hi/lo [M, ld], rows with ids<0 are zero */
int lk_gather_split_bf16(const int64_t* ids, const float* table, int64_t table_rows, void* hi, void* lo, int64_t M, int64_t E, int64_t ld,
                         cudaStream_t stream);
/* gather + masked pooling over S tokens (model/operators/pooling_operator.py:46-56): mode 0 mean, 1 max, 2 sum */
int lk_gather_pool(const int64_t* ids, const int64_t* mask, const float* table, int64_t table_rows, float* out, int64_t N, int64_t S, int64_t E,
                   int mode, cudaStream_t stream);
/* backward of the gather for trainable tables (autograd of aten::embedding): sorted-index segmented
 * scatter-add, deterministic.  dtable[ids[p],:] (+)= scale[p] * src[p / row_div,:] for valid p */
size_t lk_scatter_add_workspace_bytes(int64_t P, int64_t V, int64_t E);
int lk_scatter_add_sorted(const int64_t* ids, const int64_t* mask, const float* src, const float* scale, int64_t row_div,
                          float* dtable, int64_t P, int64_t V, int64_t E, int accumulate, void* workspace,
                          size_t workspace_bytes, cudaStream_t stream);

/* packed token ids of an item list from device-resident per-item token tables [N_items, S] (the Resampler's item cache,
 * loader/resampler.py:113-126): out_c[cu[n] + t] = table_c[items[n], t], t < cu[n+1]-cu[n].  tables / outs: HOST arrays of ncols (<= 4)
 * device pointers; items int64 [n], cu int32 [n+1] on the device. */
int lk_pack_item_tokens(const int64_t* const* tables, int64_t* const* outs, int ncols, const int64_t* items, const int32_t* cu, int64_t n,
                        int64_t S, cudaStream_t stream);

/* ---- row-sharded word table (BASELINE config 4; csrc/lk_shard.cu): the integer side of the all-to-all lookup on the device — dedup +
 *      owner bucketing into fixed-capacity send buckets (no split sizes to exchange, no host round trip), inverse index, owner-side gather.
 *      Row i lives on rank i % W at local index i / W.  Before lk_shard_plan: keys[slots] = -1, counts[W] = 0, overflow[0] = 0,
 *      send_ids[W*cap] = -1.  inverse[p] = row of the [W*cap, E] receive buffer holding position p's row (-1: unset position). */
int64_t lk_shard_hash_slots(int64_t P);
int lk_shard_plan(const int64_t* ids, int64_t P, int W, int64_t cap, int64_t* keys, int32_t* vals, int64_t slots, int32_t* counts,
                  int64_t* send_ids, int32_t* overflow, cudaStream_t stream);
int lk_shard_inverse(const int64_t* ids, int64_t P, const int64_t* keys, const int32_t* vals, int64_t slots, int64_t* inverse,
                     cudaStream_t stream);
int lk_shard_gather(const int64_t* ids, int64_t M, int W, const float* local, int64_t local_rows, int64_t E, float* out, cudaStream_t stream);

/* ---- gradient all-reduce over NVLink peer memory (csrc/lk_allreduce.cu): every rank's flat gradient bucket is peer-mapped (symmetric memory);
 *      this rank sums ITS 1/W slice over all W buckets (fixed order) and writes the (scaled) sum into all of them.  peer_ptrs: HOST array of W
 *      device pointers, index = rank.  flag_ptrs: HOST array of W device pointers to every rank's LK_ALLREDUCE_FLAG_WORDS int32 flag words
 *      (symmetric memory, zeroed once): the cross-rank rendezvous before (all gradients written) and after (all sums visible) then happen
 *      inside the launch, epoch = 1, 2, 3 ... per call and equal on all ranks.  flag_ptrs NULL: the caller brackets the launch with its own. */
#define LK_ALLREDUCE_FLAG_WORDS 4096
/* debugging aid: buf = device int64[148 * 6] (per block: SM clocks of launch->dependency, rendezvous, reduction, fence, rendezvous, total) or NULL */
int lk_allreduce_set_trace(void* buf);
/* multicast_ptr (with flag_ptrs only): this rank's NVSwitch multicast mapping of the W buckets, or NULL.  Non-NULL: the slice is reduced in the
 * switch (multimem.ld_reduce) and broadcast by it (multimem.st) — W times less NVLink traffic; the summation order is then the switch's. */
int lk_allreduce_p2p(void* const* peer_ptrs, void* const* flag_ptrs, void* multicast_ptr, int epoch, int rank, int W, int64_t n, float scale,
                     cudaStream_t stream);

/* ---- device-side Resampler (csrc/lk_resample.cu) — loader/resampler.py:139-259 taken to the device: from B impression rows to the id lists
 *      and offsets of a packed training batch in ONE launch.  Negatives = min(K, len) distinct positions of the user's negative list in random
 *      order + uniform item ids (Philox4x32-10 keyed by (seed, impression row)); histories are the users' valid clicks (padding is never
 *      materialised).  All tables are device arrays: imp_user / imp_pos [n_imps], CSR negatives (neg_off [n_users+1], neg_items), CSR histories
 *      (hist_off, hist_items), item_len [n_items] = valid tokens per item.  Outputs: items [n] (B*(K+1) candidates, then history items user by
 *      user), cu_items [n+1], cu_users [B+1], user_ids [B], meta[4] = {T token rows, n, longest item, longest history} (meta[0] = -1 and
 *      meta[1] = n when items_cap is too small). */
#define LK_MAX_NEG 16
int lk_resample_batch(const int64_t* rows, int64_t B, int K, uint64_t seed, const int64_t* imp_user, const int64_t* imp_pos,
                      const int64_t* neg_off, const int64_t* neg_items, const int64_t* hist_off, const int64_t* hist_items,
                      const int32_t* item_len, int64_t n_items, int64_t n_imps, int64_t n_users, int64_t* items, int32_t* cu_items,
                      int32_t* cu_users, int64_t* user_ids, int32_t* meta, int64_t items_cap, cudaStream_t stream);
/* HOST restatement of one impression's candidate draw (the same inline function the kernel executes; no device needed): cand_out[K+1] */
int lk_resample_reference(uint64_t seed, int64_t row, int64_t pos, const int64_t* negs, int64_t n_negs, int K, int64_t n_items,
                          int64_t* cand_out);

/* backward of the whole ConcatInputer embedding stage (concat_inputer.py:105-113 + embedding_hub.py:95-96) in one pass over
 * dx [T,D]:  dP = dx·dropout(seed)·(title id > -1) as split-bf16 planes [T, ld] (operand of the projection weight gradient),
 * g_bias [D] = column sums of dP, g_cat [n_cats, D] / g_special [n_special, D] = per-id sums of dx.  Deterministic.
 * With all three gradient pointers null the per-block partials [lk_concat_embed_bwd_blocks(T), (1+n_cats+n_special)*D] stay in `workspace`
 * (row layout: bias | categories | special tokens) for a later lk_colsum_finish_multi. */
int64_t lk_concat_embed_bwd_blocks(int64_t T);
size_t lk_concat_embed_bwd_workspace_bytes(int64_t T, int64_t D, int64_t n_cats, int64_t n_special);
int lk_concat_embed_bwd(const float* dx, const int64_t* title_ids, const int64_t* cat_ids, const int64_t* special_ids, int64_t T,
                        int64_t D, int64_t n_cats, int64_t n_special, float drop_p, uint64_t seed, void* dp_hi, void* dp_lo, int64_t ld,
                        float* g_bias, float* g_cat, float* g_special, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---- dense contractions — nn.Linear (loader/embedding_hub.py:95-96, model/operators/attention_operator.py:56,
 *      cnn_operator.py:62, model/common/attention.py:17-19), MHA in/out projections (attention_operator.py:32-37)
 *      act: 0 none, 1 tanh, 2 relu; rowmask (nullable) zeroes rows with mask<=0 after the activation */
int lk_linear_fwd(const float* X, const float* W, const float* bias, const int64_t* rowmask, float* Y, int64_t M, int64_t N,
                  int64_t K, int act, int accumulate, float drop_p, uint64_t seed, cudaStream_t stream);
/* dPre = dY * act'(Y) * dropout(seed) * rowmask, Y = saved epilogue output of lk_linear_fwd / lk_conv1d_fwd */
int lk_act_bwd(const float* dY, const float* Y, const int64_t* rowmask, float* dPre, int64_t M, int64_t N, int act, float drop_p,
               uint64_t seed, cudaStream_t stream);
/* out[i] = ids[i] > -1 (model/inputer/concat_inputer.py:108) */
int lk_valid_mask(const int64_t* ids, int64_t* out, int64_t n, cudaStream_t stream);
int lk_linear_bwd_data(const float* dY, const float* W, float* dX, int64_t M, int64_t N, int64_t K, int accumulate,
                       cudaStream_t stream);
size_t lk_linear_bwd_weight_workspace_bytes(int64_t M, int64_t N, int64_t K);
int lk_linear_bwd_weight(const float* dY, const float* X, float* dW, float* db, int64_t M, int64_t N, int64_t K,
                         int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* C[m,n] (+)= sum_z partial[z,m,n] in fixed z order (deterministic split reduction) */
int lk_splitk_reduce(const float* partial, float* C, int64_t M, int64_t N, int64_t ldc, int splits, int accumulate,
                     cudaStream_t stream);
/* Several column reductions in ONE launch: out[c] (+)= sum_i part[i*stride + c], i < nparts, fixed summation order (deterministic).
 * A training step leaves a dozen bias / small-table gradient partials behind (GEMM epilogues, the plane split, per-sequence sums of
 * the attention and pooling kernels); finishing them one launch each costs more than the work.  At most LK_COLSUM_MAX_JOBS jobs. */
#define LK_COLSUM_MAX_JOBS 16
typedef struct lk_colsum_job {
  const float* part;
  float* out;
  int64_t nparts, cols, stride;
  int accumulate;
} lk_colsum_job;
int lk_colsum_finish_multi(const lk_colsum_job* jobs, int n_jobs, cudaStream_t stream);
/* lk_split_bf16 with the column sums left as partials: part [ceil(rows/64), cols] (see lk_colsum_finish_multi) */
int lk_split_bf16_partial(const float* X, int64_t rows, int64_t cols, int64_t ld_in, void* hi, void* lo, int64_t ld_out, float* part,
                          cudaStream_t stream);
size_t lk_colsum_workspace_bytes(int64_t M, int64_t N);
int lk_colsum(const float* X, float* out, int64_t M, int64_t N, int accumulate, void* workspace, size_t workspace_bytes,
              cudaStream_t stream);

/* ---- tensor-core path for the same contractions (tcgen05.mma + TMEM + TMA, sm_100a).  Operands are split-bf16
 *      planes (hi = bf16(x), lo = bf16(x - hi)) so that three MMAs per k-step reproduce fp32 products to ~2^-17.
 *      lk_split_bf16: fp32 [rows, cols] (pitch ld_in) -> hi/lo [rows, ld_out] (transpose=0) or [cols, ld_out] (transpose=1).
 *      lk_tc_gemm: C[GM,GN] (+)= A·B with both operands K-major (a_mn=b_mn=0: A [GM,GK], B [GN,GK], reduction contiguous)
 *      or both MN-major (a_mn=b_mn=1: A [GK,GM], B [GK,GN]; the weight-gradient case, reduction over token rows). */
size_t lk_split_bf16_workspace_bytes(int64_t rows, int64_t cols);
/* colsum (nullable, transpose=0 only): also emit sum over rows of X (the bias gradient rides on the pass that reads dY) */
int lk_split_bf16(const float* X, int64_t rows, int64_t cols, int64_t ld_in, void* hi, void* lo, int64_t ld_out, int transpose,
                  float* colsum, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* several (weight) matrices in ONE launch: the parameters change every optimiser step, so their operand planes are rebuilt per
 * step — nine 7-us launches become one.  At most LK_SPLIT_MAX_SEGS segments; pad columns of the planes are zeroed. */
#define LK_SPLIT_MAX_SEGS 16
typedef struct lk_split_seg {
  const float* X;             /* fp32 [rows, cols], pitch ld_in */
  void* hi;                   /* bf16 [rows, ld_out] */
  void* lo;
  int64_t rows, cols, ld_in, ld_out;
} lk_split_seg;
int lk_split_bf16_multi(const lk_split_seg* segs, int n_segs, cudaStream_t stream);
/* Conv1d(k,'same') on the tensor cores (model/operators/cnn_operator.py:54-58): planes of the im2col image of X [rows, C] (sequences of S
 * rows): Xcol[r, j*C + c] = X[r + j - taps/2, c] inside r's sequence, else 0; hi/lo [rows, ld_out >= taps*C].  Y = Xcol·Wrᵀ is then one
 * lk_tc_gemm (bias / ReLU / dropout / row-mask epilogue), dX = im2col(dY)·Wdᵀ another, dW = dYᵀ·Xcol the third (Wr, Wd as for lk_conv1d_*). */
int lk_im2col_split_bf16(const float* X, int64_t rows, int64_t S, int64_t C, int taps, void* hi, void* lo, int64_t ld_out, cudaStream_t stream);
size_t lk_tc_gemm_workspace_bytes(int64_t GM, int64_t GN, int64_t GK);
/* Fused epilogue of lk_tc_gemm_ex, applied in this order to every result element r[m,n] of A·B:
 *   r += bias[n];  r = act(r);  r *= dropout(seed, m*GN+n);  r *= rowmask(m);  r += add_tab0[add_ids0[m], n] (+ add_tab1 ...)  where id > -1;
 *   r += C[m,n] when accumulate;  then r is stored to C (unless store_c_off), to the split-bf16 planes out_hi/out_lo (the next
 *   contraction's operand, produced without an extra pass) and summed over m into colsum[n] (bias gradients). */
typedef struct lk_gemm_epilogue {
  const float* bias;          /* [GN] */
  const int64_t* rowmask;     /* [GM]; keep row when > 0, or when > -1 if rowmask_is_ids (the row's token id) */
  int rowmask_is_ids;
  int act;                    /* 0 none, 1 tanh, 2 relu */
  float drop_p;
  uint64_t seed;
  int accumulate;             /* r += C */
  int store_c_off;            /* 1: do not store r to C (C only read for accumulate, or null) */
  const int64_t* add_ids0;    /* [GM] */
  const float* add_tab0;      /* [*, GN] */
  const int64_t* add_ids1;
  const float* add_tab1;
  void* out_hi;               /* bf16 [GM, ld_planes] */
  void* out_lo;
  int64_t ld_planes;
  float* colsum;              /* [GN] */
  float* colsum_part;         /* deferred column sums: [ceil(GM/128)*4, GN] partial sums are left here (no finishing launch);
                                 finish them later with lk_colsum_finish_multi.  Mutually exclusive with colsum. */
} lk_gemm_epilogue;
/* a_mn/b_mn: operand stored [rows, k] (0, K-major) or [k, rows] (1, MN-major); (1,0) is not instantiated */
int lk_tc_gemm_ex(const void* A_hi, const void* A_lo, int64_t lda, int a_mn, const void* B_hi, const void* B_lo, int64_t ldb, int b_mn,
                  float* C, int64_t ldc, int64_t GM, int64_t GN, int64_t GK, const lk_gemm_epilogue* ep, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream);
int lk_tc_gemm(const void* A_hi, const void* A_lo, int64_t lda, int a_mn, const void* B_hi, const void* B_lo, int64_t ldb, int b_mn,
               float* C, int64_t ldc, int64_t GM, int64_t GN, int64_t GK, const float* bias, const int64_t* rowmask, int act,
               float drop_p, uint64_t seed, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---- fused chain of up to three [M,256] x [256,256] contractions (csrc/lk_chain.cu): result g is the A operand of contraction g+1 and
 *      never leaves the SM (TMEM -> registers -> split-bf16 shared-memory operand).  Replaces, per encoder pass, the three nn.Linear calls
 *      out_proj -> linear -> additive W1 (model/operators/attention_operator.py:55-58, model/common/attention.py:31-33) and, backward, their
 *      three input-gradient contractions.  b_mn = 0: weights stored [N, K] (forward, y = x·Wᵀ); b_mn = 1: stored [K, N] (dX = dY·W).
 *      Per contraction: r = acc + bias + addsrc[m,:]; r = act(r); stored as fp32 / planes; column sums (bias gradients) left as
 *      [ceil(M/128)*4, 256] partials; rowdot_part[m, 0..3] = four partial sums of r[m,:]·dotvec (additive-attention scores). */
typedef struct lk_chain_stage {
  const void* w_hi;           /* bf16 planes of the 256 x 256 weight, pitch ldw */
  const void* w_lo;
  int64_t ldw;
  const float* bias;          /* [256] or null */
  const float* addsrc;        /* fp32 [M, 256] or null */
  int act;                    /* 0 none, 1 tanh */
  float* out_f32;             /* [M, 256] or null */
  void* out_hi;               /* bf16 [M, ld_planes] or null */
  void* out_lo;
  int64_t ld_planes;
  float* colsum_part;         /* [ceil(M/128)*4, 256] or null */
  const float* dotvec;        /* [256] or null */
  float* rowdot_part;         /* [M, 4] */
} lk_chain_stage;
int lk_tc_chain(const void* A_hi, const void* A_lo, int64_t lda, int64_t M, const lk_chain_stage* stages, int n_stages, int b_mn,
                cudaStream_t stream);
/* debug: with LK_CHAIN_TRACE=1 in the environment CTA 0 of every lk_tc_chain launch stamps clock64 at its pipeline events; this copies the
 * 4 x 256 stamps of the last launch to the host (synchronises) */
int lk_tc_chain_trace(long long* host_out, int cap);

/* ---- catalog scoring sweep with fused per-user top-k (csrc/lk_sweep.cu; BASELINE config 5): for every user the k best items of
 *      score[u, n] = <U[u,:], I[n,:]> (DotPredictor over the projected item table, model/predictors/dot_predictor.py:7-10) WITHOUT materialising
 *      the U x N scores.  Operands as split-bf16 planes [rows, 256].  The items are cut into `ranges` contiguous ranges (lk_sweep_ranges);
 *      out_val / out_idx [ranges, U, k] hold each range's k best per user (descending, ties towards the smaller item index; idx is the item's
 *      row in I), to be merged by a k-way selection — across ranges, and across ranks when the item table is row-sharded. */
#define LK_SWEEP_MAX_K 16
int64_t lk_sweep_ranges(int64_t U, int64_t N);
int lk_sweep_topk(const void* U_hi, const void* U_lo, int64_t ldu, int64_t U, const void* I_hi, const void* I_lo, int64_t ldi, int64_t N, int64_t D,
                  int k, int64_t ranges, float* out_val, int32_t* out_idx, cudaStream_t stream);

/* ---- row-wise pieces of the BERT-style blocks of TransformerOperator / FastformerOperator (transformers BertModel behind
 *      model/operators/transformer_operator.py:29-38; model/common/fastformer.py:146-226): y = LayerNorm(x + res) * w + b (res, xs nullable;
 *      xs receives x + res for the backward), exact erf GELU (mode 0 forward, 1: dy * gelu'(x)), counter-based dropout (backward = the same
 *      call on dy).  lk_layernorm_bwd leaves [lk_layernorm_bwd_parts(rows), 2*D] partials of (dw, db). */
int lk_layernorm_fwd(const float* x, const float* res, const float* w, const float* b, float* y, float* xs, float* mean, float* rstd, int64_t rows,
                     int64_t D, float eps, cudaStream_t stream);
int64_t lk_layernorm_bwd_parts(int64_t rows);
int lk_layernorm_bwd(const float* dy, const float* xs, const float* w, const float* mean, const float* rstd, float* dx, float* dwb_part, int64_t rows,
                     int64_t D, cudaStream_t stream);
int lk_gelu(const float* x, const float* dy, float* y, int64_t n, int mode, cudaStream_t stream);
int lk_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, cudaStream_t stream);

/* ---- MINER (csrc/lk_poly.cu): poly-attention pooling, model/operators/poly_attention_operator.py:52-56 (logits [B,S,C], mask [B,S], x [B,S,D] ->
 *      out [B,C,D]; w [B,C,S] saved; masked positions carry the reference's 1e-30 logit, not -inf) and the target-aware predictor,
 *      model/predictors/miner_predictor.py:30-64 (user, proj = gelu(Linear(user)) [B,C,D], items [B,K1,D] -> out [B,K1]; mode 0 weighted, 1 max, 2 mean;
 *      sc, wt [B,K1,C] saved). */
int lk_poly_pool_fwd(const float* logits, const int64_t* mask, const float* x, float* out, float* w, int64_t B, int64_t S, int64_t C, int64_t D,
                     cudaStream_t stream);
int lk_poly_pool_bwd(const float* dout, const float* w, const int64_t* mask, const float* x, float* dx, float* dlogits, int64_t B, int64_t S, int64_t C,
                     int64_t D, cudaStream_t stream);
int lk_miner_fwd(const float* user, const float* proj, const float* items, float* out, float* sc, float* wt, int64_t B, int64_t K1, int64_t C, int64_t D,
                 int mode, cudaStream_t stream);
int lk_miner_bwd(const float* dout, const float* user, const float* proj, const float* items, const float* sc, const float* wt, float* duser, float* dproj,
                 float* ditems, int64_t B, int64_t K1, int64_t C, int64_t D, int mode, cudaStream_t stream);

/* ---- Fastformer additive attention (csrc/lk_fast.cu; model/common/fastformer.py:96-143): per-head softmax pooling over the sequence with the
 *      reference's additive -10000 mask (score [B,S,H], mask [B,S], v [B,S,D] -> out [B,D]; w [B,H,S] saved), the broadcast product
 *      y[b,s,:] = a[b,s,:] * v[b,:] with its reduction dv[b,:] = sum_s dy*a, and an elementwise add. */
int lk_head_pool_fwd(const float* score, const int64_t* mask, const float* v, float* out, float* w, int64_t B, int64_t S, int64_t H, int64_t D, float scale,
                     cudaStream_t stream);
int lk_head_pool_bwd(const float* dout, const float* w, const float* v, float* dv, float* dscore, int64_t B, int64_t S, int64_t H, int64_t D, float scale,
                     cudaStream_t stream);
int lk_bcast_mul(const float* a, const float* v, float* y, int64_t B, int64_t S, int64_t D, cudaStream_t stream);
int lk_bcast_mul_dv(const float* dy, const float* a, float* dv, int64_t B, int64_t S, int64_t D, cudaStream_t stream);
int lk_add(const float* a, const float* b, float* y, int64_t n, cudaStream_t stream);

/* ---- GRU user encoder (LSTUR) — nn.GRU(1 layer, batch_first) over pack_padded_sequence, model/operators/gru_operator.py:25-54.
 *      gi [B,S,3H] = x W_ih^T + b_ih for every step (one contraction, the caller's); whhT [H,3H] = W_hh transposed; len [B] valid steps.
 *      Forward: last [B,H] = hidden state after step len-1; saved for the backward: hs [B,S,H], gates [B,S,3H] (r,z,n), hnp [B,S,H].
 *      Backward: dlast -> dgi, dgh [B,S,3H] (zero beyond len); weight / bias / input gradients are contractions of those. */
int lk_gru_fwd(const float* gi, const float* whhT, const float* bhh, const int32_t* len, float* last, float* hs, float* gates, float* hnp,
               int64_t B, int64_t S, int64_t H, cudaStream_t stream);
int lk_gru_bwd(const float* dlast, const float* whh, const int32_t* len, const float* hs, const float* gates, const float* hnp, float* dgi,
               float* dgh, int64_t B, int64_t S, int64_t H, cudaStream_t stream);

/* ---- NAML Conv1d(k,'same') as implicit-im2col GEMM — model/operators/cnn_operator.py:33-38,54-58.
 *      Wr[o, j*Cin+i] = W[o,i,j];  Wd[i, j*Cout+o] = W[o,i,taps-1-j];  rows = N*S token rows */
int lk_conv1d_fwd(const float* X, const float* Wr, const float* bias, const int64_t* rowmask, float* Y, int64_t rows,
                  int64_t S, int64_t Cin, int64_t Cout, int taps, int act, float drop_p, uint64_t seed, cudaStream_t stream);
int lk_conv1d_bwd_data(const float* dY, const float* Wd, float* dX, int64_t rows, int64_t S, int64_t Cin, int64_t Cout,
                       int taps, int accumulate, cudaStream_t stream);
size_t lk_conv1d_bwd_weight_workspace_bytes(int64_t rows, int64_t Cin, int64_t Cout, int taps);
int lk_conv1d_bwd_weight(const float* dY, const float* X, float* dWr, float* db, int64_t rows, int64_t S, int64_t Cin,
                         int64_t Cout, int taps, int accumulate, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream);

/* Sequence layouts shared by (2) and (3): dense [N, S, *] with an optional int64 validity mask [N,S] (the reference's padded
 * layout), or — when `cu` is non-null — PACKED rows with int32 cumulative offsets cu[N+1] (sequence n = rows cu[n]..cu[n+1]),
 * S then being an upper bound on the sequence length.  Packed execution skips the pad tokens / pad history slots that the
 * reference computes and masks away; results on valid positions are identical. */

/* ---- (2) NRMS multi-head self-attention core — nn.MultiheadAttention, attention_operator.py:49-55.
 *      qkv [rows,3D], ctx [rows,D], lse [rows,H]; head dim D/H in {8,16,32,64}.
 *      Forward: ctx as fp32 and/or as split-bf16 planes ctx_hi/ctx_lo [rows, D] (either may be null).
 *      Backward: needs the forward's fp32 ctx (D_i = dO_i·O_i) and lse; dqkv as fp32 and/or planes dq_hi/dq_lo [rows, 3D];
 *      colsum_part (nullable) [N, 3D]: per-sequence column sums of dqkv (in_proj bias gradient = their sum over N).
 *      Plane / column-sum outputs of the backward need S*H <= 416. ------------------------------------------------------- */
int lk_mha_fwd(const float* qkv, const int64_t* mask, const int32_t* cu, float* ctx, void* ctx_hi, void* ctx_lo, float* lse, int64_t N,
               int64_t S, int64_t D, int64_t H, float drop_p, uint64_t seed, cudaStream_t stream);
int lk_mha_bwd(const float* qkv, const int64_t* mask, const int32_t* cu, const float* ctx, const float* lse, const float* dctx,
               float* dqkv, void* dq_hi, void* dq_lo, float* colsum_part, int64_t N, int64_t S, int64_t D, int64_t H, float drop_p,
               uint64_t seed, cudaStream_t stream);

/* ---- (3) AdditiveAttention pooling — model/common/attention.py:31-38.  X [rows,D], Hd = tanh(W1 x + b1) [rows,A] ------ */
int lk_additive_pool_fwd(const float* X, const float* Hd, const float* w2, const int64_t* mask, const int32_t* cu, float* out,
                         float* alpha, int64_t N, int64_t S, int64_t D, int64_t A, cudaStream_t stream);
int lk_additive_pool_bwd(const float* X, const float* Hd, const float* w2, const float* alpha, const int32_t* cu, const float* dOut,
                         float* dX, float* dpre, float* dw2_part, int64_t N, int64_t S, int64_t D, int64_t A, int accumulate_dx,
                         cudaStream_t stream);
/* the same two kernels over the outputs of lk_tc_chain: rows X as split-bf16 planes (hi + lo) instead of fp32, scores as the chain's four
 * row-dot partials s_part [rows, 4] instead of Hd·w2 (packed rows only) */
int lk_additive_pool_fwd_planes(const void* X_hi, const void* X_lo, int64_t ldx, const float* s_part, const int32_t* cu, float* out, float* alpha,
                                int64_t N, int64_t S, int64_t D, cudaStream_t stream);
/* dpre goes to fp32 rows (dpre) and / or straight to the split-bf16 planes the next contractions read (dpre_hi/lo, pitch ld_dpre), in which
 * case dpre_colsum_part [N, A] (nullable) receives the per-sequence column sums of dpre (the b1 gradient, see lk_colsum_finish_multi) */
int lk_additive_pool_bwd_planes(const void* X_hi, const void* X_lo, int64_t ldx, const float* Hd, const float* w2, const float* alpha, const int32_t* cu,
                                const float* dOut, float* dX, float* dpre, void* dpre_hi, void* dpre_lo, int64_t ld_dpre, float* dpre_colsum_part,
                                float* dw2_part, int64_t N, int64_t S, int64_t D, int64_t A, cudaStream_t stream);
/* PoolingOperator on gathered embeddings — model/operators/pooling_operator.py:46-56 (mode 0 mean, 1 max) */
int lk_masked_pool(const float* X, const int64_t* mask, float* out, int64_t N, int64_t S, int64_t D, int mode,
                   cudaStream_t stream);
int lk_masked_mean_pool_bwd(const float* dOut, const int64_t* mask, float* dX, int64_t N, int64_t S, int64_t D,
                            cudaStream_t stream);

/* ---- (4) DotPredictor + loss — model/legommender.py:268-290, predictors/dot_predictor.py:7-10,
 *      legommender.py:114-118,254,263 -------------------------------------------------------------------- */
int lk_dot_scores(const float* U, const float* V, float* scores, int64_t B, int64_t C, int64_t D, cudaStream_t stream);
int lk_dot_ce_fwd(const float* U, const float* V, float* scores, float* probs, float* rowloss, float* loss, int64_t B,
                  int64_t C, int64_t D, cudaStream_t stream);
/* dloss: device scalar (the gradient of the mean loss), or NULL for 1 */
int lk_dot_ce_bwd(const float* U, const float* V, const float* probs, const float* dloss, float* dU, float* dV, int64_t B,
                  int64_t C, int64_t D, cudaStream_t stream);
int lk_dot_bwd(const float* U, const float* V, const float* dS, float* dU, float* dV, int64_t B, int64_t C, int64_t D,
               cudaStream_t stream);
int lk_dot_bce_fwd(const float* U, const float* V, const float* y, float* scores, float* rowloss, float* loss, int64_t B,
                   int64_t D, cudaStream_t stream);
int lk_dot_bce_bwd(const float* U, const float* V, const float* y, const float* scores, const float* dloss, float* dz,
                   float* dU, float* dV, int64_t B, int64_t D, cudaStream_t stream);

/* ---- (5) cached evaluation — model/legommender.py:153-157, 202-203; base_lego.py:373-394 -------------- */
int lk_cached_scores(const float* U, int64_t n_users, const float* I, int64_t n_items, const int64_t* uid, const int64_t* iid, float* out,
                     int64_t R, int64_t D, cudaStream_t stream);
int lk_index_rows(const float* table, int64_t table_rows, const int64_t* ids, float* out, int64_t R, int64_t D, cudaStream_t stream);

/* ---- group metrics of the evaluation phase — utils/metrics.py:88-160, 223-235, 313-369 (MetricPool.calculate with GAUC, MRR,
 *      NDCG@k).  groups: any int64 key (user id); rows of a group need not be contiguous.  ks: HOST array of nk <= 8 nDCG
 *      cut-offs; disc_prefix: DEVICE fp64 table, disc_prefix[p] = sum_{q<p} 1/log2(q+2), n_disc >= max(k)+1 entries.
 *      out (device, 3+nk doubles): mean GAUC, mean MRR, mean NDCG@ks[0..], number of groups.  per_group (nullable, device
 *      [(2+nk), R] floats, first `number of groups` entries of each row valid, groups in ascending key order). */
size_t lk_group_metrics_workspace_bytes(int64_t R, int nk);
int lk_group_metrics(const float* scores, const int64_t* labels, const int64_t* groups, int64_t R, const int32_t* ks, int nk,
                     const double* disc_prefix, int64_t n_disc, double* out, float* per_group, void* workspace,
                     size_t workspace_bytes, cudaStream_t stream);

/* ---- native training-step driver: the whole Legommender.forward + backward of the NRMS configuration
 *      (model/legommender.py:219-263 with config/model/nrms.yaml) over packed rows in ONE call; see csrc/lk_nrms_step.cu.
 *      offsets[22]: element offsets into params/grads (order documented at the definition). */
size_t lk_nrms_arena_bytes(int64_t T_max, int64_t N_max, int64_t B, int64_t C, int64_t D, int64_t A, int64_t E, int64_t H, int64_t n_cats,
                           int64_t n_special);
int lk_nrms_fwd_bwd(const int64_t* title_ids, const int64_t* cat_ids, const int64_t* special_ids, const int32_t* cu_items,
                    int64_t n_items, int64_t T, int64_t S_max, const int32_t* cu_users, int64_t B, int64_t C, int64_t H_max,
                    const float* glove_table, int64_t glove_rows, const float* params, float* grads, const int64_t* offsets, int64_t D, int64_t heads,
                    int64_t A, int64_t E, int64_t n_cats, int64_t n_special, float drop_embed, float drop_attn, uint64_t seed,
                    float* loss_out, float* scores_out, void* arena, size_t arena_bytes, cudaStream_t stream);
int lk_fill_f32(float* p, float value, int64_t n, cudaStream_t stream);

/* ---- optimiser — base_lego.py:198-204 (torch.optim.Adam defaults), one launch over a flat parameter buffer */
int lk_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 int64_t step, float grad_scale, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LEGOMMENDERS_B200_H_ */
