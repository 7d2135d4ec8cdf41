// Native NRMS training-step driver: ONE C-ABI call enqueues the whole forward + backward of
// Legommender.forward (model/legommender.py:219-263) for the NRMS configuration (config/model/nrms.yaml:
// AttentionOperator item + user encoders, ConcatInputer, DotPredictor, CrossEntropy with label 0) over PACKED rows.
//
// Why: with the kernels at a few tens of microseconds each, a Python/autograd-driven step is bound by the host (~25-40 us
// per launch).  This driver issues the same kernels in the same order from C++ (~120 launches, ~2-3 us each), owns no
// memory (the caller passes an arena), never synchronises, and writes every parameter gradient into a flat gradient
// buffer at caller-given offsets so that data-parallel training is one NCCL allreduce + one Adam launch afterwards.
//
// Row layout (see packing.py): item tokens are packed [T] with int32 offsets cu_items[n_items+1]; the first B*C items are
// the candidates, the rest are the valid history items user by user, so their encodings are directly the user encoder's
// packed token rows with offsets cu_users[B+1].
#include <cuda_bf16.h>
#include <stdlib.h>

#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {
namespace nrms {

struct Arena {
  char* base;
  size_t cap, off;
  bool ok;
  size_t high;   // high-water mark (the sizing pass runs the same allocation sequence with no memory behind it)
  void* take(size_t bytes) {
    size_t a = (off + 255) & ~(size_t)255;
    if (a + bytes > high) high = a + bytes;
    if (a + bytes > cap) { ok = false; off = a + bytes; return base; }
    off = a + bytes;
    return base + a;
  }
  float* f32(size_t n) { return (float*)take(n * 4); }
};

struct PlaneBuf {
  __nv_bfloat16 *hi, *lo;
  int64_t rows, cols, ld;
};

static inline int64_t r8(int64_t x) { return (x + 7) / 8 * 8; }

struct Ctx {
  Arena a;
  cudaStream_t st;
  void* ws;          // shared scratch for split-K partials / scatter / colsum
  size_t ws_bytes;
  int rc;
  bool dry;          // sizing pass: walk the allocations, launch nothing
  lk_colsum_job jobs[LK_COLSUM_MAX_JOBS];   // bias / small-table gradient reductions, all finished by ONE launch at the end of the step
  int n_jobs;
  // side stream: weight gradients are leaves of the backward graph (only the optimiser reads them).  Those of the user encoder
  // (1.7 k rows: 28-CTA contractions that are pure latency on the main stream) are issued here and joined once, at the end of the step.
  cudaStream_t side;
  cudaEvent_t fork_ev[12], join_ev;
  int n_fork;
  void* ws_side;
  size_t ws_side_bytes;
  bool side_used;
};

// Also the item encoder's (and the projection's) weight gradients go to the side stream: they fill the tails of the main stream's
// kernels (measured 1.152 -> 1.116 ms per step); their operands then stay allocated until the join.  LK_SIDE_ITEMS=0 keeps them inline.
static bool side_items() {
  static const bool on = [] { const char* e = getenv("LK_SIDE_ITEMS"); return !(e && e[0] == '0'); }();
  return on;
}
struct SideState { cudaStream_t st = nullptr; cudaEvent_t fork_ev[12]; cudaEvent_t join_ev; bool ok = false; bool items = false; };
static SideState& side_state() {
  static SideState s;
  if (!s.ok && !s.st) {
    const char* e = getenv("LK_SIDE_STREAM");
    const char* it = getenv("LK_SIDE_ITEMS");
    s.items = it && it[0] == '1';
    if (!(e && e[0] == '0') && cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) == cudaSuccess) {
      s.ok = true;
      for (auto& ev : s.fork_ev) s.ok = s.ok && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
      s.ok = s.ok && cudaEventCreateWithFlags(&s.join_ev, cudaEventDisableTiming) == cudaSuccess;
    }
    if (!s.st) s.st = (cudaStream_t)-1;     // tried once
  }
  return s;
}

// out[c] = sum_i part[i*stride + c]: queued; the partial buffer must stay untouched until the end of the step
static void defer_colsum(Ctx& c, const float* part, float* out, int64_t nparts, int64_t cols, int64_t stride) {
  if (c.n_jobs < LK_COLSUM_MAX_JOBS) c.jobs[c.n_jobs++] = lk_colsum_job{part, out, nparts, cols, stride, 0};
  else c.rc = c.rc ? c.rc : LK_ERR_ARG;
}

#define STEP_L(label, expr, flops)                              \
  do {                                                          \
    if (c.rc == 0 && !c.dry) {                                  \
      const bool prof_ = prof_enabled();                        \
      if (prof_) prof_begin(label, (double)(flops), c.st);      \
      c.rc = (expr);                                            \
      if (prof_) prof_end(c.st);                                \
    }                                                           \
  } while (0)
#define STEP(expr) STEP_L(#expr, expr, 0)

static PlaneBuf alloc_planes(Ctx& c, int64_t rows, int64_t cols) {
  PlaneBuf p;
  p.rows = rows; p.cols = cols; p.ld = r8(cols);
  p.hi = (__nv_bfloat16*)c.a.take((size_t)rows * p.ld * 2);
  p.lo = (__nv_bfloat16*)c.a.take((size_t)rows * p.ld * 2);
  return p;
}
// planes of X [rows, cols]; optional column sums into `colsum`
static PlaneBuf split(Ctx& c, const float* X, int64_t rows, int64_t cols, float* colsum = nullptr) {
  PlaneBuf p = alloc_planes(c, rows, cols);
  STEP(lk_split_bf16(X, rows, cols, cols, p.hi, p.lo, p.ld, 0, colsum, c.ws, c.ws_bytes, c.st));
  return p;
}
// General fused contraction: Y = epilogue(A · op(B)).  b_mn = 0: B stored [N,K] (forward); b_mn = 1: B stored [K,N] — the SAME planes of a
// weight W [out,in] serve Y = X·Wᵀ (b_mn = 0, N = out) and dX = dY·W (b_mn = 1, K = out, N = in), so no transposed copies exist.
static void gemm(Ctx& c, const char* what, const PlaneBuf& A, const PlaneBuf& B, int b_mn, float* Y, int64_t M, int64_t N, int64_t K,
                 const lk_gemm_epilogue& ep) {
  char label[40];
  snprintf(label, sizeof(label), "gemm_%s %ldx%ldx%ld", what, (long)M, (long)N, (long)K);
  STEP_L(label, lk_tc_gemm_ex(A.hi, A.lo, A.ld, 0, B.hi, B.lo, B.ld, b_mn, Y, N, M, N, K, &ep, c.ws, c.ws_bytes, c.st), 2.0 * M * N * K);
}
// dW[N,K] = dY[T,N]^T · X[T,K]
static void gemm_wgrad(Ctx& c, const PlaneBuf& dY, const PlaneBuf& X, float* dW, int64_t T, int64_t N, int64_t K, bool on_side = false) {
  char label[40];
  snprintf(label, sizeof(label), "gemm_wgrad %ldx%ldx%ld", (long)N, (long)K, (long)T);
  if (on_side && c.side && !c.dry && c.rc == 0 && !prof_enabled() && c.n_fork < 12) {
    // operands are complete at this point of the main stream: fork, contract on the side stream with its own scratch
    cudaEventRecord(c.fork_ev[c.n_fork], c.st);
    cudaStreamWaitEvent(c.side, c.fork_ev[c.n_fork], 0);
    c.n_fork++;
    c.side_used = true;
    c.rc = lk_tc_gemm(dY.hi, dY.lo, dY.ld, 1, X.hi, X.lo, X.ld, 1, dW, K, N, K, T, nullptr, nullptr, 0, 0.f, 0, 0, c.ws_side, c.ws_side_bytes,
                      c.side);
    return;
  }
  STEP_L(label, lk_tc_gemm(dY.hi, dY.lo, dY.ld, 1, X.hi, X.lo, X.ld, 1, dW, K, N, K, T, nullptr, nullptr, 0, 0.f, 0, 0, c.ws, c.ws_bytes,
                           c.st), 2.0 * T * N * K);
}
static lk_gemm_epilogue ep_planes(const PlaneBuf& out, const float* bias = nullptr, float* colsum_part = nullptr) {
  lk_gemm_epilogue ep = {};
  ep.bias = bias; ep.out_hi = out.hi; ep.out_lo = out.lo; ep.ld_planes = out.ld; ep.colsum_part = colsum_part;
  return ep;
}
static int64_t gemm_colsum_parts(int64_t M) { return (M + 127) / 128 * 4; }   // per-(128-row tile, lane quarter) partials of lk_tc_gemm
static int64_t split_colsum_parts(int64_t M) { return (M + 63) / 64; }        // per-64-row-block partials of lk_split_bf16_partial

struct EncWeights {   // one AttentionOperator
  const float *in_w, *in_b, *out_w, *out_b, *lin_w, *lin_b, *w1, *b1, *w2;
  float *g_in_w, *g_in_b, *g_out_w, *g_out_b, *g_lin_w, *g_lin_b, *g_w1, *g_b1, *g_w2;
};
struct EncPlanes { PlaneBuf in_w, out_w, lin_w, w1; };
struct EncSaved {
  PlaneBuf xp, ctxp, outp, linp;
  float *qkv, *ctx, *lse, *lin, *hid, *alpha, *rep, *spart;
  bool chained;        // out -> linear -> W1 ran as ONE fused kernel (lk_tc_chain): no fp32 `lin`, scores as row-dot partials
  const int32_t* cu;
  int64_t T, N, S;
  uint64_t seed;
};

// operand planes of a weight matrix: the conversion is queued and all weights of the step are converted by ONE launch
struct SplitQueue { lk_split_seg seg[LK_SPLIT_MAX_SEGS]; int n = 0; };
static PlaneBuf queue_split(Ctx& c, SplitQueue& q, const float* W, int64_t rows, int64_t cols) {
  PlaneBuf p = alloc_planes(c, rows, cols);
  if (q.n < LK_SPLIT_MAX_SEGS) q.seg[q.n++] = lk_split_seg{W, p.hi, p.lo, rows, cols, cols, p.ld};
  else c.rc = c.rc ? c.rc : LK_ERR_ARG;
  return p;
}
static EncPlanes weight_planes(Ctx& c, SplitQueue& q, const EncWeights& w, int64_t D, int64_t A) {
  EncPlanes p;
  p.in_w = queue_split(c, q, w.in_w, 3 * D, D);
  p.out_w = queue_split(c, q, w.out_w, D, D);
  p.lin_w = queue_split(c, q, w.lin_w, D, D);
  p.w1 = queue_split(c, q, w.w1, A, D);
  return p;
}

// attention_operator.py:46-59 over packed rows: s.xp = planes of the input rows [T,D].  Tensors that only feed the next
// contraction (ctx, out) leave their producer as split-bf16 planes; fp32 copies exist only where a SIMT kernel reads them.
// The fused chain kernel is specialised to 256-wide contractions (hidden_size = additive_hidden_size = 256, the shipped nrms.yaml);
// other widths, and LK_CHAIN=0, take the GEMM-by-GEMM path.
static bool use_chain(int64_t D, int64_t A) {
  static const bool on = [] { const char* e = getenv("LK_CHAIN"); return !(e && e[0] == '0'); }();
  return on && D == 256 && A == 256;
}
static lk_chain_stage chain_stage(const PlaneBuf& W) {
  lk_chain_stage st = {};
  st.w_hi = W.hi; st.w_lo = W.lo; st.ldw = W.ld;
  return st;
}

static void enc_fwd(Ctx& c, EncSaved& s, const EncWeights& w, const EncPlanes& wp, int64_t D, int64_t H, int64_t A, float drop_attn) {
  const int64_t T = s.T, N = s.N;
  s.chained = use_chain(D, A);
  s.qkv = c.a.f32(T * 3 * D);
  lk_gemm_epilogue ep = {};
  ep.bias = w.in_b;
  gemm(c, "qkv", s.xp, wp.in_w, 0, s.qkv, T, 3 * D, D, ep);
  s.ctx = c.a.f32(T * D);                      // fp32 ctx is kept for the backward's D_i = dO·O
  s.ctxp = alloc_planes(c, T, D);
  s.lse = c.a.f32(T * H);
  STEP(lk_mha_fwd(s.qkv, nullptr, s.cu, s.ctx, s.ctxp.hi, s.ctxp.lo, s.lse, N, s.S, D, H, drop_attn, s.seed, c.st));
  if (s.chained) {
    s.outp = alloc_planes(c, T, D);
    s.linp = alloc_planes(c, T, D);
    s.hid = c.a.f32(T * A);
    s.spart = c.a.f32(T * 4);
    s.alpha = c.a.f32(T);
    s.rep = c.a.f32(N * D);
    s.lin = nullptr;
    lk_chain_stage st[3] = {chain_stage(wp.out_w), chain_stage(wp.lin_w), chain_stage(wp.w1)};
    st[0].bias = w.out_b; st[0].out_hi = s.outp.hi; st[0].out_lo = s.outp.lo; st[0].ld_planes = s.outp.ld;
    st[1].bias = w.lin_b; st[1].out_hi = s.linp.hi; st[1].out_lo = s.linp.lo; st[1].ld_planes = s.linp.ld;
    st[2].bias = w.b1; st[2].act = 1; st[2].out_f32 = s.hid; st[2].dotvec = w.w2; st[2].rowdot_part = s.spart;
    char label[40];
    snprintf(label, sizeof(label), "chain_fwd %ldx256x256 x3", (long)T);
    STEP_L(label, lk_tc_chain(s.ctxp.hi, s.ctxp.lo, s.ctxp.ld, T, st, 3, 0, c.st), 3 * 2.0 * T * D * D);
    STEP(lk_additive_pool_fwd_planes(s.linp.hi, s.linp.lo, s.linp.ld, s.spart, s.cu, s.rep, s.alpha, N, s.S, D, c.st));
    return;
  }
  s.outp = alloc_planes(c, T, D);
  gemm(c, "out", s.ctxp, wp.out_w, 0, nullptr, T, D, D, ep_planes(s.outp, w.out_b));
  s.lin = c.a.f32(T * D);
  s.linp = alloc_planes(c, T, D);
  gemm(c, "lin", s.outp, wp.lin_w, 0, s.lin, T, D, D, ep_planes(s.linp, w.lin_b));
  s.hid = c.a.f32(T * A);
  ep = {};
  ep.bias = w.b1; ep.act = 1 /*tanh*/;
  gemm(c, "w1", s.linp, wp.w1, 0, s.hid, T, A, D, ep);
  s.alpha = c.a.f32(T);
  s.rep = c.a.f32(N * D);
  STEP(lk_additive_pool_fwd(s.lin, s.hid, w.w2, nullptr, s.cu, s.rep, s.alpha, N, s.S, D, A, c.st));
}

// backward of enc_fwd: drep [N,D] -> dX [T,D] (or nothing when dX == null); parameter gradients into w.g_*
// the five bias-type gradients of an encoder are left as partial sums in buffers that live until the end of the step
struct EncPartials { float *dw2p, *b1p, *linbp, *outbp, *binp; };
static EncPartials alloc_partials(Ctx& c, int64_t T, int64_t N, int64_t D, int64_t A) {
  EncPartials q;
  q.dw2p = c.a.f32(N * A);                          // per-sequence partials of dw2
  q.b1p = c.a.f32((split_colsum_parts(T) > N ? split_colsum_parts(T) : N) * A);     // per-sequence (chained) or per-64-row-block partials
  q.linbp = c.a.f32(gemm_colsum_parts(T) * D);
  q.outbp = c.a.f32(gemm_colsum_parts(T) * D);
  q.binp = c.a.f32(N * 3 * D);                      // per-sequence column sums of dqkv
  return q;
}

static void enc_bwd(Ctx& c, const EncSaved& s, const EncWeights& w, const EncPlanes& wp, const EncPartials& q, int64_t D, int64_t H, int64_t A,
                    float drop_attn, const float* drep, float* dX, bool side = false) {
  const int64_t T = s.T, N = s.N;
  const size_t mark = c.a.off;
  float* dlin = c.a.f32(T * D);
  PlaneBuf dprep = alloc_planes(c, T, A);
  if (s.chained) {
    // dpre leaves the pooling backward as operand planes + per-sequence column sums: no fp32 dpre, no split pass
    STEP(lk_additive_pool_bwd_planes(s.linp.hi, s.linp.lo, s.linp.ld, s.hid, w.w2, s.alpha, s.cu, drep, dlin, nullptr, dprep.hi, dprep.lo, dprep.ld,
                                     q.b1p, q.dw2p, N, s.S, D, A, c.st));
    defer_colsum(c, q.b1p, w.g_b1, N, A, A);
  } else {
    float* dpre = c.a.f32(T * A);
    STEP(lk_additive_pool_bwd(s.lin, s.hid, w.w2, s.alpha, s.cu, drep, dlin, dpre, q.dw2p, N, s.S, D, A, 0, c.st));
    STEP(lk_split_bf16_partial(dpre, T, A, A, dprep.hi, dprep.lo, dprep.ld, q.b1p, c.st));
    defer_colsum(c, q.b1p, w.g_b1, split_colsum_parts(T), A, A);
  }
  defer_colsum(c, q.dw2p, w.g_w2, N, A, A);
  gemm_wgrad(c, dprep, s.linp, w.g_w1, T, A, D, side);
  if (s.chained) {
    // dpre -> dlin (= alpha*drep + dpre·W1) -> dout -> dctx in one kernel; the two intermediate tiles leave only as the planes the weight
    // gradients contract and as column-sum partials (bias gradients)
    PlaneBuf dlinp = alloc_planes(c, T, D), doutp = alloc_planes(c, T, D);
    float* dctx = c.a.f32(T * D);
    lk_chain_stage st[3] = {chain_stage(wp.w1), chain_stage(wp.lin_w), chain_stage(wp.out_w)};
    st[0].addsrc = dlin; st[0].out_hi = dlinp.hi; st[0].out_lo = dlinp.lo; st[0].ld_planes = dlinp.ld; st[0].colsum_part = q.linbp;
    st[1].out_hi = doutp.hi; st[1].out_lo = doutp.lo; st[1].ld_planes = doutp.ld; st[1].colsum_part = q.outbp;
    st[2].out_f32 = dctx;
    char label[40];
    snprintf(label, sizeof(label), "chain_bwd %ldx256x256 x3", (long)T);
    STEP_L(label, lk_tc_chain(dprep.hi, dprep.lo, dprep.ld, T, st, 3, 1, c.st), 3 * 2.0 * T * D * D);
    defer_colsum(c, q.linbp, w.g_lin_b, gemm_colsum_parts(T), D, D);
    defer_colsum(c, q.outbp, w.g_out_b, gemm_colsum_parts(T), D, D);
    gemm_wgrad(c, dlinp, s.outp, w.g_lin_w, T, D, D, side);
    gemm_wgrad(c, doutp, s.ctxp, w.g_out_w, T, D, D, side);
    PlaneBuf dqkvp = alloc_planes(c, T, 3 * D);
    STEP(lk_mha_bwd(s.qkv, nullptr, s.cu, s.ctx, s.lse, dctx, nullptr, dqkvp.hi, dqkvp.lo, q.binp, N, s.S, D, H, drop_attn, s.seed, c.st));
    defer_colsum(c, q.binp, w.g_in_b, N, 3 * D, 3 * D);
    gemm_wgrad(c, dqkvp, s.xp, w.g_in_w, T, 3 * D, D, side);
    lk_gemm_epilogue ep0 = {};
    if (dX) gemm(c, "dx", dqkvp, wp.in_w, 1, dX, T, D, 3 * D, ep0);
    if (!side) c.a.off = mark;
    return;
  }
  // dlin = alpha*drep (already in dlin) + dpre·W1 -> only its planes and column sums are needed downstream
  PlaneBuf dlinp = alloc_planes(c, T, D);
  lk_gemm_epilogue ep = ep_planes(dlinp, nullptr, q.linbp);
  ep.accumulate = 1; ep.store_c_off = 1;
  gemm(c, "dlin", dprep, wp.w1, 1, dlin, T, D, A, ep);
  defer_colsum(c, q.linbp, w.g_lin_b, gemm_colsum_parts(T), D, D);
  gemm_wgrad(c, dlinp, s.outp, w.g_lin_w, T, D, D, side);
  PlaneBuf doutp = alloc_planes(c, T, D);
  gemm(c, "dout", dlinp, wp.lin_w, 1, nullptr, T, D, D, ep_planes(doutp, nullptr, q.outbp));
  defer_colsum(c, q.outbp, w.g_out_b, gemm_colsum_parts(T), D, D);
  gemm_wgrad(c, doutp, s.ctxp, w.g_out_w, T, D, D, side);
  float* dctx = dlin;   // dlin is dead once its planes exist
  ep = {};
  gemm(c, "dctx", doutp, wp.out_w, 1, dctx, T, D, D, ep);
  PlaneBuf dqkvp = alloc_planes(c, T, 3 * D);
  STEP(lk_mha_bwd(s.qkv, nullptr, s.cu, s.ctx, s.lse, dctx, nullptr, dqkvp.hi, dqkvp.lo, q.binp, N, s.S, D, H, drop_attn, s.seed, c.st));
  defer_colsum(c, q.binp, w.g_in_b, N, 3 * D, 3 * D);
  gemm_wgrad(c, dqkvp, s.xp, w.g_in_w, T, 3 * D, D, side);
  if (dX) gemm(c, "dx", dqkvp, wp.in_w, 1, dX, T, D, 3 * D, ep);
  // all temporaries of this backward are dead (stream order keeps reuse safe) — unless side-stream contractions still read them:
  // then they stay allocated until the join at the end of the step (the sizing pass takes the same branch)
  if (!side) c.a.off = mark;
}

}  // namespace nrms
}  // namespace lk

using namespace lk;
using namespace lk::nrms;

// shared scratch: the largest of the split-reduction partials, scatter-add and column-sum workspaces of one step
static size_t scratch_bytes(int64_t T, int64_t N, int64_t D, int64_t A, int64_t E, int64_t n_cats, int64_t n_special) {
  size_t m = 0;
  auto up = [&](size_t v) { if (v > m) m = v; };
  up(lk_tc_gemm_workspace_bytes(3 * D, D, T));
  up(lk_tc_gemm_workspace_bytes(D, D, T));
  up(lk_tc_gemm_workspace_bytes(A, D, T));
  up(lk_tc_gemm_workspace_bytes(D, E, T));
  up(lk_scatter_add_workspace_bytes(T, 64, D));
  up(lk_scatter_add_workspace_bytes(T, 24, D));   // the small-table path's block partials grow with V (category / special tables)
  up(lk_split_bf16_workspace_bytes(T, 3 * D));
  up(lk_colsum_workspace_bytes(N, A));
  up(lk_colsum_workspace_bytes(N, 3 * D));
  up(lk_concat_embed_bwd_workspace_bytes(T, D, n_cats, n_special));
  up(lk_tc_gemm_workspace_bytes(T, 3 * D, D));     // column-sum partials of the fused epilogues
  up(lk_tc_gemm_workspace_bytes(T, D, 3 * D));
  return m + (1 << 20);
}

static int nrms_run(bool dry, size_t* high_out, const int64_t* title_ids, const int64_t* cat_ids, const int64_t* special_ids,
                    const int32_t* cu_items, int64_t n_items, int64_t T, int64_t S_max, const int32_t* cu_users, int64_t B, int64_t C,
                    int64_t H_max, const float* glove_table, int64_t glove_rows, const float* params, float* grads, const int64_t* offsets, int64_t D,
                    int64_t heads, int64_t A, int64_t E, int64_t n_cats, int64_t n_special, float drop_embed, float drop_attn,
                    uint64_t seed, float* loss_out, float* scores_out, void* arena, size_t arena_bytes, cudaStream_t st) {
  Ctx c;
  c.a = Arena{(char*)arena, arena_bytes, 0, true, 0};
  c.st = st;
  c.rc = 0;
  c.dry = dry;
  c.n_jobs = 0;
  c.ws_bytes = scratch_bytes(T > 0 ? T : 1, n_items, D, A, E, n_cats, n_special);
  c.ws = c.a.take(c.ws_bytes);
  const int64_t Tu_ = n_items - B * C > 0 ? n_items - B * C : 1;
  c.ws_side_bytes = lk_tc_gemm_workspace_bytes(3 * D, D, Tu_) + (1 << 16);      // the largest user-encoder weight gradient (in_proj)
  {
    size_t o = lk_tc_gemm_workspace_bytes(D, D, Tu_), a1 = lk_tc_gemm_workspace_bytes(A, D, Tu_);
    if (o + (1 << 16) > c.ws_side_bytes) c.ws_side_bytes = o + (1 << 16);
    if (a1 + (1 << 16) > c.ws_side_bytes) c.ws_side_bytes = a1 + (1 << 16);
  }
  if (side_items()) {
    const int64_t Tt = T > 0 ? T : 1;
    size_t m = lk_tc_gemm_workspace_bytes(3 * D, D, Tt);
    auto up = [&](size_t v) { if (v > m) m = v; };
    up(lk_tc_gemm_workspace_bytes(D, D, Tt)); up(lk_tc_gemm_workspace_bytes(A, D, Tt)); up(lk_tc_gemm_workspace_bytes(D, E, Tt));
    if (m + (1 << 16) > c.ws_side_bytes) c.ws_side_bytes = m + (1 << 16);
  }
  c.ws_side = c.a.take(c.ws_side_bytes);
  c.side = nullptr; c.n_fork = 0; c.side_used = false;
  if (!dry) {
    SideState& ss = side_state();
    if (ss.ok) {
      c.side = ss.st;
      for (int i = 0; i < 12; i++) c.fork_ev[i] = ss.fork_ev[i];
      c.join_ev = ss.join_ev;
    }
  }

  auto P = [&](int i) { return params + offsets[i]; };
  auto G = [&](int i) { return grads + offsets[i]; };
  EncWeights wi{P(4), P(5), P(6), P(7), P(8), P(9), P(10), P(11), P(12), G(4), G(5), G(6), G(7), G(8), G(9), G(10), G(11), G(12)};
  EncWeights wu{P(13), P(14), P(15), P(16), P(17), P(18), P(19), P(20), P(21),
                G(13), G(14), G(15), G(16), G(17), G(18), G(19), G(20), G(21)};
  const uint64_t s_embed = seed * 4 + 0, s_item = seed * 4 + 1, s_user = seed * 4 + 2;

  // ---- weights -> split-bf16 planes (parameters change every step) ------------------------------------------------
  SplitQueue wq;
  EncPlanes pi = weight_planes(c, wq, wi, D, A);
  EncPlanes pu = weight_planes(c, wq, wu, D, A);
  PlaneBuf pg = queue_split(c, wq, P(0), D, E);
  STEP(lk_split_bf16_multi(wq.seg, wq.n, st));

  // ---- embedding stage (concat_inputer.py:92-114 + embedding_hub.py:95-96) -------------------------------------------
  // x[t] = valid(title)·dropout(W·glove[title] + b) + category[cat] + special[sp]  — one contraction whose epilogue applies the row
  // mask (title id > -1), adds the two small-table rows and writes x directly as the item encoder's operand planes
  PlaneBuf gp = alloc_planes(c, T, E);
  STEP(lk_gather_split_bf16(title_ids, glove_table, glove_rows, gp.hi, gp.lo, T, E, gp.ld, st));

  // ---- item encoder over all packed items, user encoder over the packed history encodings ---------------------------------
  EncSaved si;
  si.T = T; si.N = n_items; si.S = S_max; si.cu = cu_items; si.seed = s_item;
  si.xp = alloc_planes(c, T, D);
  {
    lk_gemm_epilogue ep = ep_planes(si.xp, P(1));
    ep.rowmask = title_ids; ep.rowmask_is_ids = 1; ep.drop_p = drop_embed; ep.seed = s_embed;
    ep.add_ids0 = cat_ids; ep.add_tab0 = P(2); ep.add_ids1 = special_ids; ep.add_tab1 = P(3);
    gemm(c, "embed", gp, pg, 0, nullptr, T, D, E, ep);
  }
  enc_fwd(c, si, wi, pi, D, heads, A, drop_attn);

  const int64_t Tu = n_items - B * C;
  EncSaved su;
  su.T = Tu; su.N = B; su.S = H_max; su.cu = cu_users; su.seed = s_user;
  su.xp = split(c, si.rep + B * C * D, Tu, D);
  enc_fwd(c, su, wu, pu, D, heads, A, drop_attn);

  // ---- DotPredictor + CrossEntropy(label 0) (legommender.py:252-254, 268-283) ----------------------------------------------
  float* scores = scores_out ? scores_out : c.a.f32(B * C);
  float* probs = c.a.f32(B * C);
  float* rowloss = c.a.f32(B);
  STEP(lk_dot_ce_fwd(su.rep, si.rep, scores, probs, rowloss, loss_out, B, C, D, st));

  // ---- backward ---------------------------------------------------------------------------------------------------------
  float* drep = c.a.f32(n_items * D);
  float* duser = c.a.f32(B * D);
  STEP(lk_dot_ce_bwd(su.rep, si.rep, probs, nullptr /* dloss = 1 */, duser, drep, B, C, D, st));          // dV -> drep[0 : B*C]
  const EncPartials qu = alloc_partials(c, Tu, B, D, A), qi = alloc_partials(c, T, n_items, D, A);
  const size_t eb_bytes = lk_concat_embed_bwd_workspace_bytes(T, D, n_cats, n_special);
  float* ebp = (float*)c.a.take(eb_bytes);                                               // per-block partials of the embedding-stage gradients
  enc_bwd(c, su, wu, pu, qu, D, heads, A, drop_attn, duser, drep + B * C * D, true);   // dX_u -> drep[B*C :]; weight gradients on the side stream
  float* dx = c.a.f32(T * D);
  enc_bwd(c, si, wi, pi, qi, D, heads, A, drop_attn, drep, dx, side_items());

  // embedding stage backward in one pass over dx: small-table gradients, dP planes, bias gradient (partials, finished below)
  PlaneBuf dpp = alloc_planes(c, T, D);
  STEP(lk_concat_embed_bwd(dx, title_ids, cat_ids, special_ids, T, D, n_cats, n_special, drop_embed, s_embed, dpp.hi, dpp.lo, dpp.ld, nullptr,
                           nullptr, nullptr, ebp, eb_bytes, st));
  {
    const int64_t nblk = lk_concat_embed_bwd_blocks(T), stride = (1 + n_cats + n_special) * D;
    defer_colsum(c, ebp, G(1), nblk, D, stride);
    defer_colsum(c, ebp + D, G(2), nblk, n_cats * D, stride);
    defer_colsum(c, ebp + (1 + n_cats) * D, G(3), nblk, n_special * D, stride);
  }
  gemm_wgrad(c, dpp, gp, G(0), T, D, E, side_items());
  STEP(lk_colsum_finish_multi(c.jobs, c.n_jobs, st));       // every bias / small-table gradient of the step in one launch
  if (c.side_used) {                                        // join: everything after this call on `st` sees the side stream's gradients
    cudaEventRecord(c.join_ev, c.side);
    cudaStreamWaitEvent(st, c.join_ev, 0);
  }

  if (high_out) *high_out = c.a.high;
  LK_REQUIRE(c.a.ok, LK_ERR_ARG, "lk_nrms_fwd_bwd: arena too small (%zu bytes given, %zu needed)", arena_bytes, c.a.high);
  return c.rc;
}

extern "C" {

// exact: the sizing pass walks the same allocation sequence as a real step (no launches, no memory behind the arena)
size_t lk_nrms_arena_bytes(int64_t T_max, int64_t N_max, int64_t B, int64_t C, int64_t D, int64_t A, int64_t E, int64_t H, int64_t n_cats,
                           int64_t n_special) {
  static const int64_t zeros[22] = {0};
  size_t high = 0;
  nrms_run(true, &high, nullptr, nullptr, nullptr, nullptr, N_max, T_max, 1, nullptr, B, C, 1, nullptr, 0, nullptr, nullptr, zeros, D, H, A,
           E, n_cats, n_special, 0.f, 0.f, 0, nullptr, nullptr, nullptr, ~(size_t)0 >> 1, nullptr);
  return high + 4096;
}

// offsets[]: element offsets into params / grads for, in order:
//   0 glove.linear.weight [D,E]   1 glove.linear.bias [D]   2 category.weight [n_cats,D]   3 special.weight [n_special,D]
//   4..12  item_op: in_proj_weight, in_proj_bias, out_proj.weight, out_proj.bias, linear.weight, linear.bias,
//                   additive.encoder.0.weight, additive.encoder.0.bias, additive.encoder.2.weight
//   13..21 user_op: same nine
int lk_nrms_fwd_bwd(const int64_t* title_ids, const int64_t* cat_ids, const int64_t* special_ids, const int32_t* cu_items,
                    int64_t n_items, int64_t T, int64_t S_max, const int32_t* cu_users, int64_t B, int64_t C, int64_t H_max,
                    const float* glove_table, int64_t glove_rows, const float* params, float* grads, const int64_t* offsets, int64_t D, int64_t heads,
                    int64_t A, int64_t E, int64_t n_cats, int64_t n_special, float drop_embed, float drop_attn, uint64_t seed,
                    float* loss_out, float* scores_out, void* arena, size_t arena_bytes, cudaStream_t st) {
  LK_REQUIRE(n_items >= B * C && B > 0 && C > 0, LK_ERR_ARG, "lk_nrms_fwd_bwd: the first B*C items must be the candidates");
  LK_REQUIRE(D % 8 == 0 && A % 8 == 0 && E % 4 == 0 && D % heads == 0, LK_ERR_SHAPE, "lk_nrms_fwd_bwd: unsupported dims");
  {   // walk the allocation sequence with the real shapes BEFORE anything is launched: an arena that is too small must never be used
    size_t need = 0;
    nrms_run(true, &need, nullptr, nullptr, nullptr, nullptr, n_items, T, S_max, nullptr, B, C, H_max, nullptr, 0, params, grads, offsets, D, heads,
             A, E, n_cats, n_special, drop_embed, drop_attn, seed, nullptr, scores_out, nullptr, ~(size_t)0 >> 1, nullptr);
    LK_REQUIRE(need <= arena_bytes, LK_ERR_ARG, "lk_nrms_fwd_bwd: arena too small (%zu bytes given, %zu needed)", arena_bytes, need);
  }
  return nrms_run(false, nullptr, title_ids, cat_ids, special_ids, cu_items, n_items, T, S_max, cu_users, B, C, H_max, glove_table, glove_rows, params,
                  grads, offsets, D, heads, A, E, n_cats, n_special, drop_embed, drop_attn, seed, loss_out, scores_out, arena,
                  arena_bytes, st);
}

}  // extern "C"
