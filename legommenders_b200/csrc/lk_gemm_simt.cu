// FP32 SIMT GEMM family (exact-fp32 contractions; the tcgen05 path in lk_gemm_tc.cu is the fast one).
//
// One templated 128x128x16 register-tiled kernel serves the three contraction layouts of the
// hot path (reference call sites: nn.Linear / MHA in_proj,out_proj / Conv1d in
// model/operators/{attention,cnn}_operator.py, loader/embedding_hub.py:95-96):
//   forward      Y[M,N]  = X[M,K] · W[N,K]^T          A k-contiguous, B k-contiguous
//   grad-input   dX[M,K] = dY[M,N] · W[N,K]           A k-contiguous, B n-contiguous
//   grad-weight  dW[N,K] = dY[M,N]^T · X[M,K]         A m-contiguous, B n-contiguous (split over M)
// plus an implicit-im2col view ("conv shift") of a row-major token matrix so that Conv1d(k,'same')
// (cnn_operator.py:33-38,54) is the same kernel with no materialised im2col.
#include "lk_common.cuh"
#include "../../include/legommenders_b200.h"

namespace lk {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, NT = 256;
constexpr int SPAD = 4;

struct ConvView {   // virtual [rows, taps*C] view of a row-major [rows, C] matrix of S-long sequences
  int S, C, taps, pad;
};

struct GemmParams {
  const float* A; const float* B; float* C;
  int M, N, K;
  int lda, ldb, ldc;
  const float* bias;         // [N] or null
  const int64_t* rowmask;    // [M] or null: rows with mask<=0 are written as 0 (mask applied after act)
  int act;                   // 0 none, 1 tanh, 2 relu
  int accumulate;            // C += result
  int kchunk;                // K range per blockIdx.z
  float* partial;            // split-K workspace [splits, M, N] or null
  ConvView conv;             // conv.taps > 0 => the row-major operand is a conv view
  float drop_p;              // dropout on the epilogue output (after act, before rowmask); 0 = off
  unsigned long long seed;
};

// 4 consecutive elements along the contiguous dim of a row-major [R, W] operand (W % 4 == 0).
__device__ __forceinline__ float4 load_rowmajor4(const float* base, int ld, int r, int c, int R, int W, const ConvView& cv) {
  if (r >= R || c >= W) return f4_zero();
  if (cv.taps > 0) {
    int j = c / cv.C, i = c - j * cv.C;
    int t = (r % cv.S) + j - cv.pad;
    if (t < 0 || t >= cv.S) return f4_zero();
    return ldg4(base + (size_t)(r + j - cv.pad) * ld + i);
  }
  return ldg4(base + (size_t)r * ld + c);
}

template <bool AK, bool BKM, bool CONV_A, bool CONV_B>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(GemmParams p) {
  pdl_prologue();
  __shared__ __align__(16) float As[2][BK][BM + SPAD];
  __shared__ __align__(16) float Bs[2][BK][BN + SPAD];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * p.kchunk;
  const int kend = min(p.K, kbeg + p.kchunk);
  const ConvView nocv{0, 0, 0, 0};
  const ConvView cva = CONV_A ? p.conv : nocv;
  const ConvView cvb = CONV_B ? p.conv : nocv;

  float4 ra[2], rb[2];

  auto load_tiles = [&](int k0) {
    if (AK) {   // A row-major [M,K]: thread -> (row = tid/4 + 64*i, kq = tid%4)
#pragma unroll
      for (int i = 0; i < 2; i++) {
        int r = (tid >> 2) + 64 * i, kq = tid & 3;
        int k = k0 + kq * 4;
        ra[i] = (k < kend) ? load_rowmajor4(p.A, p.lda, m0 + r, k, p.M, p.K, cva) : f4_zero();
      }
    } else {    // A stored [K,M] (m contiguous): thread -> (k = tid/32 + 8*i, mq = tid%32)
#pragma unroll
      for (int i = 0; i < 2; i++) {
        int kk = (tid >> 5) + 8 * i, mq = tid & 31;
        int k = k0 + kk;
        ra[i] = (k < kend) ? load_rowmajor4(p.A, p.lda, k, m0 + mq * 4, p.K, p.M, nocv) : f4_zero();
      }
    }
    if (BKM) {  // B stored [N,K] (k contiguous)
#pragma unroll
      for (int i = 0; i < 2; i++) {
        int r = (tid >> 2) + 64 * i, kq = tid & 3;
        int k = k0 + kq * 4;
        rb[i] = (k < kend) ? load_rowmajor4(p.B, p.ldb, n0 + r, k, p.N, p.K, nocv) : f4_zero();
      }
    } else {    // B stored [K,N] (n contiguous)
#pragma unroll
      for (int i = 0; i < 2; i++) {
        int kk = (tid >> 5) + 8 * i, nq = tid & 31;
        int k = k0 + kk;
        rb[i] = (k < kend) ? load_rowmajor4(p.B, p.ldb, k, n0 + nq * 4, p.K, p.N, cvb) : f4_zero();
      }
    }
  };
  auto store_tiles = [&](int buf) {
    if (AK) {
#pragma unroll
      for (int i = 0; i < 2; i++) {
        int r = (tid >> 2) + 64 * i, kq = (tid & 3) * 4;
        As[buf][kq + 0][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y;
        As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 2; i++) {
        int kk = (tid >> 5) + 8 * i, mq = (tid & 31) * 4;
        st4(&As[buf][kk][mq], ra[i]);
      }
    }
    if (BKM) {
#pragma unroll
      for (int i = 0; i < 2; i++) {
        int r = (tid >> 2) + 64 * i, kq = (tid & 3) * 4;
        Bs[buf][kq + 0][r] = rb[i].x; Bs[buf][kq + 1][r] = rb[i].y;
        Bs[buf][kq + 2][r] = rb[i].z; Bs[buf][kq + 3][r] = rb[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 2; i++) {
        int kk = (tid >> 5) + 8 * i, nq = (tid & 31) * 4;
        st4(&Bs[buf][kk][nq], rb[i]);
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  const int ty = tid >> 4, tx = tid & 15;   // 16 x 16 threads; each owns rows {ty*4..+3, 64+ty*4..+3}, cols likewise

  int nk = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;
  if (nk > 0) {
    load_tiles(kbeg);
    store_tiles(0);
  }
  __syncthreads();
  for (int it = 0; it < nk; it++) {
    int buf = it & 1;
    if (it + 1 < nk) load_tiles(kbeg + (it + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (it + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // epilogue
  const bool split = p.partial != nullptr;
  float* out = split ? p.partial + (size_t)blockIdx.z * p.M * p.N : p.C;
  const int ldo = split ? p.N : p.ldc;
#pragma unroll
  for (int i = 0; i < TM; i++) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    float rm = 1.f;
    if (!split && p.rowmask) rm = p.rowmask[m] > 0 ? 1.f : 0.f;
#pragma unroll
    for (int jh = 0; jh < 2; jh++) {
      int n = n0 + jh * 64 + tx * 4;
      if (n >= p.N) continue;   // N % 4 == 0
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      if (!split) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          float x = v[q];
          if (p.bias) x += __ldg(p.bias + n + q);
          if (p.act == 1) x = tanhf(x);
          else if (p.act == 2) x = fmaxf(x, 0.f);
          if (p.drop_p > 0.f) x *= dropout_scale(p.seed, (uint64_t)m * p.N + n + q, p.drop_p, 1.f / (1.f - p.drop_p));
          v[q] = x * rm;
        }
        if (p.accumulate) {
          float4 o = *reinterpret_cast<const float4*>(out + (size_t)m * ldo + n);
          v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
        }
      }
      st4(out + (size_t)m * ldo + n, make_float4(v[0], v[1], v[2], v[3]));
    }
  }
}

// Deterministic split-K reduction: C[m,n] (+)= sum_z partial[z,m,n].  A block owns 64 float4 outputs; its four z-lanes each sum
// every fourth partial (loads in flight in parallel) and are combined in a fixed order, so the result does not depend on timing.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, float* __restrict__ C, int M, int N, int ldc,
                                                            int splits, int accumulate) {
  pdl_prologue();
  __shared__ float4 red[3][64];
  const int q = threadIdx.x & 63, zl = threadIdx.x >> 6;
  const size_t i4 = (size_t)blockIdx.x * 64 + q;
  const size_t total4 = (size_t)M * N / 4, MN = (size_t)M * N;
  const bool ok = i4 < total4;
  const size_t e = i4 * 4;
  float4 s = f4_zero();
  if (ok) {
    int z = zl;
#pragma unroll 1
    for (; z + 12 < splits; z += 16) {
      const float4 a = ldg4_stream(partial + (size_t)z * MN + e), b = ldg4_stream(partial + (size_t)(z + 4) * MN + e);
      const float4 c = ldg4_stream(partial + (size_t)(z + 8) * MN + e), d = ldg4_stream(partial + (size_t)(z + 12) * MN + e);
      f4_add(s, a); f4_add(s, b); f4_add(s, c); f4_add(s, d);
    }
    for (; z < splits; z += 4) f4_add(s, ldg4_stream(partial + (size_t)z * MN + e));
  }
  if (zl > 0) red[zl - 1][q] = s;
  __syncthreads();
  if (zl == 0 && ok) {
    f4_add(s, red[0][q]); f4_add(s, red[1][q]); f4_add(s, red[2][q]);
    const int m = (int)(e / N), n = (int)(e % N);
    float* o = C + (size_t)m * ldc + n;
    if (accumulate) f4_add(s, *reinterpret_cast<const float4*>(o));
    st4(o, s);
  }
}

// Column sums of a row-major [M,N] matrix, deterministic two-stage (db = sum_m dY[m,:]).
// block = 32 columns x 8 row-lanes over `rows_per_block` rows; the row-lanes are combined in a fixed order (deterministic)
__global__ void __launch_bounds__(256) colsum_stage1(const float* __restrict__ X, float* __restrict__ part, int M, int N, int rows_per_block) {
  pdl_prologue();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (n < N) {
#pragma unroll 4
    for (int r = r0 + ty; r < r1; r += 8) s += __ldg(X + (size_t)r * N + n);
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += red[i][tx];
    part[(size_t)blockIdx.y * N + n] = t;
  }
}
__global__ void colsum_stage2(const float* __restrict__ part, float* __restrict__ out, int nparts, int N, int accumulate) {
  pdl_prologue();
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int i = 0; i < nparts; i++) s += part[(size_t)i * N + n];
  out[n] = accumulate ? out[n] + s : s;
}

// dPre = dY * act'(Y) * dropout_scale * rowmask   (Y is the saved epilogue OUTPUT, i.e. after act/dropout/mask)
__global__ void act_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ Y, const int64_t* __restrict__ rowmask,
                               float* __restrict__ dPre, int64_t M, int N, int act, float drop_p, unsigned long long seed) {
  pdl_prologue();
  int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int N4 = N >> 2;
  if (i4 >= M * N4) return;
  int64_t m = i4 / N4;
  int n = (int)(i4 % N4) * 4;
  float rm = rowmask ? (rowmask[m] > 0 ? 1.f : 0.f) : 1.f;
  float4 g = *reinterpret_cast<const float4*>(dY + m * N + n);
  float ge[4] = {g.x, g.y, g.z, g.w};
  float ye[4] = {0.f, 0.f, 0.f, 0.f};
  if (act != 0) {
    float4 y = *reinterpret_cast<const float4*>(Y + m * N + n);
    ye[0] = y.x; ye[1] = y.y; ye[2] = y.z; ye[3] = y.w;
  }
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    float sc = rm;
    if (drop_p > 0.f) sc *= dropout_scale(seed, (uint64_t)m * N + n + q, drop_p, inv_keep);
    float d;
    if (act == 1) {
      // Y = tanh(pre) * sc  ->  tanh(pre) = Y / sc where sc != 0 (only used with sc in {0,1} when drop_p == 0)
      float t = sc != 0.f ? ye[q] / sc : 0.f;
      d = 1.f - t * t;
    } else if (act == 2) {
      d = ye[q] > 0.f ? 1.f : 0.f;   // relu output > 0 iff pre > 0 and kept
    } else {
      d = 1.f;
    }
    ge[q] = ge[q] * d * sc;
  }
  st4(dPre + m * N + n, make_float4(ge[0], ge[1], ge[2], ge[3]));
}

__global__ void valid_mask_kernel(const int64_t* __restrict__ ids, int64_t* __restrict__ out, int64_t n) {
  pdl_prologue();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = ids[i] > -1 ? 1 : 0;
}

template <bool AK, bool BKM, bool CA, bool CB>
static int launch(const GemmParams& p, int splits, cudaStream_t st) {
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, splits);
  LK_LAUNCH((gemm_simt_kernel<AK, BKM, CA, CB>), grid, NT, 0, st, p);
  return check_launch("gemm_simt");
}

static int choose_splits(int M, int N, int K) {
  long tiles = (long)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  if (tiles >= 2 * kNumSMs || K <= 4 * BK * 8) return 1;
  long want = (2L * kNumSMs + tiles - 1) / tiles;
  long maxs = (K + 8 * BK - 1) / (8 * BK);
  long s = want < maxs ? want : maxs;
  return (int)(s < 1 ? 1 : (s > 512 ? 512 : s));
}

}  // namespace lk

using namespace lk;

extern "C" {

size_t lk_linear_bwd_weight_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  // grad-weight GEMM is [N,K] output reduced over M
  int splits = choose_splits((int)N, (int)K, (int)M);
  size_t a = (size_t)splits * N * K * sizeof(float);
  size_t rows_per_block = 512;
  size_t b = ((size_t)(M + rows_per_block - 1) / rows_per_block) * N * sizeof(float);
  return a + b + 256;
}

int lk_linear_fwd(const float* X, const float* W, const float* bias, const int64_t* rowmask, float* Y,
                  int64_t M, int64_t N, int64_t K, int act, int accumulate, float drop_p, uint64_t seed, cudaStream_t st) {
  LK_REQUIRE(K % 4 == 0 && N % 4 == 0, LK_ERR_SHAPE, "lk_linear_fwd: K=%ld and N=%ld must be multiples of 4", (long)K, (long)N);
  if (M == 0) return LK_OK;
  GemmParams p{X, W, Y, (int)M, (int)N, (int)K, (int)K, (int)K, (int)N, bias, rowmask, act, accumulate, (int)K + BK, nullptr, {0, 0, 0, 0}, drop_p, (unsigned long long)seed};
  return launch<true, true, false, false>(p, 1, st);
}

int lk_linear_bwd_data(const float* dY, const float* W, float* dX, int64_t M, int64_t N, int64_t K, int accumulate,
                       cudaStream_t st) {
  LK_REQUIRE(K % 4 == 0 && N % 4 == 0, LK_ERR_SHAPE, "lk_linear_bwd_data: K=%ld and N=%ld must be multiples of 4", (long)K, (long)N);
  if (M == 0) return LK_OK;
  // dX[M,K] = dY[M,N] · W[N,K]: contraction over N; B element (k'=n, n'=k) = W[n*K + k] -> n'-contiguous
  GemmParams p{dY, W, dX, (int)M, (int)K, (int)N, (int)N, (int)K, (int)K, nullptr, nullptr, 0, accumulate, (int)N + BK, nullptr, {0, 0, 0, 0}, 0.f, 0ULL};
  return launch<true, false, false, false>(p, 1, st);
}

int lk_linear_bwd_weight(const float* dY, const float* X, float* dW, float* db, int64_t M, int64_t N, int64_t K,
                         int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  LK_REQUIRE(K % 4 == 0 && N % 4 == 0, LK_ERR_SHAPE, "lk_linear_bwd_weight: K=%ld and N=%ld must be multiples of 4", (long)K, (long)N);
  LK_REQUIRE(workspace_bytes >= lk_linear_bwd_weight_workspace_bytes(M, N, K), LK_ERR_ARG, "lk_linear_bwd_weight: workspace too small");
  if (M == 0) {
    if (!accumulate) {
      cudaMemsetAsync(dW, 0, (size_t)N * K * sizeof(float), st);
      if (db) cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st);
    }
    return LK_OK;
  }
  int splits = choose_splits((int)N, (int)K, (int)M);
  float* ws = (float*)workspace;
  // dW[N,K] = dY^T[N,M] · X[M,K]: A element (m'=n, k'=m) = dY[m*N + n] (m'-contiguous), B (k'=m, n'=k) = X[m*K+k]
  int kchunk = (int)(((M + splits - 1) / splits + BK - 1) / BK * BK);
  GemmParams p{dY, X, dW, (int)N, (int)K, (int)M, (int)N, (int)K, (int)K, nullptr, nullptr, 0, accumulate, kchunk,
               splits > 1 ? ws : nullptr, {0, 0, 0, 0}, 0.f, 0ULL};
  int rc = launch<false, false, false, false>(p, splits, st);
  if (rc) return rc;
  if (splits > 1) {
    size_t total4 = (size_t)N * K / 4;
    LK_LAUNCH((splitk_reduce_kernel), (unsigned)((total4 + 63) / 64), 256, 0, st, ws, dW, (int)N, (int)K, (int)K, splits, accumulate);
    rc = check_launch("splitk_reduce");
    if (rc) return rc;
  }
  if (db) {
    float* part = ws + (size_t)splits * N * K;
    int rpb = 512;
    int nparts = (int)((M + rpb - 1) / rpb);
    dim3 g1((unsigned)((N + 31) / 32), nparts);
    LK_LAUNCH((colsum_stage1), g1, 256, 0, st, dY, part, (int)M, (int)N, rpb);
    LK_LAUNCH((colsum_stage2), (unsigned)((N + 127) / 128), 128, 0, st, part, db, nparts, (int)N, accumulate);
    rc = check_launch("colsum", 2);
  }
  return rc;
}

int lk_act_bwd(const float* dY, const float* Y, const int64_t* rowmask, float* dPre, int64_t M, int64_t N, int act, float drop_p,
               uint64_t seed, cudaStream_t st) {
  LK_REQUIRE(N % 4 == 0, LK_ERR_SHAPE, "lk_act_bwd: N=%ld must be a multiple of 4", (long)N);
  if (M == 0) return LK_OK;
  int64_t total = M * (N / 4);
  LK_LAUNCH((act_bwd_kernel), (unsigned)((total + 255) / 256), 256, 0, st, dY, Y, rowmask, dPre, M, (int)N, act, drop_p, (unsigned long long)seed);
  return check_launch("act_bwd");
}

int lk_valid_mask(const int64_t* ids, int64_t* out, int64_t n, cudaStream_t st) {
  if (n == 0) return LK_OK;
  LK_LAUNCH((valid_mask_kernel), (unsigned)((n + 255) / 256), 256, 0, st, ids, out, n);
  return check_launch("valid_mask");
}

int lk_splitk_reduce(const float* partial, float* C, int64_t M, int64_t N, int64_t ldc, int splits, int accumulate, cudaStream_t st) {
  LK_REQUIRE(N % 4 == 0 && ldc % 4 == 0, LK_ERR_SHAPE, "lk_splitk_reduce: N and ldc must be multiples of 4");
  size_t total4 = (size_t)M * N / 4;
  if (total4 == 0) return LK_OK;
  LK_LAUNCH((splitk_reduce_kernel), (unsigned)((total4 + 63) / 64), 256, 0, st, partial, C, (int)M, (int)N, (int)ldc, splits, accumulate);
  return check_launch("splitk_reduce");
}

size_t lk_colsum_workspace_bytes(int64_t M, int64_t N) { return ((size_t)(M + 511) / 512) * N * sizeof(float) + 256; }

int lk_colsum(const float* X, float* out, int64_t M, int64_t N, int accumulate, void* workspace, size_t workspace_bytes,
              cudaStream_t st) {
  LK_REQUIRE(workspace_bytes >= lk_colsum_workspace_bytes(M, N), LK_ERR_ARG, "lk_colsum: workspace too small");
  if (M == 0) {
    if (!accumulate) cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), st);
    return LK_OK;
  }
  int rpb = 512;
  int nparts = (int)((M + rpb - 1) / rpb);
  dim3 g1((unsigned)((N + 31) / 32), nparts);
  LK_LAUNCH((colsum_stage1), g1, 256, 0, st, X, (float*)workspace, (int)M, (int)N, rpb);
  LK_LAUNCH((colsum_stage2), (unsigned)((N + 127) / 128), 128, 0, st, (const float*)workspace, out, nparts, (int)N, accumulate);
  return check_launch("colsum", 2);
}

// ---- Conv1d(k, 'same') over S-long sequences as implicit-im2col GEMMs --------------------------------
// Wr[o, j*Cin + i] = W[o,i,j] (forward / grad-weight layout); Wd[i, j*Cout + o] = W[o,i,taps-1-j] (grad-input layout).
int lk_conv1d_fwd(const float* X, const float* Wr, const float* bias, const int64_t* rowmask, float* Y, int64_t rows,
                  int64_t S, int64_t Cin, int64_t Cout, int taps, int act, float drop_p, uint64_t seed, cudaStream_t st) {
  LK_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0 && taps % 2 == 1 && rows % S == 0, LK_ERR_SHAPE, "lk_conv1d_fwd: bad shape");
  if (rows == 0) return LK_OK;
  int K = (int)(taps * Cin);
  GemmParams p{X, Wr, Y, (int)rows, (int)Cout, K, (int)Cin, K, (int)Cout, bias, rowmask, act, 0, K + BK, nullptr,
               {(int)S, (int)Cin, taps, (taps - 1) / 2}, drop_p, (unsigned long long)seed};
  return launch<true, true, true, false>(p, 1, st);
}

int lk_conv1d_bwd_data(const float* dY, const float* Wd, float* dX, int64_t rows, int64_t S, int64_t Cin, int64_t Cout,
                       int taps, int accumulate, cudaStream_t st) {
  LK_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0 && taps % 2 == 1 && rows % S == 0, LK_ERR_SHAPE, "lk_conv1d_bwd_data: bad shape");
  if (rows == 0) return LK_OK;
  int K = (int)(taps * Cout);
  GemmParams p{dY, Wd, dX, (int)rows, (int)Cin, K, (int)Cout, K, (int)Cin, nullptr, nullptr, 0, accumulate, K + BK, nullptr,
               {(int)S, (int)Cout, taps, (taps - 1) / 2}, 0.f, 0ULL};
  return launch<true, true, true, false>(p, 1, st);
}

size_t lk_conv1d_bwd_weight_workspace_bytes(int64_t rows, int64_t Cin, int64_t Cout, int taps) {
  return lk_linear_bwd_weight_workspace_bytes(rows, Cout, (int64_t)taps * Cin);
}

int lk_conv1d_bwd_weight(const float* dY, const float* X, float* dWr, float* db, int64_t rows, int64_t S, int64_t Cin,
                         int64_t Cout, int taps, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  LK_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0 && taps % 2 == 1 && rows % S == 0, LK_ERR_SHAPE, "lk_conv1d_bwd_weight: bad shape");
  LK_REQUIRE(workspace_bytes >= lk_conv1d_bwd_weight_workspace_bytes(rows, Cin, Cout, taps), LK_ERR_ARG,
             "lk_conv1d_bwd_weight: workspace too small");
  int Kw = (int)(taps * Cin);
  if (rows == 0) {
    if (!accumulate) {
      cudaMemsetAsync(dWr, 0, (size_t)Cout * Kw * sizeof(float), st);
      if (db) cudaMemsetAsync(db, 0, (size_t)Cout * sizeof(float), st);
    }
    return LK_OK;
  }
  int splits = choose_splits((int)Cout, Kw, (int)rows);
  float* ws = (float*)workspace;
  int kchunk = (int)(((rows + splits - 1) / splits + BK - 1) / BK * BK);
  // dWr[Cout, taps*Cin] = dY^T · Xview ; B (k'=row, n'=kk) = Xview[row, kk] (n'-contiguous, conv view)
  GemmParams p{dY, X, dWr, (int)Cout, Kw, (int)rows, (int)Cout, (int)Cin, Kw, nullptr, nullptr, 0, accumulate, kchunk,
               splits > 1 ? ws : nullptr, {(int)S, (int)Cin, taps, (taps - 1) / 2}, 0.f, 0ULL};
  int rc = launch<false, false, false, true>(p, splits, st);
  if (rc) return rc;
  if (splits > 1) {
    size_t total4 = (size_t)Cout * Kw / 4;
    LK_LAUNCH((splitk_reduce_kernel), (unsigned)((total4 + 63) / 64), 256, 0, st, ws, dWr, (int)Cout, Kw, Kw, splits, accumulate);
    rc = check_launch("splitk_reduce");
    if (rc) return rc;
  }
  if (db) {
    float* part = ws + (size_t)splits * Cout * Kw;
    int rpb = 512;
    int nparts = (int)((rows + rpb - 1) / rpb);
    dim3 g1((unsigned)((Cout + 31) / 32), nparts);
    LK_LAUNCH((colsum_stage1), g1, 256, 0, st, dY, part, (int)rows, (int)Cout, rpb);
    LK_LAUNCH((colsum_stage2), (unsigned)((Cout + 127) / 128), 128, 0, st, part, db, nparts, (int)Cout, accumulate);
    rc = check_launch("colsum", 2);
  }
  return rc;
}

}  // extern "C"
