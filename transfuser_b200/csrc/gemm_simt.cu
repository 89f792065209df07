// fp32 CUDA-core GEMM (exact-fp32 path): C = alpha * op(A) * op(B) + beta * C (+ bias[n]) (ReLU optional),
// row-major, arbitrary leading dimensions, two-level strided batching (used for per-(sample, head) attention
// products and for every shape the tcgen05 path does not take: tiny M, odd leading dimensions, parity mode).
// Replaces the cuBLAS calls behind nn.Linear / 1x1 nn.Conv2d / torch.matmul in the reference
// (/root/reference/team_code_transfuser/transfuser.py:510-527, 538-543; model.py:592-605).
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(int M, int N, int K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                 float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int relu, float alpha, float beta,
                 int nbi, int64_t sAo, int64_t sAi, int64_t sBo, int64_t sBi, int64_t sCo, int64_t sCi) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int z = blockIdx.z, zo = z / nbi, zi = z % nbi;
  A += zo * sAo + zi * sAi;
  B += zo * sBo + zi * sBi;
  C += zo * sCo + zi * sCi;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // A tile -> As[k][m]
    if (!TA) {  // A[m*lda + k]: k contiguous
      for (int e = tid; e < BM * BK; e += 256) {
        int kk = e % BK, mm = e / BK;
        int m = m0 + mm, k = k0 + kk;
        As[kk][mm] = (m < M && k < K) ? A[(int64_t)m * lda + k] : 0.f;
      }
    } else {    // A[k*lda + m]: m contiguous
      for (int e = tid; e < BM * BK; e += 256) {
        int mm = e % BM, kk = e / BM;
        int m = m0 + mm, k = k0 + kk;
        As[kk][mm] = (m < M && k < K) ? A[(int64_t)k * lda + m] : 0.f;
      }
    }
    if (!TB) {  // B[k*ldb + n]: n contiguous
      for (int e = tid; e < BN * BK; e += 256) {
        int nn = e % BN, kk = e / BN;
        int n = n0 + nn, k = k0 + kk;
        Bs[kk][nn] = (n < N && k < K) ? B[(int64_t)k * ldb + n] : 0.f;
      }
    } else {    // B[n*ldb + k]: k contiguous
      for (int e = tid; e < BN * BK; e += 256) {
        int kk = e % BK, nn = e / BK;
        int n = n0 + nn, k = k0 + kk;
        Bs[kk][nn] = (n < N && k < K) ? B[(int64_t)n * ldb + k] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
      a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
      b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= N) continue;
      float v = alpha * acc[i][j];
      if (bias) v += bias[n];
      float* c = C + (int64_t)m * ldc + n;
      if (beta != 0.f) v += beta * (*c);
      if (relu) v = fmaxf(v, 0.f);
      *c = v;
    }
  }
}

}  // namespace

TFB_API int tfb_gemm_f32_simt(int transA, int transB, int M, int N, int K, const float* A, int64_t lda, const float* B,
                              int64_t ldb, float* C, int64_t ldc, const float* bias, int relu, float alpha, float beta,
                              int batch_outer, int batch_inner, int64_t sAo, int64_t sAi, int64_t sBo, int64_t sBi,
                              int64_t sCo, int64_t sCi, cudaStream_t stream) {
  TFB_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch_outer >= 1 && batch_inner >= 1);
  if (M == 0 || N == 0) return TFB_OK;
  TFB_REQUIRE(A && B && C);
  int64_t nz = (int64_t)batch_outer * batch_inner;
  TFB_REQUIRE(nz <= 65535);
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, (unsigned)nz);
  TFB_REQUIRE(grid.y <= 65535);
#define LAUNCH(TA, TB) gemm_simt_kernel<TA, TB><<<grid, 256, 0, stream>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, relu, alpha, beta, \
                                                                           batch_inner, sAo, sAi, sBo, sBi, sCo, sCi)
  if (!transA && !transB) LAUNCH(false, false);
  else if (!transA && transB) LAUNCH(false, true);
  else if (transA && !transB) LAUNCH(true, false);
  else LAUNCH(true, true);
#undef LAUNCH
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}
