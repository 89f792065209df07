/* transfuser_b200 — C-ABI of the B200-native (sm_100a) TransFuser training hot path.
 *
 * The reference (autonomousvision/transfuser) is pure Python and has no FFI layer: its "operator interface" for this
 * path is the set of torch.nn / torch.nn.functional calls made by transfuser.py and model.py. Each entry point below
 * cites the reference call site it replaces (paths relative to /root/reference/team_code_transfuser/).
 *
 * Conventions: plain device pointers + sizes, no torch types; every call is asynchronous on `stream`, never
 * allocates, never synchronises; returns 0 on success or a negative TFB_ERR_* code (tfb_last_error() has the text).
 * Activations are NHWC ("channels-last") fp32 unless stated; weights keep the reference's (PyTorch) layouts.
 */
#ifndef TFB200_H_
#define TFB200_H_
#include <stdint.h>

#ifdef __cplusplus
#define TFB_EXPORT extern "C" __attribute__((visibility("default")))
#else
#define TFB_EXPORT
#endif

#ifdef __CUDACC__
#include <cuda_runtime.h>
typedef cudaStream_t tfb_stream_t;
#else
typedef void* tfb_stream_t; /* a cudaStream_t */
#endif

#define TFB_OK 0
#define TFB_ERR_ARG (-1)
#define TFB_ERR_LAUNCH (-2)
#define TFB_ERR_UNSUPPORTED (-3)
#define TFB_ERR_DRIVER (-4)

TFB_EXPORT const char* tfb_last_error(void);
TFB_EXPORT int tfb_abi_version(void);

/* LiDAR points -> (2,256,256) BEV histogram, bit-exact with data.py:446-470 (`lidar_to_histogram_features`).
 * points: [batch][n_max][4] (x,y,z,intensity) fp32 or fp64 (is_f64); n_valid: optional [batch] int32 point counts;
 * counts_ws: workspace batch*2*256*256 uint32; out: [batch][2][256][256] fp32 in {0,.2,.4,.6,.8,1}. */
TFB_EXPORT int tfb_bev_histogram(const void* points, int is_f64, const int* n_valid, int batch, int n_max,
                                 unsigned int* counts_ws, float* out, tfb_stream_t stream);

/* C = alpha*op(A)*op(B) + beta*C (+bias[n]) (ReLU), row-major fp32 on CUDA cores, two-level strided batch.
 * op(A)[m][k] = transA ? A[k*lda+m] : A[m*lda+k];  op(B)[k][n] = transB ? B[n*ldb+k] : B[k*ldb+n].
 * Replaces torch.matmul / nn.Linear in SelfAttention.forward (transfuser.py:510-527) and small heads (model.py:592-605). */
TFB_EXPORT int tfb_gemm_f32_simt(int transA, int transB, int M, int N, int K, const float* A, int64_t lda, const float* B,
                                 int64_t ldb, float* C, int64_t ldc, const float* bias, int relu, float alpha, float beta,
                                 int batch_outer, int batch_inner, int64_t sAo, int64_t sAi, int64_t sBo, int64_t sBi,
                                 int64_t sCo, int64_t sCi, tfb_stream_t stream);

/* Same contract (no batching) on the 5th-gen tensor cores: TMA -> 128B-swizzled smem -> tcgen05.mma kind::tf32 with the
 * accumulator in TMEM. splits > 1 = split-K with fp32 atomics (wgrad). Needs 16-byte aligned bases / leading dimensions.
 * Replaces nn.Linear (transfuser.py:498-506,538-543) and the 1x1 nn.Conv2d of timm's RegNet blocks (transfuser.py:380,442). */
TFB_EXPORT int tfb_gemm_tf32_tc(int transA, int transB, int M, int N, int K, const float* A, int64_t lda, const float* B,
                                int64_t ldb, float* C, int64_t ldc, const float* bias, int relu, float alpha, float beta,
                                int splits, tfb_stream_t stream);
/* bf16 operands (kind::f16), fp32 accumulate/output. */
TFB_EXPORT int tfb_gemm_bf16_tc(int transA, int transB, int M, int N, int K, const void* A, int64_t lda, const void* B,
                                int64_t ldb, float* C, int64_t ldc, const float* bias, int relu, float alpha, float beta,
                                int splits, tfb_stream_t stream);

#endif /* TFB200_H_ */
