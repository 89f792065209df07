// BatchNorm2d (training mode, NHWC as a [M, C] matrix) and LayerNorm, forward + backward. HBM-bound.
// Replaces cuDNN/ATen batch_norm inside timm's BatchNormAct2d (every conv of the RegNetY trunks, transfuser.py:136-184)
// and nn.LayerNorm in Block / GPT.ln_f (transfuser.py:533-534, 321).
//   BN fwd : stats (per-channel sum, sum of squares; fp32 partials combined in fp64) -> normalise (+ReLU) + running stats
//   BN bwd : reduce (sum g, sum g*xhat with the ReLU mask recomputed from x) -> dx, dgamma, dbeta
#include "common.cuh"

namespace {

// ---------------- per-channel column reductions over a [M, C] matrix ----------------
// block (32, 8): threadIdx.x -> channel inside a 32-wide slab, threadIdx.y -> row phase. grid (slabs, row splits).
// MODE 0: (sum x, sum x^2)            MODE 1: BN backward (sum g, sum g*xhat), g = dy * relu_mask
template <int MODE>
__global__ void __launch_bounds__(256)
colreduce_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dy, int64_t M, int C, double* __restrict__ out,
                 const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int relu) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a0 = 0.f, a1 = 0.f;
  double d0 = 0.0, d1 = 0.0;
  if (c < C) {
    float mu = 0.f, is = 0.f, ga = 0.f, be = 0.f;
    if (MODE == 1) { mu = mean[c]; is = invstd[c]; ga = gamma[c]; be = beta[c]; }
    int cnt = 0;
    for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < M; r += (int64_t)gridDim.y * 8) {
      const float v = x[r * ldx + c];
      if (MODE == 0) {
        a0 += v; a1 = fmaf(v, v, a1);
      } else {
        const float xh = (v - mu) * is;
        float g = dy[r * C + c];
        if (relu && !(fmaf(xh, ga, be) > 0.f)) g = 0.f;
        a0 += g; a1 = fmaf(g, xh, a1);
      }
      if (++cnt == 64) { d0 += a0; d1 += a1; a0 = a1 = 0.f; cnt = 0; }  // bound fp32 partial length
    }
    d0 += a0; d1 += a1;
  }
  // combine the 8 row phases in fp64
  __shared__ double sd0[8][33], sd1[8][33];
  sd0[threadIdx.y][threadIdx.x] = d0;
  sd1[threadIdx.y][threadIdx.x] = d1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { t0 += sd0[j][threadIdx.x]; t1 += sd1[j][threadIdx.x]; }
    atomicAdd(&out[c], t0);
    atomicAdd(&out[C + c], t1);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, int64_t M, int C, float eps, float momentum,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / (double)M;
  double var = sums[C + c] / (double)M - mean * mean;
  if (var < 0.0) var = 0.0;
  save_mean[c] = (float)mean;
  save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// y = (x - mean) * invstd * gamma + beta (ReLU); 4 channels per thread (C % 4 == 0).
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total4, int C4, const float* __restrict__ mean,
                const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta, int relu) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 is = *reinterpret_cast<const float4*>(invstd + c);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
    const float4 be = *reinterpret_cast<const float4*>(beta + c);
    float4 o;
    o.x = fmaf((v.x - mu.x) * is.x, ga.x, be.x);
    o.y = fmaf((v.y - mu.y) * is.y, ga.y, be.y);
    o.z = fmaf((v.z - mu.z) * is.z, ga.z, be.z);
    o.w = fmaf((v.w - mu.w) * is.w, ga.w, be.w);
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    reinterpret_cast<float4*>(y)[i] = o;
  }
}

// dx = gamma * invstd * (g - sum_g/M - xhat * sum_gx/M);  block 0 also writes dgamma = sum_gx, dbeta = sum_g.
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int64_t M, int C,
                    const double* __restrict__ sums, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ gamma, const float* __restrict__ beta, int relu, float* __restrict__ dgamma,
                    float* __restrict__ dbeta) {
  const int64_t total = M * C;
  const double invM = 1.0 / (double)M;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float mu = mean[c], is = invstd[c], ga = gamma[c];
    const float xh = (x[i] - mu) * is;
    float g = dy[i];
    if (relu && !(fmaf(xh, ga, beta[c]) > 0.f)) g = 0.f;
    const float mg = (float)(sums[c] * invM), mgx = (float)(sums[C + c] * invM);
    dx[i] = ga * is * (g - mg - xh * mgx);
  }
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      dgamma[c] = (float)sums[C + c];
      dbeta[c] = (float)sums[c];
    }
  }
}

// ---------------- LayerNorm over the last dim of [R, C] ----------------
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int C, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const float* xr = x + r * C;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s += xr[c];
  const float mean = block_sum(s, red) / (float)C;
  float v = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { const float d = xr[c] - mean; v = fmaf(d, d, v); }
  const float var = block_sum(v, red) / (float)C;
  const float rstd = rsqrtf(var + eps);
  float* yr = y + r * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) yr[c] = fmaf((xr[c] - mean) * rstd, gamma[c], beta[c]);
  if (threadIdx.x == 0) { save_mean[r] = mean; save_rstd[r] = rstd; }
}

// dx = rstd * (g*gamma - mean_c(g*gamma) - xhat * mean_c(g*gamma*xhat))
__global__ void __launch_bounds__(256)
ln_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int C,
                 const float* __restrict__ gamma, const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                 int accumulate) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const float mean = save_mean[r], rstd = save_rstd[r];
  const float* xr = x + r * C;
  const float* gr = dy + r * C;
  float s1 = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float gg = gr[c] * gamma[c];
    s1 += gg;
    s2 = fmaf(gg, (xr[c] - mean) * rstd, s2);
  }
  const float m1 = block_sum(s1, red) / (float)C;
  const float m2 = block_sum(s2, red) / (float)C;
  float* dr = dx + r * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float xh = (xr[c] - mean) * rstd;
    const float v = rstd * (gr[c] * gamma[c] - m1 - xh * m2);
    dr[c] = accumulate ? dr[c] + v : v;
  }
}

// dgamma[c] = sum_r dy*xhat, dbeta[c] = sum_r dy  (column reduction; rows are few thousand)
__global__ void __launch_bounds__(256)
ln_bwd_param_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t R, int C, const float* __restrict__ save_mean,
                    const float* __restrict__ save_rstd, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float sg[8][33], sb[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f, b = 0.f;
  if (c < C) {
    for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < R; r += (int64_t)gridDim.y * 8) {
      const float g = dy[r * C + c];
      a = fmaf(g, (x[r * C + c] - save_mean[r]) * save_rstd[r], a);
      b += g;
    }
  }
  sg[threadIdx.y][threadIdx.x] = a;
  sb[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float ta = 0.f, tb = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { ta += sg[j][threadIdx.x]; tb += sb[j][threadIdx.x]; }
    atomicAdd(&dgamma[c], ta);
    atomicAdd(&dbeta[c], tb);
  }
}

__global__ void double_to_float_kernel(const double* __restrict__ in, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

int colreduce_splits(int64_t M, int slabs) {
  int64_t want = (4LL * tfb_num_sms() + slabs - 1) / slabs;
  int64_t maxs = ceil_div64(M, 8 * 16);  // at least 16 rows per thread
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  return (int)want;
}

}  // namespace

TFB_API int tfb_bn_fwd(const float* x, float* y, int64_t M, int C, const float* gamma, const float* beta, float eps,
                       float momentum, int relu, float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                       double* sums_ws, cudaStream_t stream) {
  TFB_REQUIRE(x && y && gamma && beta && save_mean && save_invstd && sums_ws && M > 0 && C > 0 && C % 4 == 0);
  if (cudaMemsetAsync(sums_ws, 0, 2 * (size_t)C * sizeof(double), stream) != cudaSuccess) return TFB_ERR_DRIVER;
  const int slabs = (C + 31) / 32;
  dim3 grid(slabs, colreduce_splits(M, slabs)), block(32, 8);
  colreduce_kernel<0><<<grid, block, 0, stream>>>(x, C, nullptr, M, C, sums_ws, nullptr, nullptr, nullptr, nullptr, 0);
  TFB_CHECK_LAUNCH();
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(sums_ws, M, C, eps, momentum, save_mean, save_invstd, running_mean, running_var);
  TFB_CHECK_LAUNCH();
  const int64_t total4 = M * C / 4;
  bn_apply_kernel<<<tfb_grid(total4, 256), 256, 0, stream>>>(x, y, total4, C / 4, save_mean, save_invstd, gamma, beta, relu);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}

// eval-mode BN: normalise with the running statistics (mean, 1/sqrt(var+eps) computed by the caller into save_*).
TFB_API int tfb_bn_apply(const float* x, float* y, int64_t M, int C, const float* gamma, const float* beta, const float* mean,
                         const float* invstd, int relu, cudaStream_t stream) {
  TFB_REQUIRE(x && y && gamma && beta && mean && invstd && M > 0 && C > 0 && C % 4 == 0);
  const int64_t total4 = M * C / 4;
  bn_apply_kernel<<<tfb_grid(total4, 256), 256, 0, stream>>>(x, y, total4, C / 4, mean, invstd, gamma, beta, relu);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}

TFB_API int tfb_bn_bwd(const float* x, const float* dy, float* dx, int64_t M, int C, const float* gamma, const float* beta,
                       const float* save_mean, const float* save_invstd, int relu, float* dgamma, float* dbeta, double* sums_ws,
                       cudaStream_t stream) {
  TFB_REQUIRE(x && dy && dx && gamma && beta && save_mean && save_invstd && dgamma && dbeta && sums_ws && M > 0 && C > 0);
  if (cudaMemsetAsync(sums_ws, 0, 2 * (size_t)C * sizeof(double), stream) != cudaSuccess) return TFB_ERR_DRIVER;
  const int slabs = (C + 31) / 32;
  dim3 grid(slabs, colreduce_splits(M, slabs)), block(32, 8);
  colreduce_kernel<1><<<grid, block, 0, stream>>>(x, C, dy, M, C, sums_ws, save_mean, save_invstd, gamma, beta, relu);
  TFB_CHECK_LAUNCH();
  bn_bwd_apply_kernel<<<tfb_grid(M * C, 256), 256, 0, stream>>>(x, dy, dx, M, C, sums_ws, save_mean, save_invstd, gamma, beta, relu,
                                                                dgamma, dbeta);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}

// out[c] = sum_r x[r][c]  (bias gradients of Linear / 1x1 conv layers)
// x is [M, C] with row stride ldx (elements).
TFB_API int tfb_colsum(const float* x, int64_t ldx, int64_t M, int C, float* out, double* sums_ws, cudaStream_t stream) {
  TFB_REQUIRE(x && out && sums_ws && M > 0 && C > 0 && ldx >= C);
  if (cudaMemsetAsync(sums_ws, 0, 2 * (size_t)C * sizeof(double), stream) != cudaSuccess) return TFB_ERR_DRIVER;
  const int slabs = (C + 31) / 32;
  dim3 grid(slabs, colreduce_splits(M, slabs)), block(32, 8);
  colreduce_kernel<0><<<grid, block, 0, stream>>>(x, ldx, nullptr, M, C, sums_ws, nullptr, nullptr, nullptr, nullptr, 0);
  TFB_CHECK_LAUNCH();
  double_to_float_kernel<<<(C + 127) / 128, 128, 0, stream>>>(sums_ws, out, C);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}

TFB_API int tfb_layernorm_fwd(const float* x, float* y, int64_t R, int C, const float* gamma, const float* beta, float eps,
                              float* save_mean, float* save_rstd, cudaStream_t stream) {
  TFB_REQUIRE(x && y && gamma && beta && save_mean && save_rstd && R > 0 && C > 0);
  ln_fwd_kernel<<<(unsigned)R, 256, 0, stream>>>(x, y, C, gamma, beta, eps, save_mean, save_rstd);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}

// dx (+)= LN backward; dgamma/dbeta are overwritten.
TFB_API int tfb_layernorm_bwd(const float* x, const float* dy, float* dx, int64_t R, int C, const float* gamma,
                              const float* save_mean, const float* save_rstd, float* dgamma, float* dbeta, int accumulate_dx,
                              cudaStream_t stream) {
  TFB_REQUIRE(x && dy && dx && gamma && save_mean && save_rstd && dgamma && dbeta && R > 0 && C > 0);
  ln_bwd_dx_kernel<<<(unsigned)R, 256, 0, stream>>>(x, dy, dx, C, gamma, save_mean, save_rstd, accumulate_dx);
  TFB_CHECK_LAUNCH();
  if (cudaMemsetAsync(dgamma, 0, (size_t)C * sizeof(float), stream) != cudaSuccess) return TFB_ERR_DRIVER;
  if (cudaMemsetAsync(dbeta, 0, (size_t)C * sizeof(float), stream) != cudaSuccess) return TFB_ERR_DRIVER;
  const int slabs = (C + 31) / 32;
  int splits = (int)ceil_div64(R, 8 * 8);
  if (splits > 64) splits = 64;
  if (splits < 1) splits = 1;
  dim3 grid(slabs, splits), block(32, 8);
  ln_bwd_param_kernel<<<grid, block, 0, stream>>>(x, dy, R, C, save_mean, save_rstd, dgamma, dbeta);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}
