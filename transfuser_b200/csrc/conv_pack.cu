// Weight packing for the tensor-core 3x3 convolution (csrc/conv_tc.cu): fp32 PyTorch-layout weights -> the bf16 B-operand layout
// the implicit-GEMM kernel reads, per conv (tfb_conv3x3_pack_weights) or for EVERY registered conv of the model in one launch
// (tfb_conv3x3_pack_weights_batched: once per training step, right after the optimizer changed the weights, instead of one
// launch per conv and direction inside forward and backward). CUDA-core code only: also built into the CPU emulation.
#include "common.cuh"

namespace {

// out[gb][chunk][tap][j][kk] (bf16): the B operand rows for output channel j of block gb and reduction channel kk.
// mode 0 (forward): value = w[co = gb*nb_real + j][ci - group(co)*Cig][tap],      ci = gb*c_step + chunk*KC + kk (same group only)
// mode 1 (dgrad)  : value = w[co = gb*c_step + chunk*KC + kk][ci' - group*Cig][8 - tap], ci' = gb*nb_real + j (conv input channel)
struct PackArgs {
  int Cout, Cin, groups, mode, NB, KC, c_step, nchunks, nb_real, gblocks;
};

__device__ __forceinline__ float pack_value(const float* __restrict__ w, int64_t i, const PackArgs& a) {
  const int Cig = a.Cin / a.groups, Cog = a.Cout / a.groups;
  const int kk = (int)(i % a.KC);
  const int j = (int)((i / a.KC) % a.NB);
  const int tap = (int)((i / ((int64_t)a.KC * a.NB)) % 9);
  const int chunk = (int)((i / ((int64_t)a.KC * a.NB * 9)) % a.nchunks);
  const int gb = (int)(i / ((int64_t)a.KC * a.NB * 9 * a.nchunks));
  float v = 0.f;
  const int oc = gb * a.nb_real + j;                  // channel of the tensor the conv kernel WRITES
  const int rc = gb * a.c_step + chunk * a.KC + kk;   // channel of the tensor the conv kernel READS
  if (j < a.nb_real) {
    if (a.mode == 0) {
      if (oc < a.Cout && rc < a.Cin) {
        const int g = oc / Cog;
        if (rc / Cig == g) v = w[((int64_t)oc * Cig + (rc - g * Cig)) * 9 + tap];
      }
    } else {
      if (oc < a.Cin && rc < a.Cout) {
        const int g = oc / Cig;
        if (rc / Cog == g) v = w[((int64_t)rc * Cig + (oc - g * Cig)) * 9 + (8 - tap)];
      }
    }
  }
  return v;
}

__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, PackArgs a) {
  const int64_t total = (int64_t)a.gblocks * a.nchunks * 9 * a.NB * a.KC;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(pack_value(w, i, a));
}

// One row of the device-resident descriptor table of the batched pack: 12 x int64
//   [0] weight pointer (fp32), [1] packed output pointer (bf16), [2..11] Cout, Cin, groups, mode, NB, KC, c_step, nchunks, nb_real, gblocks
constexpr int PACK_DESC_WORDS = 12;

__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const int64_t* __restrict__ table) {
  const int64_t* d = table + (int64_t)blockIdx.y * PACK_DESC_WORDS;
  const float* w = reinterpret_cast<const float*>(d[0]);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d[1]);
  PackArgs a = {(int)d[2], (int)d[3], (int)d[4], (int)d[5], (int)d[6], (int)d[7], (int)d[8], (int)d[9], (int)d[10], (int)d[11]};
  const int64_t total = (int64_t)a.gblocks * a.nchunks * 9 * a.NB * a.KC;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(pack_value(w, i, a));
}

// col[m][g][c][tap] (bf16) = x[n, ho*stride - 1 + tap/3, wo*stride - 1 + tap%3, g*Cg + c] (0 outside), m = (n, ho, wo): the im2col
// matrix whose column windows are the B operands of the per-group wgrad GEMMs (dW_g = dy_g^T col_g). Column order (c, tap) with
// the tap fastest = the order of PyTorch's weight layout [co][ci][kh][kw], so the GEMM writes dW in place (no permute pass).
// One thread = 8 consecutive channels x 9 taps of one (pixel, group): 18 16-byte loads, 144 contiguous bytes stored (Cg % 8 == 0).
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ col, int N, int H, int W, int C, int Ho, int Wo, int stride,
                 int groups) {
  const int Cg = C / groups, Cg8 = Cg / 8;
  const int64_t total = (int64_t)N * Ho * Wo * groups * Cg8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % Cg8);
    const int g = (int)((i / Cg8) % groups);
    const int64_t m = i / ((int64_t)Cg8 * groups);
    const int wo = (int)(m % Wo), ho = (int)((m / Wo) % Ho), n = (int)(m / ((int64_t)Wo * Ho));
    float v[9][8];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int hi = ho * stride - 1 + tap / 3, wi = wo * stride - 1 + tap % 3;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
        const float4* p = reinterpret_cast<const float4*>(x + (((int64_t)n * H + hi) * W + wi) * C + g * Cg + c8 * 8);
        a = p[0];
        b = p[1];
      }
      v[tap][0] = a.x; v[tap][1] = a.y; v[tap][2] = a.z; v[tap][3] = a.w;
      v[tap][4] = b.x; v[tap][5] = b.y; v[tap][6] = b.z; v[tap][7] = b.w;
    }
    // 72 outputs in (channel, tap) order, eight per 16-byte store
    uint4* dst = reinterpret_cast<uint4*>(col + m * 9 * (int64_t)C + ((int64_t)g * Cg + c8 * 8) * 9);
#pragma unroll
    for (int s = 0; s < 9; ++s) {
      uint32_t q[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int e0 = s * 8 + 2 * h, e1 = e0 + 1;
        __nv_bfloat162 pr = __floats2bfloat162_rn(v[e0 % 9][e0 / 9], v[e1 % 9][e1 / 9]);
        q[h] = *reinterpret_cast<uint32_t*>(&pr);
      }
      dst[s] = make_uint4(q[0], q[1], q[2], q[3]);
    }
  }
}

// y[m][0..Cp) (bf16) = x[m][0..C) zero-padded to Cp columns (narrow dy of the last decoder layers -> 16-byte rows for TMA)
__global__ void __launch_bounds__(256) cast_pad_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t M, int C, int Cp) {
  const int64_t total = M * Cp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const int64_t m = i / Cp;
    y[i] = __float2bfloat16_rn(c < C ? x[m * C + c] : 0.f);
  }
}

}  // namespace

// Packs fp32 PyTorch-layout 3x3 weights [Cout][Cin/groups][3][3] into the bf16 B-operand layout [gblocks][nchunks][9][NB][KC].
TFB_API int tfb_conv3x3_pack_weights(const float* w, void* out_bf16, int Cout, int Cin, int groups, int mode, int NB, int KC,
                                     int c_step, int nchunks, int nb_real, int gblocks, cudaStream_t stream) {
  TFB_REQUIRE(w && out_bf16 && Cout > 0 && Cin > 0 && groups > 0 && (mode == 0 || mode == 1) && NB > 0 && KC > 0);
  const int64_t total = (int64_t)gblocks * nchunks * 9 * NB * KC;
  PackArgs a = {Cout, Cin, groups, mode, NB, KC, c_step, nchunks, nb_real, gblocks};
  pack_weights_kernel<<<tfb_grid(total, 256), 256, 0, stream>>>(w, (__nv_bfloat16*)out_bf16, a);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}

// The same packing for n convs in ONE launch. table_dev: n rows of 12 int64 in device memory (see PACK_DESC_WORDS: the two
// pointers, then the ten arguments of tfb_conv3x3_pack_weights); blocks_x: CTAs per conv (grid-stride over its elements).
TFB_API int tfb_conv3x3_pack_weights_batched(const void* table_dev, int n, int blocks_x, cudaStream_t stream) {
  TFB_REQUIRE(table_dev && n > 0 && n <= 65535 && blocks_x > 0);
  pack_weights_batched_kernel<<<dim3((unsigned)blocks_x, (unsigned)n), 256, 0, stream>>>((const int64_t*)table_dev);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}

// im2col for the tensor-core wgrad: x fp32 NHWC -> col bf16 [N*Ho*Wo][groups][C/groups][9] (stride 1 or 2, pad 1; tap fastest).
TFB_API int tfb_im2col3x3_bf16(const float* x, void* col_bf16, int N, int H, int W, int C, int stride, int groups, cudaStream_t stream) {
  TFB_REQUIRE(x && col_bf16 && N > 0 && H > 0 && W > 0 && C > 0 && groups > 0 && (stride == 1 || stride == 2) && C % groups == 0 &&
              (C / groups) % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(col_bf16) & 15) == 0);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const int64_t total = (int64_t)N * Ho * Wo * (C / 8);
  im2col3x3_kernel<<<tfb_grid(total, 256, 16), 256, 0, stream>>>(x, (__nv_bfloat16*)col_bf16, N, H, W, C, Ho, Wo, stride, groups);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}

// Narrow dy of the last decoder layers (7 / 1 channels) -> zero-padded bf16 rows of Cp channels (16-byte rows for TMA).
TFB_API int tfb_cast_bf16_pad(const float* x, void* y_bf16, int64_t M, int C, int Cp, cudaStream_t stream) {
  TFB_REQUIRE(x && y_bf16 && M > 0 && C > 0 && Cp >= C);
  cast_pad_kernel<<<tfb_grid(M * Cp, 256), 256, 0, stream>>>(x, (__nv_bfloat16*)y_bf16, M, C, Cp);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}
