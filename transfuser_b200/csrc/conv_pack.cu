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
