// Implicit-GEMM 3x3 convolution (stride 1, pad 1, NHWC) on the tcgen05 tensor cores — no im2col buffer.
// One CTA = one 8x16 spatial tile (128 output pixels = the UMMA M dimension) of one image x one block of output channels.
// For each of the 9 taps and each input-channel chunk, TMA loads the tap-SHIFTED activation patch straight from the NHWC
// bf16 tensor with a 4-D tensor map (box {KC channels, 16 w, 8 h, 1 n}; halo / image-border pixels are TMA out-of-bounds
// zero fill = the conv's zero padding) into 128B/64B-swizzled shared memory, where it is already a K-major A operand;
// the packed weights (K-major B operand, bf16, zero-padded, block-diagonal for grouped convs) come through a 2-D map.
// All taps / chunks accumulate into ONE fp32 TMEM accumulator; the epilogue adds bias / ReLU and writes fp32 NHWC.
// dgrad (stride 1) is the same kernel on dy with tap-flipped, transposed packed weights.
// Grouped RegNetY convs (group width 24): two groups share one CTA (N = 48 output channels, 64-channel K window).
// Replaces cuDNN behind the grouped 3x3 of timm's RegNetY blocks (transfuser.py:145-184) and the dense 3x3 of the
// decoders / heads (transfuser.py:214-281; model.py:93-99, 581-585). Warp roles as in gemm_tc.cu.
#include "common.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int TH = 8, TW = 16, BM = TH * TW;

__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t addr, int row_bytes) {
  // K-major operand, rows of `row_bytes` (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B); 8-row groups are row_bytes*8 apart.
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                                    // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(((uint32_t)row_bytes * 8) >> 4) << 32;     // SBO
  d |= (uint64_t)1 << 46;                                    // descriptor version (Blackwell)
  d |= (uint64_t)(row_bytes == 128 ? 2 : 4) << 61;           // SWIZZLE_128B / SWIZZLE_64B
  return d;
}

template <int KC, int NB, int STAGES>
__global__ void __launch_bounds__(192)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w, float* __restrict__ y,
                  const float* __restrict__ bias, int H, int W, int Cout, int tiles_w, int tiles_h, int c_step, int nchunks,
                  int nb_real, int relu, int gblocks, int total_tiles, int stride, double* __restrict__ stats) {
  // H, W: OUTPUT height / width. stride 2: the activation map steps two pixels per box element, so tile pixel (h, w) reads the
  // input pixel (2h + dy, 2w + dx) of tap (dy, dx).
  // PERSISTENT: tile t = (spatial tile, channel block), channel block fastest (CTAs running together re-use the same activation
  // patch out of L2); double-buffered TMEM accumulator so the epilogue of tile i overlaps the MMAs of tile i+1.
  constexpr int ROWB = KC * 2;                 // bytes per smem row
  constexpr int A_BYTES = BM * ROWB, B_BYTES = NB * ROWB, STAGE = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* stage_out = reinterpret_cast<float*>(smem + STAGES * STAGE + 256);
  float* s_stat = stage_out + 4 * 32 * 36;      // [2][Cout] per-channel statistics accumulators (present only when stats != nullptr)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int iters = nchunks * 9;
  constexpr uint32_t kAccCols = NB <= 32 ? 32 : NB <= 64 ? 64 : 128;
  constexpr uint32_t kTmemCols = 2 * kAccCols;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tma_x);
    tc::tma_prefetch_desc(&tma_w);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&tmem_full_bar[s], 1); tc::mbar_init(&tmem_empty_bar[s], 4); }
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int t, int& gb, int& n, int& h0, int& w0) {
    gb = t % gblocks;
    int sp = t / gblocks;
    const int tw = sp % tiles_w; sp /= tiles_w;
    const int th = sp % tiles_h; sp /= tiles_h;
    n = sp; h0 = th * TH; w0 = tw * TW;
  };

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int gb, n, h0, w0;
        decode(t, gb, n, h0, w0);
        for (int i = 0; i < iters; ++i, ++it) {
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          const int chunk = i / 9, tap = i % 9;
          tc::mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * STAGE;
          tc::mbar_expect_tx(&full_bar[s], STAGE);
          tc::tma_load_4d(&tma_x, &full_bar[s], sa, gb * c_step + chunk * KC, w0 * stride + tap % 3 - 1, h0 * stride + tap / 3 - 1, n);
          tc::tma_load_2d(&tma_w, &full_bar[s], sa + A_BYTES, 0, ((gb * nchunks + chunk) * 9 + tap) * NB);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc(1u, 0u, 0u, BM, NB);  // bf16 x bf16 -> fp32, both K-major
      int it = 0, lt = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
        const int as = lt & 1, aph = (lt >> 1) & 1;
        tc::mbar_wait(&tmem_empty_bar[as], aph ^ 1);
        tc::fence_after_sync();
        const uint32_t acc_addr = tmem_base + (uint32_t)as * kAccCols;
        for (int i = 0; i < iters; ++i, ++it) {
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          tc::mbar_wait(&full_bar[s], ph);
          tc::fence_after_sync();
          const uint32_t sa = tc::smem_u32(smem + s * STAGE), sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < KC / 16; ++k)
            tc::umma_f16(acc_addr, smem_desc_kmajor(sa + k * 32, ROWB), smem_desc_kmajor(sb + k * 32, ROWB), idesc, (i > 0 || k > 0) ? 1u : 0u);
          tc::umma_commit(&empty_bar[s]);
        }
        tc::umma_commit(&tmem_full_bar[as]);
      }
    }
  } else {
    const int q = warp & 3;
    float* stage = stage_out + q * (32 * 36);
    int lt = 0;
    if (stats) tc::stat_clear(s_stat, Cout, (int)threadIdx.x - 64);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
      int gb, n, h0, w0;
      decode(t, gb, n, h0, w0);
      const int as = lt & 1, aph = (lt >> 1) & 1;
      tc::mbar_wait(&tmem_full_bar[as], aph);
      tc::fence_after_sync();
      const uint32_t acc_addr = tmem_base + (uint32_t)as * kAccCols + ((uint32_t)(q * 32) << 16);
      const int nvalid = min(nb_real, Cout - gb * nb_real);
      float* ytile = y + (int64_t)gb * nb_real;
      const float* bp = bias ? bias + gb * nb_real : nullptr;
      const bool fast = (Cout & 3) == 0 && ((gb * nb_real) & 3) == 0 && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
      if (fast) {
        // accumulator row r of this warp's band = pixel (h0 + (q*32 + r) / 16, w0 + (q*32 + r) % 16)  (W fastest in the TMA box)
#pragma unroll 1
        for (int c0 = 0; c0 < NB; c0 += 32) {
          const int cols_valid = nvalid - c0;
          if (cols_valid <= 0) break;
          tc::epilogue_chunk32(acc_addr + (uint32_t)c0, stage,
                               [=](int row) -> int64_t {
                                 const int r = q * 32 + row, h = h0 + r / TW, w = w0 + r % TW;
                                 return (h < H && w < W) ? (((int64_t)n * H + h) * W + w) * Cout + c0 : (int64_t)-1;
                               },
                               cols_valid, bp ? bp + c0 : nullptr, 1.f, relu, lane, ytile, nullptr,
                               stats ? s_stat + gb * nb_real + c0 : nullptr, stats ? s_stat + Cout + gb * nb_real + c0 : nullptr);
        }
      } else {
        const int r = q * 32 + lane;
        const int h = h0 + r / TW, w = w0 + r % TW;
        const bool pix_ok = h < H && w < W;
        float* yp = ytile + (((int64_t)n * H + h) * W + w) * Cout;
#pragma unroll 1
        for (int c0 = 0; c0 < NB; c0 += 16) {
          float v[16];
          tc::tmem_ld16(acc_addr + (uint32_t)c0, v);
          if (pix_ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (c0 + j < nvalid) {
                float o = v[j] + (bp ? bp[c0 + j] : 0.f);
                if (relu) o = fmaxf(o, 0.f);
                yp[c0 + j] = o;
              }
            }
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty_bar[as]);
    }
    if (stats) tc::stat_flush(s_stat, stats, Cout, (int)threadIdx.x - 64);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int KC, int NB>
int launch_conv(const void* x16, const void* wpack, const float* bias, float* y, int N, int Hin, int Win, int Cx, int Cy, int c_step,
                int nchunks, int nb_real, int gblocks, int relu, int stride, double* stats, cudaStream_t stream) {
  const int H = (Hin - 1) / stride + 1, W = (Win - 1) / stride + 1;       // output size (kernel 3, pad 1)
  constexpr int STAGES = NB <= 64 ? 6 : 4;
  constexpr int STAGE = BM * KC * 2 + NB * KC * 2;
  constexpr int SMEM0 = STAGES * STAGE + 256 + 4 * 32 * 36 * 4 + 1024;   // ring + barriers/TMEM slot + epilogue staging + alignment slack
  const int SMEM = SMEM0 + (stats ? 2 * Cy * (int)sizeof(float) : 0);    // + per-channel statistics accumulators
  CUtensorMap mx, mw;
  const uint64_t xd[4] = {(uint64_t)Cx, (uint64_t)Win, (uint64_t)Hin, (uint64_t)N};
  const uint64_t xs[3] = {(uint64_t)Cx * 2, (uint64_t)Win * Cx * 2, (uint64_t)Hin * Win * Cx * 2};
  const uint32_t xb[4] = {(uint32_t)KC, (uint32_t)(TW * stride), (uint32_t)(TH * stride), 1};
  const uint32_t xe[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
  const CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  if (!tc::make_map(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x16, xd, xs, xb, swz, xe)) {
    tfb_set_last_error("cuTensorMapEncodeTiled(activation, 4-D) failed");
    return TFB_ERR_DRIVER;
  }
  const uint64_t wd[2] = {(uint64_t)KC, (uint64_t)gblocks * nchunks * 9 * NB};
  const uint64_t ws[1] = {(uint64_t)KC * 2};
  const uint32_t wb[2] = {(uint32_t)KC, (uint32_t)NB};
  if (!tc::make_map(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wpack, wd, ws, wb, swz)) {
    tfb_set_last_error("cuTensorMapEncodeTiled(packed weights) failed");
    return TFB_ERR_DRIVER;
  }
  auto kern = conv3x3_tc_kernel<KC, NB, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM0 + 2 * 2048 * (int)sizeof(float)) != cudaSuccess) return TFB_ERR_DRIVER;
    attr_done = true;
  }
  const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
  const int64_t total = (int64_t)tiles_w * tiles_h * N * gblocks;
  if (total > 0x7fffffff) { tfb_set_last_error("too many tiles"); return TFB_ERR_ARG; }
  // NB <= 64 tiles use < 100 KB of shared memory: two persistent CTAs per SM hide each other's pipeline bubbles
  const int per_sm = (SMEM <= 110 * 1024) ? 2 : 1;
  const int64_t cap = (int64_t)tfb_num_sms() * per_sm;
  const int grid = (int)(total < cap ? total : cap);
  kern<<<grid, 192, SMEM, stream>>>(mx, mw, y, bias, H, W, Cy, tiles_w, tiles_h, c_step, nchunks, nb_real, relu, gblocks, (int)total, stride, stats);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}

}  // namespace

// y[N,Ho,Wo,Cy] (fp32) = conv3x3(x16[N,H,W,Cx] bf16, packed weights; pad 1, stride 1 or 2) (+bias) (ReLU), Ho = (H-1)/stride + 1.
// Block gb reads channels [gb*c_step + chunk*KC, +KC) for chunk < nchunks and writes channels [gb*nb_real, +min(nb_real, Cy - gb*nb_real)).
// stride 2 (first block of every RegNetY stage): the TMA map traverses the activation with element strides {1, 2, 2, 1}.
// stats (optional, 2*Cy doubles, zeroed by the caller once per step): per-channel sum / sum of squares of y accumulated in the
// epilogue (training-mode BatchNorm statistics without a pass over y); needs Cy % 4 == 0, nb_real % 4 == 0, no bias / ReLU.
TFB_API int tfb_conv3x3_tc_strided(const void* x16, const void* wpack, const float* bias, float* y, int N, int H, int W, int Cx, int Cy,
                                   int NB, int KC, int c_step, int nchunks, int nb_real, int gblocks, int relu, int stride, double* stats,
                                   cudaStream_t stream) {
  TFB_REQUIRE(x16 && wpack && y && N > 0 && H > 0 && W > 0 && Cx > 0 && Cy > 0 && Cx % 8 == 0 && (stride == 1 || stride == 2));
  TFB_REQUIRE(!stats || (Cy % 4 == 0 && Cy <= 2048 && nb_real % 4 == 0 && !bias && !relu && (reinterpret_cast<uintptr_t>(y) & 15) == 0));
  TFB_REQUIRE((reinterpret_cast<uintptr_t>(x16) & 15) == 0 && (reinterpret_cast<uintptr_t>(wpack) & 15) == 0);
#define CASE(KC_, NB_) if (KC == KC_ && NB == NB_) return launch_conv<KC_, NB_>(x16, wpack, bias, y, N, H, W, Cx, Cy, c_step, nchunks, nb_real, gblocks, relu, stride, stats, stream)
  CASE(64, 16); CASE(64, 32); CASE(64, 48); CASE(64, 64); CASE(64, 128);
  CASE(32, 16); CASE(32, 32); CASE(32, 64);
#undef CASE
  tfb_set_last_error("tfb_conv3x3_tc: unsupported (KC, NB) tile");
  return TFB_ERR_UNSUPPORTED;
}

// Stride-1 form of tfb_conv3x3_tc_strided.
TFB_API int tfb_conv3x3_tc(const void* x16, const void* wpack, const float* bias, float* y, int N, int H, int W, int Cx, int Cy,
                           int NB, int KC, int c_step, int nchunks, int nb_real, int gblocks, int relu, cudaStream_t stream) {
  return tfb_conv3x3_tc_strided(x16, wpack, bias, y, N, H, W, Cx, Cy, NB, KC, c_step, nchunks, nb_real, gblocks, relu, 1, nullptr, stream);
}
