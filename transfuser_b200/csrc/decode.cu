// CenterNet box decode for the inference path: LidarCenterNetHead.get_bboxes / decode_heatmap (model.py:376-497) with
// mmdet 2.25's get_local_maximum (3x3 max-pool == self), get_topk_from_heatmap (top-k over the flattened map) and
// transpose_and_gather_feat folded into ONE launch that reads the raw [B,H,W,9+nb] head output once:
//   sigmoid(heat logit) -> 3x3 peak test -> per-sample top-k (bitonic sort of (score, cell) in shared memory)
//   -> gather the k cells' regressions -> argmax yaw bin / brake, class2angle (model.py:269-283), x4 to LiDAR-BEV pixels.
// One CTA per sample (the map is 64x64 = 4096 cells: 32 KB of keys+indices); the reference does this with ~25 ATen
// launches and a host sync. Ordering: descending score, ties broken by ascending cell index (torch.topk leaves ties
// unspecified); non-peak cells carry score 0 exactly as heat * (hmax == heat) does.
#include "common.cuh"

namespace {

constexpr int kMaxCells = 4096;
constexpr int kDecodeThreads = 1024;

__device__ __forceinline__ bool ranks_before(float sa, int ia, float sb, int ib) { return sa > sb || (sa == sb && ia < ib); }

// preds channel layout: 0 heat logit | 1-2 wh | 3-4 offset | 5..5+nb-1 yaw class | 5+nb yaw res | 6+nb velocity | 7+nb,8+nb brake
__global__ void __launch_bounds__(kDecodeThreads) centernet_decode_kernel(const float* __restrict__ preds, int H, int W, int nb, int k, int npad,
                                                                           float ratio, float angle_per_class, float* __restrict__ boxes,
                                                                           int* __restrict__ labels) {
  __shared__ float key[kMaxCells];
  __shared__ int cell[kMaxCells];
  const int b = blockIdx.x, HW = H * W, C = 9 + nb;
  const float* p = preds + (size_t)b * HW * C;
  for (int i = threadIdx.x; i < npad; i += blockDim.x) key[i] = (i < HW) ? 1.f / (1.f + expf(-p[(size_t)i * C])) : -1.f;
  __syncthreads();
  // peak test against the 8 neighbours (max_pool2d pads with -inf, i.e. out-of-map neighbours never win)
  float kept[kMaxCells / kDecodeThreads];
#pragma unroll
  for (int r = 0; r < kMaxCells / kDecodeThreads; ++r) {
    const int i = threadIdx.x + r * kDecodeThreads;
    float s = -1.f;
    if (i < HW) {
      s = key[i];
      const int y = i / W, x = i - y * W;
      bool peak = true;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int yy = y + dy, xx = x + dx;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W && key[yy * W + xx] > s) peak = false;
        }
      if (!peak) s = 0.f;
    }
    kept[r] = s;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kMaxCells / kDecodeThreads; ++r) {
    const int i = threadIdx.x + r * kDecodeThreads;
    if (i < npad) { key[i] = kept[r]; cell[i] = i; }
  }
  __syncthreads();
  // bitonic sort of npad (power of two) entries into ranks_before order
  for (int size = 2; size <= npad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (npad >> 1); t += blockDim.x) {
        const int i = 2 * t - (t & (stride - 1)), j = i + stride;
        const bool forward = (i & size) == 0;
        const float si = key[i], sj = key[j];
        const int ci = cell[i], cj = cell[j];
        if (ranks_before(sj, cj, si, ci) == forward) { key[i] = sj; key[j] = si; cell[i] = cj; cell[j] = ci; }
      }
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < k; t += blockDim.x) {
    const int c = cell[t];
    const float* q = p + (size_t)c * C;
    const int y = c / W, x = c - y * W;
    int best = 0;
    float bv = q[5];
    for (int a = 1; a < nb; ++a) if (q[5 + a] > bv) { bv = q[5 + a]; best = a; }
    float yaw = __fadd_rn(__fmul_rn((float)best, angle_per_class), q[5 + nb]);
    if (yaw > 3.14159274101257324f) yaw = __fadd_rn(yaw, -6.28318548202514648f);
    float* o = boxes + ((size_t)b * k + t) * 8;
    o[0] = ((float)x + q[3]) * ratio;
    o[1] = ((float)y + q[4]) * ratio;
    o[2] = q[1] * ratio;
    o[3] = q[2] * ratio;
    o[4] = yaw;
    o[5] = q[6 + nb];
    o[6] = q[8 + nb] > q[7 + nb] ? 1.f : 0.f;
    o[7] = key[t];
    labels[(size_t)b * k + t] = 0;  // single class (model.py:597: num_classes 1) -> flat index / (H*W) == 0
  }
}

}  // namespace

// preds: raw head output [B,H,W,9+num_dir_bins] (NHWC, heat as a logit). boxes: [B,k,8] = (x, y, w, h, yaw, velocity,
// brake class, score), labels: [B,k] int32. Requires H*W <= 4096 and k <= H*W.
TFB_API int tfb_centernet_decode(const float* preds, int B, int H, int W, int num_dir_bins, int k, float ratio, float* boxes, int* labels,
                                 cudaStream_t stream) {
  TFB_REQUIRE(preds && boxes && labels && B > 0 && H > 0 && W > 0 && num_dir_bins > 0);
  TFB_REQUIRE((int64_t)H * W <= kMaxCells && k > 0 && k <= H * W);
  int npad = 2;
  while (npad < H * W) npad <<= 1;
  const float apc = (float)(2.0 * 3.14159265358979323846 / (double)num_dir_bins);
  centernet_decode_kernel<<<B, kDecodeThreads, 0, stream>>>(preds, H, W, num_dir_bins, k, npad, ratio, apc, boxes, labels);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}
