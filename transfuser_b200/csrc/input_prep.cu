// GPU side of the reference's per-sample data preparation (CARLA_Data.__getitem__, data.py:103-356), fed with COMPACT
// inputs (uint8 camera frames, the target point, raw LiDAR points + one 4x4 pose transform) instead of the expanded fp32 /
// int64 tensors the reference ships from its CPU workers. SURVEY.md §8f rank 2. What each kernel replaces:
//   tfb_draw_target_point  data.py:616-630  draw_target_point: cv2.circle(radius 5, thickness 3) on a 256x256 map
//   tfb_camera_prep        data.py:225-238, 358-372, 536-551, 561-576: crop_image_cv2 / crop_seg with the augmentation
//                          shift, HWC->CHW, get_depth, the class converter LUT, and (optionally) normalize_imagenet + NHWC
//                          (transfuser.py:419-428) so the backbone's own prep pass disappears
// (align + histogram fused: tfb_bev_histogram_aligned in bev_hist.cu.)
// All integer / byte work is exact; the depth decode follows the reference's float64 arithmetic and rounds to fp32 once.
#include "common.cuh"

namespace {

constexpr int kMap = 256;
// cv2.circle(img, c, radius=5, thickness=3) rasterised by OpenCV 4.13 around c: 15 rows, bit j = column c.x - 7 + j.
// Generated and checked for translation invariance / border clipping over all 257x257 clipped centres by
// oracle/make_golden.py (target stamp); the same table lives in oracle/pipeline_oracle.py.
__constant__ unsigned short kTargetStamp[15] = {992, 4088, 8188, 16382, 16382, 32319, 31775, 31775, 31775, 32319, 16382, 16382, 8188, 4088, 992};

__device__ __forceinline__ int to_int32_like_numpy(double v) {
  // ndarray.astype(np.int32) on x86-64: cvttsd2si -> truncation, 0x80000000 for NaN / out of range
  if (!(v > -2147483649.0 && v < 2147483648.0)) return (int)0x80000000;
  return (int)v;
}

__global__ void __launch_bounds__(256) draw_target_point_kernel(const double* __restrict__ tp, float* __restrict__ out) {
  const int b = blockIdx.y;
  // data.py:621-627, operation by operation (float64)
  double p0 = tp[2 * b + 0], p1 = tp[2 * b + 1];
  p1 = p1 + 1.3;
  p0 = p0 * 8.0;
  p1 = p1 * 8.0;
  p1 = p1 * -1.0;
  p1 = 256.0 - p1;
  p0 = p0 + 128.0;
  int cx = to_int32_like_numpy(p0), cy = to_int32_like_numpy(p1);
  cx = min(max(cx, 0), 256);
  cy = min(max(cy, 0), 256);
  float* o = out + (size_t)b * kMap * kMap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kMap * kMap; i += gridDim.x * blockDim.x) {
    const int r = i >> 8, c = i & 255;
    const int dr = r - cy + 7, dc = c - cx + 7;
    float v = 0.f;
    if (dr >= 0 && dr < 15 && dc >= 0 && dc < 15 && ((kTargetStamp[dr] >> dc) & 1)) v = 1.f;   // 255 / 255
    o[i] = v;
  }
}

// One thread per cropped pixel. rgb / depth: [B,H,W,3] uint8 (RGB order, as after cv2.cvtColor BGR2RGB), seg: [B,H,W] uint8.
__global__ void __launch_bounds__(256) camera_prep_kernel(const uint8_t* __restrict__ rgb, const uint8_t* __restrict__ depth,
                                                          const uint8_t* __restrict__ seg, const int* __restrict__ crop_shift,
                                                          const uint8_t* __restrict__ lut, int B, int H, int W, int ch, int cw,
                                                          float* __restrict__ rgb_nchw, float* __restrict__ rgb_nhwc_norm,
                                                          float* __restrict__ depth_out, int64_t* __restrict__ seg_out) {
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  const int64_t total = (int64_t)B * ch * cw;
  const int y0 = H / 2 - ch / 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % cw);
    const int y = (int)((i / cw) % ch);
    const int b = (int)(i / ((int64_t)cw * ch));
    const int x0 = W / 2 - cw / 2 + (crop_shift ? crop_shift[b] : 0);
    const int64_t src = ((int64_t)b * H + y0 + y) * W + x0 + x;
    if (rgb) {
      const uint8_t* s = rgb + src * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = (float)s[c];
        if (rgb_nchw) rgb_nchw[(((int64_t)b * 3 + c) * ch + y) * cw + x] = v;
        if (rgb_nhwc_norm) rgb_nhwc_norm[i * 3 + c] = ((v / 255.0f) - mean[c]) / stdv[c];   // == image_prep_kernel
      }
    }
    if (depth && depth_out) {
      const uint8_t* s = depth + src * 3;
      // get_depth (data.py:358-372) in float64: dot with [65536, 256, 1] (exact), / (256^3 - 1), clip to [0, 0.05], * 20
      double d = (double)s[0] * 65536.0 + (double)s[1] * 256.0 + (double)s[2];
      d = d / 16777215.0;
      d = fmin(fmax(d, 0.0), 0.05);
      depth_out[i] = (float)(d * 20.0);
    }
    if (seg && seg_out) seg_out[i] = (int64_t)lut[seg[src]];
  }
}

}  // namespace

// target_point: [B,2] float64 (local_command_point, data.py:352-355) -> out [B,1,256,256] float32 in {0,1}.
TFB_API int tfb_draw_target_point(const double* target_point, int B, float* out, cudaStream_t stream) {
  TFB_REQUIRE(target_point && out && B > 0 && B <= 65535);
  dim3 grid(16, B);
  draw_target_point_kernel<<<grid, 256, 0, stream>>>(target_point, out);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}

// Any of the input / output groups may be NULL. The caller guarantees 0 <= W/2 - cw/2 + crop_shift[b] <= W - cw
// (transfuser_b200.pipeline checks it on the host, where the shifts are produced).
TFB_API int tfb_camera_prep(const uint8_t* rgb, const uint8_t* depth, const uint8_t* seg, const int* crop_shift, const uint8_t* lut256,
                            int B, int H, int W, int crop_h, int crop_w, float* rgb_nchw, float* rgb_nhwc_norm, float* depth_out,
                            int64_t* seg_out, cudaStream_t stream) {
  TFB_REQUIRE(B > 0 && H > 0 && W > 0 && crop_h > 0 && crop_w > 0 && crop_h <= H && crop_w <= W);
  TFB_REQUIRE(!(seg && seg_out) || lut256);
  TFB_REQUIRE((rgb && (rgb_nchw || rgb_nhwc_norm)) || (depth && depth_out) || (seg && seg_out));
  camera_prep_kernel<<<tfb_grid((int64_t)B * crop_h * crop_w, 256), 256, 0, stream>>>(rgb, depth, seg, crop_shift, lut256, B, H, W, crop_h, crop_w,
                                                                                     rgb_nchw, rgb_nhwc_norm, depth_out, seg_out);
  TFB_CHECK_LAUNCH();
  return TFB_OK;
}
