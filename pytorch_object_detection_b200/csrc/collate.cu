// N4 — the step before the path: the datasets' collate_fn (dataset/voc.py:141-173, dataset/coco.py:135-165)
// on the device, so that a batch reaches FCOSGenTargets / FCOSTargetLoss without per-image host work.
//
//   pack_gt_kernel          ragged per-image GT lists (one flat upload + offsets) -> gt_boxes [B,M,4] and
//                           gt_labels [B,M], padded with -1 (torch.nn.functional.pad(..., value=-1) + stack)
//   collate_images_kernel   per-image [C,h,w] -> [B,C,H,W]: zero pad to the batch maximum, THEN
//                           transforms.Normalize(mean, std), so a padded pixel is (0 - mean) / std exactly as in
//                           the reference (it normalises the padded tensor); (x - mean) / std with each op rounded.
// Both are pure streams: every output element is written once, every input element read once.
#include "common.cuh"

namespace b200det {
namespace {

__global__ void __launch_bounds__(256)
pack_gt_kernel(const float4* __restrict__ flat_boxes, const long long* __restrict__ flat_labels,
               const int32_t* __restrict__ offsets, const int max_gt, float4* __restrict__ gt_boxes,
               long long* __restrict__ gt_labels) {
  const int b = blockIdx.y;
  const int lo = offsets[b], n = offsets[b + 1] - lo;
  for (int m = blockIdx.x * 256 + threadIdx.x; m < max_gt; m += gridDim.x * 256) {
    const bool real = m < n;
    gt_boxes[(size_t)b * max_gt + m] = real ? flat_boxes[lo + m] : make_float4(-1.f, -1.f, -1.f, -1.f);
    gt_labels[(size_t)b * max_gt + m] = real ? flat_labels[lo + m] : -1ll;
  }
}

constexpr int kCollateMaxBatch = 128;      // images per launch (pointer table passed by value)
constexpr int kCollateMaxChannels = 4;

struct CollateTable {
  const float* img[kCollateMaxBatch];
  int h[kCollateMaxBatch], w[kCollateMaxBatch];
  float mean[kCollateMaxChannels], std[kCollateMaxChannels];
};

__global__ void __launch_bounds__(256)
collate_images_kernel(const CollateTable t, const int channels, const int out_h, const int out_w,
                      float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int h = t.h[b], w = t.w[b];
  const float mean = t.mean[c], sd = t.std[c];
  const float* __restrict__ src = t.img[b] + (size_t)c * h * w;
  float* __restrict__ dst = out + ((size_t)b * channels + c) * out_h * out_w;
  const float pad = __fdiv_rn(__fsub_rn(0.f, mean), sd);
  const int quads = (out_w + 3) / 4;                         // 4 consecutive pixels of a row per thread
  const bool vec = (out_w % 4 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
  for (int i = blockIdx.x * 256 + threadIdx.x; i < out_h * quads; i += gridDim.x * 256) {
    const int y = i / quads, x0 = (i - y * quads) * 4;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + k;
      v[k] = (y < h && x < w) ? __fdiv_rn(__fsub_rn(ldg_stream_f1(src + (size_t)y * w + x), mean), sd) : pad;
    }
    float* o = dst + (size_t)y * out_w + x0;
    if (vec) {
      stg_stream_f4(o, make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (x0 + k < out_w) stg_stream_f1(o + k, v[k]);
    }
  }
}

}  // namespace
}  // namespace b200det

using namespace b200det;

extern "C" int b200det_pack_gt(const float* flat_boxes, const int64_t* flat_labels, const int32_t* offsets,
                               int batch, int max_gt, float* gt_boxes, int64_t* gt_labels, void* stream) {
  if (batch <= 0 || batch > 65535 || max_gt < 0 || !offsets) return B200DET_ERR_ARG;
  if (max_gt == 0) return B200DET_OK;
  if (!gt_boxes || !gt_labels || !aligned16(flat_boxes) || !aligned16(gt_boxes)) return B200DET_ERR_ARG;
  pack_gt_kernel<<<dim3((max_gt + 255) / 256, batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(flat_boxes), reinterpret_cast<const long long*>(flat_labels), offsets, max_gt,
      reinterpret_cast<float4*>(gt_boxes), reinterpret_cast<long long*>(gt_labels));
  return check_launch();
}

extern "C" int b200det_collate_images(const void* const* images, const int32_t* image_hw, int batch, int channels,
                                      int out_h, int out_w, const float* mean, const float* std, float* out,
                                      void* stream) {
  if (!images || !image_hw || batch <= 0 || channels <= 0 || channels > kCollateMaxChannels || out_h <= 0 ||
      out_w <= 0 || !mean || !std || !out)
    return B200DET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int b0 = 0; b0 < batch; b0 += kCollateMaxBatch) {
    const int nb = batch - b0 < kCollateMaxBatch ? batch - b0 : kCollateMaxBatch;
    CollateTable t = {};
    for (int i = 0; i < nb; ++i) {
      const int h = image_hw[2 * (b0 + i)], w = image_hw[2 * (b0 + i) + 1];
      if (!images[b0 + i] || h <= 0 || w <= 0 || h > out_h || w > out_w) return B200DET_ERR_ARG;
      t.img[i] = static_cast<const float*>(images[b0 + i]);
      t.h[i] = h;
      t.w[i] = w;
    }
    for (int c = 0; c < channels; ++c) {
      if (!(std[c] != 0.f)) return B200DET_ERR_ARG;
      t.mean[c] = mean[c];
      t.std[c] = std[c];
    }
    const int quads = out_h * ((out_w + 3) / 4);
    int gx = (quads + 255) / 256;
    if (gx > 1184) gx = 1184;                                 // 8 CTAs x 148 SMs; the loop strides over the rest
    collate_images_kernel<<<dim3(gx, channels, nb), 256, 0, st>>>(t, channels, out_h, out_w,
                                                                  out + (size_t)b0 * channels * out_h * out_w);
    const int rc = check_launch();
    if (rc) return rc;
  }
  return B200DET_OK;
}
