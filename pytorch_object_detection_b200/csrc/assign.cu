// K4a — FCOS target assignment (FCOSGenTargets.generate_target, model/modules/head.py:235-316).
//
// The reference materialises [B, HW, M, 4] offsets and reduces over M.  Here ONE launch of
// assign_stream_kernel<kLoss = false> (assign_stream.cuh, which describes the design): a CTA owns a tile of one level
// of one image; its fill warp writes the tile as negatives with bulk copies while the other warps gather the few
// ground-truth boxes whose centre can reach the tile, evaluate the reference's exact fp32 expressions
//   l = x-x0, t = y-y0, r = x1-x, b = y1-y;  area = (l+r)*(t+b)            head.py:261-268
//   positive iff min(l,t,r,b) > 0, lo < max(l,t,r,b) <= hi, centre mask    head.py:272-283
//   winner = smallest area, lowest GT index on ties (torch.min's first index; head.py:285-286)
//   centerness = sqrt(min(l,r)*min(t,b) / (max(l,r)*max(t,b) + 1e-10))      head.py:294-299
//   no positive -> (0, -1, -1)                                             head.py:308-314
// for (box, point in reach) pairs only and patch the positives on top of the fill.  28 (+4) bytes written per
// point, nothing re-read.  (A positive box with area >= 99999999 px^2 would lose to the reference's sentinel;
// images are far smaller than 10^4 x 10^4, so that case is not modelled.)
#include "common.cuh"

B200DET_TRACE_BUFFER(assign)

#include "assign_stream.cuh"

#include <stdlib.h>

extern "C" int b200det_assign_targets(const int32_t* level_hw, const int32_t* strides, const float* limit_lo,
                                      const float* limit_hi, const float* radius_px, int n_levels, int batch,
                                      int max_gt, const float* gt_boxes, const int64_t* gt_labels, int64_t* cls_t,
                                      float* cnt_t, float* reg_t, int32_t* gt_index, void* stream) {
  using namespace b200det;
  if (!level_hw || !strides || !limit_lo || !limit_hi || !radius_px || n_levels <= 0 ||
      n_levels > B200DET_MAX_LEVELS || batch <= 0 || max_gt < 0 || !cls_t || !cnt_t || !reg_t)
    return B200DET_ERR_ARG;
  if (max_gt > 0 && (!gt_boxes || !gt_labels)) return B200DET_ERR_ARG;
  if (!aligned16(gt_boxes) || !aligned16(reg_t)) return B200DET_ERR_ARG;
  const size_t smem = (size_t)max_gt * (sizeof(GtEntry) + 3 * sizeof(int));
  if (smem > 160 * 1024) return B200DET_ERR_UNSUPPORTED;
  if (batch > 65535) return B200DET_ERR_UNSUPPORTED;
  const int pts = stream_points_per_thread(level_hw, n_levels, batch);
  AssignTable at;
  if (!make_assign_table(level_hw, strides, limit_lo, limit_hi, radius_px, n_levels, kStreamThreads * pts, &at))
    return B200DET_ERR_ARG;
  const int n_tiles = at.tile_off[B200DET_MAX_LEVELS];
  StreamArgs a = {};
  a.M = max_gt;
  a.n_tiles = n_tiles;
  a.gt_boxes = gt_boxes;
  a.gt_labels = reinterpret_cast<const long long*>(gt_labels);
  a.cls_t = reinterpret_cast<long long*>(cls_t);
  a.cnt_t = cnt_t;
  a.reg_t = reg_t;
  a.gt_index = gt_index;
  a.batch = batch;
  const LossMaps lm = {};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static const bool no_pdl = getenv("B200DET_NO_PDL") && getenv("B200DET_NO_PDL")[0] == '1';
  auto go = [&](auto kernel) -> int {
    cudaError_t e = launch_stream_kernel(kernel, dim3(n_tiles, batch), smem, st, at, lm, a, !no_pdl);
    if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
    return check_launch();
  };
  return pts == 4 ? go(assign_stream_kernel<4, false, false>)
         : pts == 6 ? go(assign_stream_kernel<6, false, false>)
                    : go(assign_stream_kernel<8, false, false>);
}
