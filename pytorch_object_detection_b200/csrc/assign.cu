// K4a — FCOS target assignment (FCOSGenTargets.generate_target, model/modules/head.py:235-316).
//
// The reference materialises [B, HW, M, 4] offsets and reduces over M.  Here a CTA owns 2048
// consecutive points of one level of one image.  The centre-sampling mask
// (max(|x-cx|, |y-cy|) < 1.5*stride, head.py:275-283) means only ground-truth boxes whose centre
// row lies within 1.5 strides of the tile's rows can be positive for it, so the CTA first
// gathers those few boxes into shared memory (a conservative, rounding-safe pre-filter) and every
// point then evaluates the reference's exact fp32 expressions against that short list:
//   l = x-x0, t = y-y0, r = x1-x, b = y1-y;  area = (l+r)*(t+b)            head.py:261-268
//   positive iff min(l,t,r,b) > 0, lo < max(l,t,r,b) <= hi, centre mask    head.py:272-283
//   winner = smallest area, lowest GT index on ties (torch.min's first index; head.py:285-286)
//   centerness = sqrt(min(l,r)*min(t,b) / (max(l,r)*max(t,b) + 1e-10))      head.py:294-299
//   no positive -> (0, -1, -1)                                             head.py:308-314
// The kernel is a pure, fully coalesced write stream: 28 (+4) bytes per point, nothing re-read.
// (A positive box with area >= 99999999 px^2 would lose to the reference's sentinel; images are
// far smaller than 10^4 x 10^4, so that case is not modelled.)
#include "assign_body.cuh"

B200DET_TRACE_BUFFER(assign)
#ifdef B200DET_TRACE
#define g_trace_nlist(slot, v) (b200det::g_trace[slot] = (v))
#else
#define g_trace_nlist(slot, v) ((void)0)
#endif

namespace b200det {
namespace {

// Tile shape (threads x points per thread), chosen per launch: small batches are dominated by the
// per-CTA prologue, so they get wide CTAs with short tiles; large batches get more, leaner CTAs.
// Measured on B200 (28 B written per point): <256,6> 6.6 us at B=32 (one wave; <256,4> is 1.46 waves and
// 7.2 us); <128,8> 5.0 TB/s at B=128.
constexpr long long kAssignSmallPoints = 1500000;   // B*P below this -> <256,4>

template <int kAssignThreads, int kAssignPts>
__global__ void __launch_bounds__(kAssignThreads, 768 / kAssignThreads)
assign_targets_kernel(const AssignTable at, const int M, const float* __restrict__ gt_boxes,
                      const long long* __restrict__ gt_labels, long long* __restrict__ cls_t,
                      float* __restrict__ cnt_t, float* __restrict__ reg_t, int32_t* __restrict__ gt_index) {
  constexpr int kAssignTile = kAssignThreads * kAssignPts;               // points of one level of one image per CTA
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GtEntry* gts = reinterpret_cast<GtEntry*>(smem_raw);                   // [M] the image's boxes, by GT index
  int* cand = reinterpret_cast<int*>(smem_raw + (size_t)M * sizeof(GtEntry));   // [M] GT indices relevant to this tile
  __shared__ unsigned long long keys[kAssignTile];                       // per point: (area bits << 32) | GT index
  __shared__ int s_n;

  // grid = (image, tile) with the tile order reversed: the coarse levels are scheduled first and
  // consecutive CTAs (same tile, different images) cost the same, so expensive tiles spread over SMs.
  const int b = blockIdx.x;
  const int tile = (int)gridDim.y - 1 - (int)blockIdx.y;
  int l = 0;
#pragma unroll
  for (int i = 1; i < B200DET_MAX_LEVELS; ++i) l += (i < at.n_levels && tile >= at.tile_off[i]) ? 1 : 0;
  const int hw = at.hw[l], w = at.w[l], h = at.h[l], s = at.stride[l];
  const int t0 = (tile - at.tile_off[l]) * kAssignTile;
  const int t1 = min(t0 + kAssignTile, hw) - 1;
  const float lo = at.lo[l], hi = at.hi[l], radius = at.radius[l];

  const bool traced = blockIdx.x == 0 && (tile == 0 || tile == (int)gridDim.y - 1);
  const int tslot = tile == 0 ? 0 : 8;
  B200DET_STAMP_IF(traced, tslot + 0);
  if (threadIdx.x == 0) s_n = 0;
#pragma unroll
  for (int q = 0; q < kAssignPts; ++q) keys[threadIdx.x + q * kAssignThreads] = kNoWinner;
  __syncthreads();
  {
    // Stage the image's boxes; keep the indices of those that can be positive somewhere in this tile
    // (gt_may_hit: a conservative, rounding-safe pre-filter on rows, level range and padding rows).
    const float4* g4 = reinterpret_cast<const float4*>(gt_boxes) + (size_t)b * M;
    const long long* lab = gt_labels + (size_t)b * M;
    for (int m = threadIdx.x; m < M; m += kAssignThreads) {
      const GtEntry g = make_gt_entry(g4[m], m, (int)lab[m]);
      gts[m] = g;
      if (gt_may_hit(g, t0 / w, t1 / w, s, lo, hi, radius)) cand[atomicAdd(&s_n, 1)] = m;
    }
  }
  __syncthreads();
  const int n_list = s_n;
  B200DET_STAMP_IF(traced, tslot + 1);
  if (traced && threadIdx.x == 0) g_trace_nlist(tslot + 3, n_list);

  // Box-centric pass: only the points within the centre radius of a box can be positive for it
  // (head.py:275-283), i.e. a (2*hwin+1)^2 window of grid points around its centre cell (hwin has
  // one cell of slack).  Every (box, window point) pair evaluates the reference's exact fp32
  // expressions; positives race with a 64-bit atomicMin on (area, GT index): smallest area, lowest
  // index on ties = torch.min's first index on the masked areas (head.py:285-286).
  const int hwin = window_half(radius, s);
  const int wcount = (2 * hwin + 1) * (2 * hwin + 1);
  for (int pi = threadIdx.x; pi < n_list * wcount; pi += kAssignThreads) {
    const int e = pi / wcount, k = pi - e * wcount;
    window_vote(gts[cand[e]], k, hwin, s, w, h, t0, t1, lo, hi, radius, keys);
  }
  __syncthreads();

  // Point pass: one coalesced write stream.  (row, col) advance incrementally in fp32 (exact below 2^24).
  const size_t out0 = (size_t)b * at.num_points + at.point_off[l];
  const int p_first = t0 + threadIdx.x;
  int row = p_first / w, col = p_first - row * w;
  const int drow = kAssignThreads / w, dcol = kAssignThreads - drow * w;
#pragma unroll
  for (int q = 0; q < kAssignPts; ++q) {
    const int pos = p_first + q * kAssignThreads;            // strided: every store instruction is coalesced
    if (pos < hw) {
      const unsigned long long key = keys[pos - t0];
      long long label = 0;
      float cnt = -1.f;
      float4 reg = make_float4(-1.f, -1.f, -1.f, -1.f);
      int best_m = -1;
      if (key != kNoWinner) {
        best_m = (int)(unsigned)(key & 0xffffffffull);
        const GtEntry g = gts[best_m];
        positive_targets(g, col, row, s, &reg, &cnt);
        label = (long long)g.label;
      }
      const size_t o = out0 + pos;
      stg_stream_s64(cls_t + o, label);
      stg_stream_f1(cnt_t + o, cnt);
      stg_stream_f4(reg_t + 4 * o, reg);
      if (gt_index) gt_index[o] = best_m;
    }
    row += drow;
    col += dcol;
    if (col >= w) { col -= w; ++row; }
  }
  B200DET_STAMP_IF(traced, tslot + 2);
}

}  // namespace
}  // namespace b200det

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
  const size_t smem = (size_t)max_gt * (sizeof(GtEntry) + sizeof(int));
  if (smem > 200 * 1024) return B200DET_ERR_UNSUPPORTED;
  long long total_points = 0;
  for (int l = 0; l < n_levels; ++l) total_points += (long long)level_hw[2 * l] * level_hw[2 * l + 1];
  // Tile shape.  Small problems are a handful of waves, so the shape is chosen against WAVE QUANTISATION: a
  // grid of 1.46 waves (config 3 with <256,4>: 864 CTAs on 592 resident slots) idles a quarter of the machine.
  // Candidates <256 threads, 4 / 6 / 8 points>; cost model = waves * (fixed CTA latency + per-point time),
  // constants from the phase trace (scripts/trace_assign.py): ~1.2 us fixed, ~0.21 us per point per thread.
  // (Writing every point as a negative first, so that the store stream starts before the GT staging, and
  // patching the positives afterwards was measured too: 8.1 us against 6.6 us — the early stores delay the
  // GT loads behind them.)
  const bool small = (long long)batch * total_points < kAssignSmallPoints;
  int pts = 8;
  if (small) {
    static int sm_count = 0;
    if (!sm_count) {
      int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        sm_count = 148;
    }
    const long long slots = (long long)sm_count * 4;                   // 256 threads, <= 64 registers: 4 CTAs per SM
    double best = 1e30;
    for (int cand = 4; cand <= 8; cand += 2) {
      long long tiles = 0;
      for (int l = 0; l < n_levels; ++l)
        tiles += ((long long)level_hw[2 * l] * level_hw[2 * l + 1] + 256 * cand - 1) / (256 * cand);
      const long long waves = (tiles * batch + slots - 1) / slots;
      const double cost = (double)waves * (1.2 + 0.21 * cand);
      if (cost < best) { best = cost; pts = cand; }
    }
  }
  const int tile_points = small ? 256 * pts : 128 * 8;
  AssignTable at;
  if (!make_assign_table(level_hw, strides, limit_lo, limit_hi, radius_px, n_levels, tile_points, &at))
    return B200DET_ERR_ARG;
  const int toff = at.tile_off[B200DET_MAX_LEVELS];
  if (toff > 65535) return B200DET_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto launch = [&](auto kernel, int threads) -> int {
    if (smem > 40 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
    }
    kernel<<<dim3(batch, toff), threads, smem, st>>>(at, max_gt, gt_boxes, reinterpret_cast<const long long*>(gt_labels),
                                                     reinterpret_cast<long long*>(cls_t), cnt_t, reg_t, gt_index);
    return B200DET_OK;
  };
  const int rc = !small     ? launch(assign_targets_kernel<128, 8>, 128)
                 : pts == 4 ? launch(assign_targets_kernel<256, 4>, 256)
                 : pts == 6 ? launch(assign_targets_kernel<256, 6>, 256)
                            : launch(assign_targets_kernel<256, 8>, 256);
  if (rc) return rc;
  return check_launch();
}
