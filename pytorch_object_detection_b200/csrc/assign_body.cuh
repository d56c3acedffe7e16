// The FCOS assignment rule (FCOSGenTargets.generate_target, model/modules/head.py:235-316) as device
// functions shared by assign.cu (targets only) and train_fused.cu (targets + losses + gradients).
//   l = x-x0, t = y-y0, r = x1-x, b = y1-y;  area = (l+r)*(t+b)            head.py:261-268
//   positive iff min(l,t,r,b) > 0, lo < max(l,t,r,b) <= hi, centre mask    head.py:272-283
//   winner = smallest area, lowest GT index on ties (torch.min's first index; head.py:285-286)
//   centerness = sqrt(min(l,r)*min(t,b) / (max(l,r)*max(t,b) + 1e-10))      head.py:294-299
//   no positive -> (0, -1, -1)                                             head.py:308-314
// Every fp32 operation is individually rounded (no FMA contraction), so labels, GT indices and
// regression targets are bit-exact with the reference.
#pragma once
#include "common.cuh"

namespace b200det {

// Level geometry of one launch (kernel parameter, by value).
struct AssignTable {
  int h[B200DET_MAX_LEVELS], w[B200DET_MAX_LEVELS], stride[B200DET_MAX_LEVELS], hw[B200DET_MAX_LEVELS];
  int point_off[B200DET_MAX_LEVELS + 1], tile_off[B200DET_MAX_LEVELS + 1];
  float lo[B200DET_MAX_LEVELS], hi[B200DET_MAX_LEVELS], radius[B200DET_MAX_LEVELS];
  int n_levels, num_points;
};

// Fills the table; tiles are `tile_points` consecutive points of one level.  Returns false on bad sizes.
inline bool make_assign_table(const int32_t* level_hw, const int32_t* strides, const float* limit_lo,
                              const float* limit_hi, const float* radius_px, int n_levels, int tile_points,
                              AssignTable* at) {
  long long off = 0;
  int toff = 0;
  for (int l = 0; l < B200DET_MAX_LEVELS; ++l) {
    const bool on = l < n_levels;
    if (on && (level_hw[2 * l] <= 0 || level_hw[2 * l + 1] <= 0 || strides[l] <= 0)) return false;
    at->h[l] = on ? level_hw[2 * l] : 0;
    at->w[l] = on ? level_hw[2 * l + 1] : 0;
    at->stride[l] = on ? strides[l] : 0;
    at->hw[l] = at->h[l] * at->w[l];
    at->lo[l] = on ? limit_lo[l] : 0.f;
    at->hi[l] = on ? limit_hi[l] : 0.f;
    at->radius[l] = on ? radius_px[l] : 0.f;
    at->point_off[l] = (int)off;
    at->tile_off[l] = toff;
    off += at->hw[l];
    toff += (at->hw[l] + tile_points - 1) / tile_points;
    if (off > (1ll << 30)) return false;
  }
  at->point_off[B200DET_MAX_LEVELS] = (int)off;
  at->tile_off[B200DET_MAX_LEVELS] = toff;
  at->n_levels = n_levels;
  at->num_points = (int)off;
  return true;
}

struct GtEntry {
  float x0, y0, x1, y1;
  float cx, cy;          // (x0+x1)/2, (y0+y1)/2 as the reference rounds them (head.py:276-277)
  int idx;
  int label;             // class label (fits int32), staged so the epilogue has no dependent global load
};

constexpr unsigned long long kNoWinner = ~0ull;

__device__ __forceinline__ GtEntry make_gt_entry(const float4 g, const int m, const int label) {
  const float cx = __fmul_rn(__fadd_rn(g.x, g.z), 0.5f);
  const float cy = __fmul_rn(__fadd_rn(g.y, g.w), 0.5f);
  return GtEntry{g.x, g.y, g.z, g.w, cx, cy, m, label};
}

// Conservative, rounding-safe pre-filter: can box g be positive for any point of a tile whose point
// centres span rows [row0, row1] of a level with this stride / range / radius?  (Margins of 1 px dwarf
// any fp32 rounding at image scale.)
//  * the centre mask needs |y - cy| < radius;
//  * a point strictly inside a box has max(l,t,r,b) in [max(w,h)/2, max(w,h)), so the level's (lo, hi]
//    range can only be met when max(w,h) > lo and max(w,h)/2 <= hi;
//  * side > 0 drops the -1 padding rows: a point cannot be strictly inside a degenerate box.
__device__ __forceinline__ bool gt_may_hit(const GtEntry& g, const int row0, const int row1, const int s,
                                           const float lo, const float hi, const float radius) {
  const int half = s / 2;
  const float ymin = (float)(row0 * s + half) - radius - 1.0f;
  const float ymax = (float)(row1 * s + half) + radius + 1.0f;
  const float side = fmaxf(g.x1 - g.x0, g.y1 - g.y0);
  return g.cy >= ymin && g.cy <= ymax && side > 0.f && side > lo - 1.0f && 0.5f * side <= hi + 1.0f;
}

// Half-width (in cells) of the window of grid points around a box centre that can pass the centre mask
// (one cell of slack).
__device__ __forceinline__ int window_half(const float radius, const int s) {
  return (int)ceilf(0.5f + radius / (float)s);
}

// One (box, window point k) pair: evaluates the reference's exact fp32 expressions.  Returns true when
// the window point exists in the level and is positive for the box; *pos = its row-major index in the
// level, *area = (l+r)*(t+b) (> 0, so its bit pattern is monotone).
__device__ __forceinline__ bool window_point_positive(const GtEntry& g, const int k, const int hwin, const int s,
                                                      const int w, const int h, const float lo, const float hi,
                                                      const float radius, int* pos, float* area) {
  const int wside = 2 * hwin + 1;
  const int half = s / 2;
  const float sf = (float)s;
  const int j = (int)floorf(g.cx / sf) + (k % wside) - hwin;
  const int i = (int)floorf(g.cy / sf) + (k / wside) - hwin;
  if (j < 0 || j >= w || i < 0 || i >= h) return false;
  const float x = (float)(j * s + half), y = (float)(i * s + half);
  // max(x-cx, y-cy, cx-x, cy-y) = max(|x-cx|, |y-cy|) exactly (fp32 subtraction is odd-symmetric)
  const float cmax = fmaxf(fabsf(__fsub_rn(x, g.cx)), fabsf(__fsub_rn(y, g.cy)));
  if (!(cmax < radius)) return false;
  const float lf = __fsub_rn(x, g.x0), tf = __fsub_rn(y, g.y0);
  const float rf = __fsub_rn(g.x1, x), bf = __fsub_rn(g.y1, y);
  const float omin = fminf(fminf(lf, tf), fminf(rf, bf));
  const float omax = fmaxf(fmaxf(lf, tf), fmaxf(rf, bf));
  if (!((omin > 0.f) && (omax > lo) && (omax <= hi))) return false;
  *pos = i * w + j;
  *area = __fmul_rn(__fadd_rn(lf, rf), __fadd_rn(tf, bf));
  return true;
}

// Targets of a positive point (col, row) of a level with this stride, assigned to box g.
__device__ __forceinline__ void positive_targets(const GtEntry& g, const int col, const int row, const int s,
                                                 float4* reg, float* cnt) {
  const int half = s / 2;
  const float x = (float)(col * s + half), y = (float)(row * s + half);
  *reg = make_float4(__fsub_rn(x, g.x0), __fsub_rn(y, g.y0), __fsub_rn(g.x1, x), __fsub_rn(g.y1, y));
  const float lr_min = fminf(reg->x, reg->z), lr_max = fmaxf(reg->x, reg->z);
  const float tb_min = fminf(reg->y, reg->w), tb_max = fmaxf(reg->y, reg->w);
  *cnt = __fsqrt_rn(__fdiv_rn(__fmul_rn(lr_min, tb_min), __fadd_rn(__fmul_rn(lr_max, tb_max), 1e-10f)));
}

}  // namespace b200det
