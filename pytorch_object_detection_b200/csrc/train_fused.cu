// K4 fused — target assignment + box (IoU / GIoU) loss + centerness BCE loss, forward AND backward.
// Replaces, for one training step,
//   FCOSGenTargets.forward            model/modules/head.py:218-316
//   compute_reg_loss (+ giou / iou)   model/loss.py:116-177      and its autograd backward
//   compute_cnt_loss                  model/loss.py:29-57        and its autograd backward
//   the two `.mean()` of FCOSLoss     model/loss.py:210-213
//
// The gradient of a positive carries 1 / num_pos of its image, so num_pos must be known before the
// single streaming pass can write final gradients.  Three launches chained by programmatic dependent
// launch (each starts while its predecessor runs and waits only where it needs the predecessor's result):
//   count_positives_kernel  one cluster of 8 CTAs per image; a warp expands a (box, level) pair, its lanes
//                           the window points; positives set a bit in the image's bitmap, held in the
//                           shared memory of the cluster's CTA 0 (32-bit atomicOr over distributed shared
//                           memory); popcount -> num_pos.  ~4 us of latency, hidden behind:
//   assign_loss_tile_kernel a CTA owns 1024 points of one level of one image (the tiling of assign.cu):
//                           stage the GT boxes, box-centric vote into shared memory, then ONE write
//                           stream: every point's targets (28 B) and gradients (20 B; zeros off the
//                           positives).  Predictions are fetched at positives only (all of a thread's
//                           loads in flight together); loss terms and unscaled gradients are evaluated
//                           before the wait for num_pos, only the scaling and the stores come after it.
//   finalize_losses_kernel  one CTA: tile partials added in tile order, per-image losses in image order
//                           (deterministic), batch means.
// Nothing is re-read from HBM: the assignment never leaves shared memory.
// HBM traffic: 48 B written per point + ~36 B read per positive; GT boxes are read once per CTA.
// Measured (B200, config 3: B=32, P=23265, M<=100): 16.9 us for the three launches against 26.6 us for
// assign + box-loss forward + backward as separate kernels (scripts/time_fused.py).
// Tried and dropped: one cluster per image doing everything (the cluster barrier is a GPU-scope fence
// that waits for the CTA's streaming stores to drain; 64-bit atomicMin over distributed shared memory is
// not atomic against the local CAS loop) and in-kernel tickets for the final reduction (5 us tail).
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "assign_body.cuh"
#include "loss_terms.cuh"

B200DET_TRACE_BUFFER(train)

namespace cg = cooperative_groups;

namespace b200det {
namespace {

struct LossMaps {
  const float* reg[B200DET_MAX_LEVELS];
  const float* reg_scale[B200DET_MAX_LEVELS];   // ScaleExp folded in (common.cuh); NULL = reg holds the distances
  const float* cnt[B200DET_MAX_LEVELS];
  float* greg[B200DET_MAX_LEVELS];
  float* gcnt[B200DET_MAX_LEVELS];
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- num_pos[b] = clamp(#points of image b with at least one positive box, 1) ------------------------
// One cluster of kCountSlices CTAs per image.  The (box, level) pairs of the image are dealt to the CTAs;
// one thread tests one pair against the level-range / padding pre-filter, survivors go to a shared list
// which the warps expand round-robin, lanes = window points.  A positive point sets its bit in the
// image's bitmap, which lives in the shared memory of CTA 0 of the cluster (32-bit atomicOr through
// distributed shared memory); after the cluster barrier CTA 0 popcounts it.
constexpr int kCountSlices = 8;
constexpr int kCountThreads = 256;

__global__ void __cluster_dims__(kCountSlices, 1, 1) __launch_bounds__(kCountThreads)
count_positives_kernel(const AssignTable at, const int M, const float* __restrict__ gt_boxes,
                       float* __restrict__ num_pos) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GtEntry* gts = reinterpret_cast<GtEntry*>(smem_raw);                                      // [M]
  unsigned* bitmap = reinterpret_cast<unsigned*>(smem_raw + (size_t)M * sizeof(GtEntry));   // [ceil(P / 32)], CTA 0's is used
  __shared__ float s_red[32];
  __shared__ int s_list[kCountThreads];
  __shared__ int s_list_n;
  pdl_launch_dependents();                       // the streaming kernel may start; it waits before reading num_pos
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y, tid = threadIdx.x;
  const int words = (at.num_points + 31) / 32;
  if (rank == 0)
    for (int i = tid; i < words; i += kCountThreads) bitmap[i] = 0u;
  const float4* g4 = reinterpret_cast<const float4*>(gt_boxes) + (size_t)b * M;
  for (int m = tid; m < M; m += kCountThreads) gts[m] = make_gt_entry(g4[m], m, 0);
  cluster.sync();                                // CTA 0's bitmap is clear; every CTA of the cluster runs
  unsigned* image_bits = cluster.map_shared_rank(bitmap, 0);

  // This CTA's pairs are q = rank, rank + kCountSlices, ... (levels and boxes mix evenly).
  const int lane = tid & 31, warp = tid >> 5;
  const int n_pairs = M * at.n_levels;
  for (int q0 = 0; q0 < n_pairs; q0 += kCountSlices * kCountThreads) {
    __syncthreads();
    if (tid == 0) s_list_n = 0;
    __syncthreads();
    const int q = q0 + tid * kCountSlices + rank;
    if (q < n_pairs) {
      const int l = q / M, m = q - l * M;
      const float side = fmaxf(gts[m].x1 - gts[m].x0, gts[m].y1 - gts[m].y0);   // pre-filter of gt_may_hit
      if (side > 0.f && side > at.lo[l] - 1.0f && 0.5f * side <= at.hi[l] + 1.0f)
        s_list[atomicAdd(&s_list_n, 1)] = q;
    }
    __syncthreads();
    const int n_list = s_list_n;
    for (int e = warp; e < n_list; e += kCountThreads / 32) {
      const int pair = s_list[e];
      const int l = pair / M, m = pair - l * M;
      const GtEntry g = gts[m];
      const int s = at.stride[l];
      const float radius = at.radius[l];
      const int hwin = window_half(radius, s);
      const int wcount = (2 * hwin + 1) * (2 * hwin + 1);
      for (int k = lane; k < wcount; k += 32) {
        int pos;
        float area;
        if (window_point_positive(g, k, hwin, s, at.w[l], at.h[l], at.lo[l], at.hi[l], radius, &pos, &area)) {
          const int p = at.point_off[l] + pos;
          atomicOr(image_bits + (p >> 5), 1u << (p & 31));
        }
      }
    }
  }
  cluster.sync();                                // every vote has landed in CTA 0
  if (rank != 0) return;
  int c = 0;
  for (int i = tid; i < words; i += kCountThreads) c += __popc(bitmap[i]);
  const float total = block_sum_f((float)c, s_red);             // <= P < 2^24: exact in fp32
  if (tid == 0) num_pos[b] = fmaxf(total, 1.f);
}

// ---- the streaming kernel -----------------------------------------------------------------------------
constexpr int kTrainThreads = 256;

constexpr int kTrainPts = 4;                  // 8 points per thread spills at the 64 registers 4 CTAs / SM allow
constexpr int kTrainTile = kTrainThreads * kTrainPts;

template <bool kScaleExp>                      // some level carries a folded ScaleExp (raw regression outputs)
__global__ void __launch_bounds__(kTrainThreads, 3)
assign_loss_tile_kernel(const AssignTable at, const LossMaps lm, const int has_cnt, const int M,
                        const float* __restrict__ gt_boxes, const long long* __restrict__ gt_labels, const int mode,
                        const float* __restrict__ grad_box, const float* __restrict__ grad_cnt, const int grad_mode,
                        const float inv_batch,
                        const float* num_pos, long long* __restrict__ cls_t, float* __restrict__ cnt_t,
                        float* __restrict__ reg_t, float* partial, const int use_pdl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GtEntry* gts = reinterpret_cast<GtEntry*>(smem_raw);                   // [M] the image's boxes, by GT index
  int* cand = reinterpret_cast<int*>(smem_raw + (size_t)M * sizeof(GtEntry));   // [M] GT indices relevant to this tile
  __shared__ unsigned long long keys[kTrainTile];                        // per point: (area bits << 32) | GT index
  __shared__ float s_red[32];
  __shared__ int s_n;

  // grid = (image, tile), tile order reversed (coarse levels first), as in assign.cu
  const int b = blockIdx.x;
  const int n_tiles = (int)gridDim.y;
  const int tile = n_tiles - 1 - (int)blockIdx.y;
  const int tid = threadIdx.x;
  int l = 0;
#pragma unroll
  for (int i = 1; i < B200DET_MAX_LEVELS; ++i) l += (i < at.n_levels && tile >= at.tile_off[i]) ? 1 : 0;
  const int hw = at.hw[l], w = at.w[l], h = at.h[l], s = at.stride[l];
  const int t0 = (tile - at.tile_off[l]) * kTrainTile;
  const int t1 = min(t0 + kTrainTile, hw) - 1;
  const float lo = at.lo[l], hi = at.hi[l], radius = at.radius[l];
  const float* __restrict__ reg_p = lm.reg[l] + (size_t)b * 4 * hw;
  const float* __restrict__ cnt_p = has_cnt ? lm.cnt[l] + (size_t)b * hw : nullptr;
  float* __restrict__ greg = lm.greg[l] + (size_t)b * 4 * hw;
  float* __restrict__ gcnt = has_cnt ? lm.gcnt[l] + (size_t)b * hw : nullptr;
  const bool traced = b == 0 && (tile == 0 || tile == n_tiles - 1);
  const int tslot = tile == 0 ? 0 : 16;
  B200DET_STAMP_IF(traced, tslot + 0);
  if (use_pdl) pdl_launch_dependents();          // finalize_losses_kernel may become resident; it waits for this grid

  if (tid == 0) s_n = 0;
#pragma unroll
  for (int q = 0; q < kTrainPts; ++q) keys[tid + q * kTrainThreads] = kNoWinner;
  __syncthreads();
  {
    const float4* g4 = reinterpret_cast<const float4*>(gt_boxes) + (size_t)b * M;
    const long long* lab = gt_labels + (size_t)b * M;
    for (int m = tid; m < M; m += kTrainThreads) {
      const GtEntry g = make_gt_entry(g4[m], m, (int)lab[m]);
      gts[m] = g;
      if (gt_may_hit(g, t0 / w, t1 / w, s, lo, hi, radius)) cand[atomicAdd(&s_n, 1)] = m;
    }
  }
  __syncthreads();
  const int n_list = s_n;
  B200DET_STAMP_IF(traced, tslot + 1);
  const int hwin = window_half(radius, s);
  const int wcount = (2 * hwin + 1) * (2 * hwin + 1);
  for (int pi = tid; pi < n_list * wcount; pi += kTrainThreads) {
    const int e = pi / wcount, k = pi - e * wcount;
    window_vote(gts[cand[e]], k, hwin, s, w, h, t0, t1, lo, hi, radius, keys);
  }
  __syncthreads();
  B200DET_STAMP_IF(traced, tslot + 2);

  // ---- pass A: the targets of every point, zero gradients of the negatives; at positives the
  //      predictions are fetched (all of a thread's points in flight together) and the loss terms and
  //      unscaled gradients evaluated — everything that does not need num_pos ------------------------------
  const size_t out0 = (size_t)b * at.num_points + at.point_off[l];
  const int p_first = t0 + tid;
  unsigned pos_mask = 0;                                      // bit q: this thread's q-th point is positive
  float4 pr[kTrainPts], tg[kTrainPts];
  float px[kTrainPts], ct[kTrainPts];
  {
    int row = p_first / w, col = p_first - row * w;
    const int drow = kTrainThreads / w, dcol = kTrainThreads - drow * w;
#pragma unroll
    for (int q = 0; q < kTrainPts; ++q) {
      const int pos = p_first + q * kTrainThreads;            // strided: every store instruction is coalesced
      pr[q] = tg[q] = make_float4(-1.f, -1.f, -1.f, -1.f);
      px[q] = 0.f;
      ct[q] = -1.f;
      if (pos < hw) {
        const unsigned long long key = keys[pos - t0];
        long long label = 0;
        if (key != kNoWinner) {
          const GtEntry g = gts[(unsigned)(key & 0xffffffffull)];
          positive_targets(g, col, row, s, &tg[q], &ct[q]);
          label = (long long)g.label;
          pos_mask |= 1u << q;
          pr[q] = make_float4(reg_p[pos], reg_p[hw + pos], reg_p[2 * hw + pos], reg_p[3 * hw + pos]);
          if (has_cnt) px[q] = cnt_p[pos];
        } else {
          stg_stream_f1(greg + pos, 0.f);
          stg_stream_f1(greg + hw + pos, 0.f);
          stg_stream_f1(greg + 2 * hw + pos, 0.f);
          stg_stream_f1(greg + 3 * hw + pos, 0.f);
          if (has_cnt) stg_stream_f1(gcnt + pos, 0.f);
        }
        const size_t o = out0 + pos;
        stg_stream_s64(cls_t + o, label);
        stg_stream_f1(cnt_t + o, ct[q]);
        stg_stream_f4(reg_t + 4 * o, tg[q]);
      }
      row += drow;
      col += dcol;
      if (col >= w) { col -= w; ++row; }
    }
  }
  float acc_box = 0.f, acc_cnt = 0.f, acc_dsc = 0.f;
  if (pos_mask) {
    const float sc = (kScaleExp && lm.reg_scale[l]) ? __ldg(lm.reg_scale[l]) : 0.f;
#pragma unroll
    for (int q = 0; q < kTrainPts; ++q) {
      if (!(pos_mask & (1u << q))) continue;
      float4 g;
      if (kScaleExp && lm.reg_scale[l]) {
        // raw regression output x: distances d = exp(x * sc) (ScaleExp); dL/dx = dL/dd * d * sc,
        // dL/dsc += dL/dd * d * x
        const float4 x4 = pr[q];
        const float4 d4 = make_float4(scale_exp_f32(x4.x, sc), scale_exp_f32(x4.y, sc), scale_exp_f32(x4.z, sc),
                                      scale_exp_f32(x4.w, sc));
        acc_box += box_term<true>(d4, tg[q], mode, &g);
        g = make_float4(g.x * d4.x, g.y * d4.y, g.z * d4.z, g.w * d4.w);
        acc_dsc += (g.x * x4.x + g.y * x4.y) + (g.z * x4.z + g.w * x4.w);
        g = make_float4(g.x * sc, g.y * sc, g.z * sc, g.w * sc);
      } else {
        acc_box += box_term<true>(pr[q], tg[q], mode, &g);
      }
      pr[q] = g;                                              // unscaled d(loss term) / d(reg map)
      if (has_cnt) {
        acc_cnt += bce_term(px[q], ct[q]);
        px[q] = sigmoid_f32(px[q]) - ct[q];
      }
    }
  }
  B200DET_STAMP_IF(traced, tslot + 3);

  // this tile's loss partials: fixed shuffle tree per warp, then the warps in order (one barrier)
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    acc_box += __shfl_xor_sync(0xffffffffu, acc_box, d);
    acc_cnt += __shfl_xor_sync(0xffffffffu, acc_cnt, d);
    acc_dsc += __shfl_xor_sync(0xffffffffu, acc_dsc, d);
  }
  if ((tid & 31) == 0) {
    s_red[3 * (tid >> 5)] = acc_box;
    s_red[3 * (tid >> 5) + 1] = acc_cnt;
    s_red[3 * (tid >> 5) + 2] = acc_dsc;
  }

  // ---- pass B: scale the positives' gradients by grad_loss[b] / num_pos[b] and write them --------------
  if (use_pdl) pdl_wait();                       // count_positives_kernel has completed and is visible
  if (pos_mask) {
    const float np = __ldcg(num_pos + b);
    const float scale_box = upstream_of(grad_box, grad_mode, b, inv_batch) / np;
    const float scale_cnt = upstream_of(grad_cnt, grad_mode, b, inv_batch) / np;
#pragma unroll
    for (int q = 0; q < kTrainPts; ++q) {
      if (!(pos_mask & (1u << q))) continue;
      const int pos = p_first + q * kTrainThreads;
      stg_stream_f1(greg + pos, pr[q].x * scale_box);
      stg_stream_f1(greg + hw + pos, pr[q].y * scale_box);
      stg_stream_f1(greg + 2 * hw + pos, pr[q].z * scale_box);
      stg_stream_f1(greg + 3 * hw + pos, pr[q].w * scale_box);
      if (has_cnt) stg_stream_f1(gcnt + pos, scale_cnt * px[q]);
    }
  }
  B200DET_STAMP_IF(traced, tslot + 4);

  // finalize_losses_kernel adds the tile partials in tile order
  __syncthreads();
  if (tid == 0) {
    float tb = 0.f, tc = 0.f, td = 0.f;
#pragma unroll
    for (int wi = 0; wi < kTrainThreads / 32; ++wi) {
      tb += s_red[3 * wi];
      tc += s_red[3 * wi + 1];
      td += s_red[3 * wi + 2];
    }
    *reinterpret_cast<float4*>(partial + ((size_t)b * n_tiles + tile) * 4) = make_float4(tb, tc, td, 0.f);
  }
  B200DET_STAMP_IF(traced, tslot + 5);
}

// ---- per-image losses (tile partials added in tile order), batch means (image order) ------------------
// One CTA, launched as a programmatic dependent of the streaming kernel: it is resident before that kernel
// ends and proceeds as soon as its partials are complete and visible.
constexpr int kFinalThreads = 256;
constexpr int kFinalStage = 1024;            // tile partials (float4: box, cnt, d/d scale, -) staged per round
constexpr int kFinalImages = 128;            // images per round at most

struct TileLevels {
  int tile_off[B200DET_MAX_LEVELS + 1];
  int n_levels;
};

__global__ void __launch_bounds__(kFinalThreads)
finalize_losses_kernel(const int batch, const int n_tiles, const TileLevels tl, const float* __restrict__ partial,
                       const float* __restrict__ num_pos, const float* __restrict__ grad_box, const int grad_mode,
                       const float inv_batch, float* __restrict__ box_loss, float* __restrict__ cnt_loss, float* __restrict__ mean_out,
                       float* __restrict__ reg_scale_grad, const int use_pdl) {
  __shared__ float4 stage[kFinalStage];
  __shared__ float2 img[kFinalImages];
  __shared__ float img_dsc[kFinalImages][B200DET_MAX_LEVELS];    // per image, per level: scaled d/d scale
  if (use_pdl) pdl_wait();
  const int tid = threadIdx.x;
  const int per_round = min(kFinalImages, kFinalStage / n_tiles);   // images per round (n_tiles <= kFinalStage)
  float mb = 0.f, mc = 0.f, my_dsc = 0.f;
  for (int i0 = 0; i0 < batch; i0 += per_round) {
    const int n = min(per_round, batch - i0);
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(partial) + (size_t)i0 * n_tiles;
    for (int t = tid; t < n * n_tiles; t += kFinalThreads) stage[t] = __ldcg(src + t);
    __syncthreads();
    for (int i = tid; i < n; i += kFinalThreads) {
      const float np = __ldcg(num_pos + i0 + i);
      const float up = upstream_of(grad_box, grad_mode, i0 + i, inv_batch) / np;
      float tb = 0.f, tc = 0.f;
      for (int l = 0; l < tl.n_levels; ++l) {
        float td = 0.f;
        for (int t = tl.tile_off[l]; t < tl.tile_off[l + 1]; ++t) {
          const float4 v = stage[i * n_tiles + t];
          tb += v.x;
          tc += v.y;
          td += v.z;
        }
        img_dsc[i][l] = td * up;
      }
      tb /= np;
      tc /= np;
      box_loss[i0 + i] = tb;
      if (cnt_loss) cnt_loss[i0 + i] = tc;
      img[i] = make_float2(tb, tc);
    }
    __syncthreads();
    if (tid == 0)
      for (int i = 0; i < n; ++i) {
        mb += img[i].x;
        mc += img[i].y;
      }
    if (tid < tl.n_levels)                                     // thread l: level l's scale gradient, image order
      for (int i = 0; i < n; ++i) my_dsc += img_dsc[i][tid];
  }
  if (tid == 0 && mean_out) {
    mean_out[0] = mb / (float)batch;
    mean_out[1] = mc / (float)batch;
  }
  if (reg_scale_grad && tid < tl.n_levels) reg_scale_grad[tid] = my_dsc;
}

// In-place multiply of up to kMaxScaleMaps arrays, each by its own device scalar; a map whose scalar is
// exactly 1 costs nothing (the usual case: the fused kernel already applied d(mean)/d(loss[b]) = 1/B
// and backward() feeds 1).
constexpr int kMaxScaleMaps = 16;
struct ScaleTable {
  float* map[kMaxScaleMaps];
  const float* factor[kMaxScaleMaps];
  long long numel[kMaxScaleMaps];
};

__global__ void __launch_bounds__(256) scale_maps_kernel(const ScaleTable t) {
  const float f = *t.factor[blockIdx.y];
  if (f == 1.0f) return;
  float* m = t.map[blockIdx.y];
  const long long n = t.numel[blockIdx.y];
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) m[i] *= f;
}

// Autograd backward of the eager-gradient steps (the forward kernels wrote final gradients for an ASSUMED
// upstream gradient): map *= got / assumed, skipped entirely when the two are equal — which they are from the
// second step on, because the assumed value is then overwritten with the one that arrived (a GradScaler's loss
// scale is constant for thousands of steps).  ONE launch of a fixed, small grid (the usual case is the no-op):
// every CTA walks all maps with 128-bit accesses; the last CTA to finish (a ticket next to each assumed value)
// stores the new assumption, after every CTA has consumed the old one.
constexpr int kMaxStates = 4;
constexpr int kRescaleCtas = 74;             // the usual launch is a no-op: keep it small
struct RescaleTable {
  void* map[kMaxScaleMaps];
  long long numel[kMaxScaleMaps];
  int state_of[kMaxScaleMaps];
  const float* got[kMaxStates];
  float* state[kMaxStates];                  // {assumed upstream gradient, ticket (as bits of an unsigned)}
};

// map[i0 + k * step] *= f for one map, 128 bits per access where the map is 16-byte aligned
template <typename T> struct RescaleVec;
template <> struct RescaleVec<float> {
  static constexpr int kPer = 4;
  static __device__ __forceinline__ uint4 mul(uint4 v, const float f) {
    float4 x = *reinterpret_cast<float4*>(&v);
    x.x *= f; x.y *= f; x.z *= f; x.w *= f;
    return *reinterpret_cast<uint4*>(&x);
  }
  static __device__ __forceinline__ void mul1(float* p, const float f) { *p *= f; }
};
template <> struct RescaleVec<__half> {
  static constexpr int kPer = 8;
  static __device__ __forceinline__ uint32_t pair(uint32_t w, const float f) {
    const float2 x = __half22float2(*reinterpret_cast<__half2*>(&w));
    const __half2 r = __floats2half2_rn(x.x * f, x.y * f);
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  static __device__ __forceinline__ uint4 mul(uint4 v, const float f) {
    return make_uint4(pair(v.x, f), pair(v.y, f), pair(v.z, f), pair(v.w, f));
  }
  static __device__ __forceinline__ void mul1(__half* p, const float f) { *p = __float2half_rn(__half2float(*p) * f); }
};
template <> struct RescaleVec<__nv_bfloat16> {
  static constexpr int kPer = 8;
  static __device__ __forceinline__ uint32_t pair(uint32_t w, const float f) {
    const float2 x = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w));
    const __nv_bfloat162 r = __floats2bfloat162_rn(x.x * f, x.y * f);
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  static __device__ __forceinline__ uint4 mul(uint4 v, const float f) {
    return make_uint4(pair(v.x, f), pair(v.y, f), pair(v.z, f), pair(v.w, f));
  }
  static __device__ __forceinline__ void mul1(__nv_bfloat16* p, const float f) {
    *p = __float2bfloat16_rn(__bfloat162float(*p) * f);
  }
};

template <typename T>
__device__ __forceinline__ void rescale_map(T* m, const long long n, const float f, const long long i0, const long long step) {
  using V = RescaleVec<T>;
  if ((reinterpret_cast<uintptr_t>(m) & 15u) == 0) {
    uint4* m4 = reinterpret_cast<uint4*>(m);
    const long long n4 = n / V::kPer;
    long long j = i0;
    for (; j + 3 * step < n4; j += 4 * step) {                 // four 128-bit loads in flight per thread
      const uint4 v0 = m4[j], v1 = m4[j + step], v2 = m4[j + 2 * step], v3 = m4[j + 3 * step];
      m4[j] = V::mul(v0, f); m4[j + step] = V::mul(v1, f); m4[j + 2 * step] = V::mul(v2, f); m4[j + 3 * step] = V::mul(v3, f);
    }
    for (; j < n4; j += step) m4[j] = V::mul(m4[j], f);
    for (long long k = n4 * V::kPer + i0; k < n; k += step) V::mul1(m + k, f);
  } else {
    for (long long k = i0; k < n; k += step) V::mul1(m + k, f);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) rescale_maps_kernel(const RescaleTable t, const int n_maps, const int n_states) {
  // every thread reads the (<= 4) pairs itself: the usual launch finds them equal and ends here, without a
  // barrier or an atomic.  The assumed values only change after ALL CTAs have read them (ticket below), so
  // every CTA takes the same decision.
  float f[kMaxStates];
  bool any = false;
#pragma unroll
  for (int s = 0; s < kMaxStates; ++s) {
    f[s] = 1.0f;
    if (s < n_states) {
      const float got = __ldcg(t.got[s]), assumed = __ldcg(t.state[s]);
      if (got != assumed) {
        f[s] = got / assumed;
        any = true;
      }
    }
  }
  if (!any) return;
  const long long step = (long long)gridDim.x * 256;
  const long long i0 = (long long)blockIdx.x * 256 + threadIdx.x;
  for (int i = 0; i < n_maps; ++i) {
    const int so = t.state_of[i];
    const float fi = so == 0 ? f[0] : so == 1 ? f[1] : so == 2 ? f[2] : f[3];
    if (fi != 1.0f) rescale_map(static_cast<T*>(t.map[i]), t.numel[i], fi, i0, step);
  }
  // remember the upstream gradient that arrived (a zero / non-finite one is not a usable assumption for the
  // next forward: its gradients could not be rescaled); only states that differed take part
  __syncthreads();
  if (threadIdx.x < n_states) {
    const float g = __ldcg(t.got[threadIdx.x]);
    if (g != __ldcg(t.state[threadIdx.x])) {
      unsigned* ticket = reinterpret_cast<unsigned*>(t.state[threadIdx.x] + 1);
      __threadfence();
      if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
        if (g != 0.f && isfinite(g)) *t.state[threadIdx.x] = g;
        *ticket = 0u;
      }
    }
  }
}

}  // namespace
}  // namespace b200det

using namespace b200det;

namespace {
// workspace: [ticket] [tile partials B x tiles x 2]
size_t ticket_bytes() { return 256; }
int train_tiles(int num_points) { return (num_points + kTrainTile - 1) / kTrainTile + B200DET_MAX_LEVELS; }
}  // namespace

extern "C" size_t b200det_assign_loss_workspace_bytes(int batch, int num_points) {
  if (batch <= 0 || num_points <= 0) return 0;
  return ticket_bytes() + align_up((size_t)batch * train_tiles(num_points) * 4 * sizeof(float), 256);
}

extern "C" int b200det_assign_loss_fused(const b200det_level* levels, float* const* reg_grads, float* const* cnt_grads,
                                         int n_levels, const float* limit_lo, const float* limit_hi,
                                         const float* radius_px, int batch, int max_gt, const float* gt_boxes,
                                         const int64_t* gt_labels, int mode, const float* grad_box,
                                         const float* grad_cnt, int grad_mode, int64_t* cls_t, float* cnt_t, float* reg_t,
                                         float* box_loss, float* cnt_loss, float* num_pos, float* mean_out,
                                         float* reg_scale_grad, void* workspace, size_t workspace_bytes,
                                         void* stream) {
  if (!levels || n_levels <= 0 || n_levels > B200DET_MAX_LEVELS || !limit_lo || !limit_hi || !radius_px ||
      batch <= 0 || batch > 65535 || max_gt < 0 || !reg_grads || !cls_t || !cnt_t || !reg_t || !box_loss ||
      !num_pos || !workspace)
    return B200DET_ERR_ARG;
  if (max_gt > 0 && (!gt_boxes || !gt_labels)) return B200DET_ERR_ARG;
  if (!aligned16(gt_boxes) || !aligned16(reg_t) || !aligned16(workspace)) return B200DET_ERR_ARG;
  if ((mode != 0 && mode != 1) || (grad_mode != 0 && grad_mode != 1)) return B200DET_ERR_UNSUPPORTED;
  const bool has_cnt = cnt_grads != nullptr;
  if (has_cnt != (cnt_loss != nullptr)) return B200DET_ERR_ARG;
  int32_t level_hw[2 * B200DET_MAX_LEVELS], strides[B200DET_MAX_LEVELS];
  LossMaps lm = {};
  for (int l = 0; l < n_levels; ++l) {
    if (!levels[l].reg || !reg_grads[l] || (has_cnt && (!levels[l].cnt || !cnt_grads[l]))) return B200DET_ERR_ARG;
    level_hw[2 * l] = levels[l].h;
    level_hw[2 * l + 1] = levels[l].w;
    strides[l] = levels[l].stride;
    lm.reg[l] = static_cast<const float*>(levels[l].reg);
    lm.reg_scale[l] = static_cast<const float*>(levels[l].reg_scale);
    lm.cnt[l] = has_cnt ? static_cast<const float*>(levels[l].cnt) : nullptr;
    lm.greg[l] = reg_grads[l];
    lm.gcnt[l] = has_cnt ? cnt_grads[l] : nullptr;
  }
  const int tile_points = kTrainThreads * kTrainPts;
  AssignTable at;
  if (!make_assign_table(level_hw, strides, limit_lo, limit_hi, radius_px, n_levels, tile_points, &at))
    return B200DET_ERR_ARG;
  const int n_tiles = at.tile_off[B200DET_MAX_LEVELS];
  if (n_tiles > 65535 || n_tiles > train_tiles(at.num_points)) return B200DET_ERR_UNSUPPORTED;
  if (workspace_bytes < b200det_assign_loss_workspace_bytes(batch, at.num_points)) return B200DET_ERR_WORKSPACE;
  if (n_tiles > kFinalStage) return B200DET_ERR_UNSUPPORTED;           // finalize_losses_kernel stages whole images
  float* partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + ticket_bytes());

  const size_t smem_count = (size_t)max_gt * sizeof(GtEntry) + (size_t)((at.num_points + 31) / 32) * 4;
  const size_t smem_tile = (size_t)max_gt * (sizeof(GtEntry) + sizeof(int));
  if (smem_count > 200 * 1024 || smem_tile > 180 * 1024) return B200DET_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaSuccess;
  if (smem_count > 40 * 1024)
    e = cudaFuncSetAttribute(count_positives_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count);
  bool scale_exp = false;
  for (int l = 0; l < n_levels; ++l) scale_exp |= levels[l].reg_scale != nullptr;
  auto tile_kernel = scale_exp ? assign_loss_tile_kernel<true> : assign_loss_tile_kernel<false>;
  if (e == cudaSuccess && smem_tile > 20 * 1024)
    e = cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tile);
  if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }

  count_positives_kernel<<<dim3(kCountSlices, batch), kCountThreads, smem_count, st>>>(at, max_gt, gt_boxes, num_pos);
  int rc = check_launch();
  if (rc) return rc;

  // Programmatic dependent launch: the streaming kernel starts while count_positives_kernel runs and
  // waits (griddepcontrol.wait) only before it reads num_pos.  B200DET_NO_PDL=1 serialises the two.
  static const bool no_pdl = getenv("B200DET_NO_PDL") && getenv("B200DET_NO_PDL")[0] == '1';
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(batch, n_tiles);
  cfg.blockDim = dim3(kTrainThreads);
  cfg.dynamicSmemBytes = smem_tile;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  e = cudaLaunchKernelEx(&cfg, tile_kernel, at, lm, has_cnt ? 1 : 0, max_gt, gt_boxes,
                         reinterpret_cast<const long long*>(gt_labels), mode, grad_box, grad_cnt, grad_mode,
                         1.0f / (float)batch, (const float*)num_pos, reinterpret_cast<long long*>(cls_t), cnt_t, reg_t,
                         partial, no_pdl ? 0 : 1);
  if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
  rc = check_launch();
  if (rc) return rc;
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(kFinalThreads);
  cfg.dynamicSmemBytes = 0;
  TileLevels tl;
  for (int l = 0; l <= B200DET_MAX_LEVELS; ++l) tl.tile_off[l] = at.tile_off[l];
  tl.n_levels = n_levels;
  e = cudaLaunchKernelEx(&cfg, finalize_losses_kernel, batch, n_tiles, tl, (const float*)partial,
                         (const float*)num_pos, grad_box, grad_mode, 1.0f / (float)batch, box_loss, cnt_loss, mean_out,
                         reg_scale_grad, no_pdl ? 0 : 1);
  if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
  return check_launch();
}

extern "C" int b200det_scale_maps(float* const* maps, const int64_t* numel, const float* const* factors, int n_maps,
                                  void* stream) {
  if (!maps || !numel || !factors || n_maps <= 0 || n_maps > kMaxScaleMaps) return B200DET_ERR_ARG;
  ScaleTable t = {};
  for (int i = 0; i < n_maps; ++i) {
    if (!maps[i] || !factors[i] || numel[i] < 0) return B200DET_ERR_ARG;
    t.map[i] = maps[i];
    t.factor[i] = factors[i];
    t.numel[i] = numel[i];
  }
  scale_maps_kernel<<<dim3(74, n_maps), 256, 0, static_cast<cudaStream_t>(stream)>>>(t);
  return check_launch();
}

extern "C" int b200det_rescale_maps(void* const* maps, const int64_t* numel, const int32_t* state_of, int dtype,
                                    int n_maps, const float* const* got, float* const* state, int n_states,
                                    void* stream) {
  if (!maps || !numel || !state_of || !got || !state || n_maps <= 0 || n_maps > kMaxScaleMaps || n_states <= 0 ||
      n_states > kMaxStates)
    return B200DET_ERR_ARG;
  if (dtype != B200DET_F32 && dtype != B200DET_F16 && dtype != B200DET_BF16) return B200DET_ERR_UNSUPPORTED;
  RescaleTable t = {};
  for (int i = 0; i < n_maps; ++i) {
    if (!maps[i] || numel[i] < 0 || state_of[i] < 0 || state_of[i] >= n_states) return B200DET_ERR_ARG;
    t.map[i] = maps[i];
    t.numel[i] = numel[i];
    t.state_of[i] = state_of[i];
  }
  for (int s = 0; s < n_states; ++s) {
    if (!got[s] || !state[s]) return B200DET_ERR_ARG;
    for (int r = 0; r < s; ++r)
      if (state[r] == state[s]) return B200DET_ERR_ARG;        // one ticket per state: list each state once
    t.got[s] = got[s];
    t.state[s] = state[s];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200DET_F32) rescale_maps_kernel<float><<<kRescaleCtas, 256, 0, st>>>(t, n_maps, n_states);
  else if (dtype == B200DET_F16) rescale_maps_kernel<__half><<<kRescaleCtas, 256, 0, st>>>(t, n_maps, n_states);
  else rescale_maps_kernel<__nv_bfloat16><<<kRescaleCtas, 256, 0, st>>>(t, n_maps, n_states);
  return check_launch();
}
