// The streaming kernel behind b200det_assign_targets (kLoss = false) and b200det_assign_loss_fused (kLoss = true).
//
// FCOSGenTargets.forward (model/modules/head.py:218-316) writes three dense target maps of which ~1 % of the
// points (the positives) carry information; compute_reg_loss / compute_cnt_loss (model/loss.py:29-57,116-177) and
// their backward add two dense gradient maps that are zero off the positives.  So the dense part of the step does
// not depend on the ground truth at all, and the CTA (256 threads, one tile of one level of one image) splits in two:
//   FILL WARP (warp 7)   writes the tile as if every point were a negative — cls_t = 0, cnt_t = -1, reg_t = -1,
//                        gradients = 0: 48 B per point, nothing read — with BULK COPIES from two constant
//                        shared-memory buffers (cp.async.bulk shared -> global, the TMA engine), one segment per
//                        lane, and waits for their completion.
//   CHAIN (warps 0-6)    loads the image's boxes, runs the box-centric vote of assign_body.cuh into shared memory,
//                        compacts the tile's positives into an ORDERED list (the loss sum must not depend on the
//                        order of atomics), and (kLoss) fetches the predictions of the positives and evaluates their
//                        loss terms and unscaled gradients.  Its barriers are named barriers of 224 threads.
//   join (__syncthreads) then the chain PATCHES the positives on top of the fill: targets and scaled gradients.
// Why the split: a CTA barrier also waits for the warp's outstanding global stores, so with plain stores the first
// barrier of the vote sits out the whole fill (the first version therefore wrote its dense stream LAST and exposed
// the ground-truth latency and the vote instead), and a thread issuing bulk copies blocks while the copy queue is
// full.  This way the ~3-5 us the memory system needs for the fill hide everything else.
// The gradient of a positive carries 1 / num_pos of its IMAGE.  Each CTA adds (1 << 32 | positives of its tile) to
// a 64-bit counter of its image right after the vote; a CTA that owns positives then reads num_pos from the counter's
// low word once all tiles of the image have arrived.  The grid is image-major, so these are neighbouring CTAs that
// run the same schedule; the wait is BOUNDED (a few microseconds), after which the CTA recounts the image's
// positives itself (all boxes x levels x window points into a bitmap in shared memory) — no unbounded spinning, no
// co-residency assumption, no count kernel in front.
// finalize_losses_kernel (train_fused.cu; one CTA, a programmatic dependent: it is resident before this grid ends) adds
// the tile partials in tile order and the per-image losses in image order (deterministic), publishes num_pos[] and
// the batch means and clears the counters for the next call.  (Folding it into this kernel — a ticket, the last CTA
// or the last tile of every image reducing — was built and measured twice: every variant is a chain of 4-6 global
// round trips of ~0.7 us behind the last tile, 3.3 us of tail against 3.3 us for the dependent kernel with its launch
// boundary, and costs every CTA a fence + atomic round trip on top.)
// The kernel may be launched as a programmatic dependent of whatever precedes it: it touches global memory only
// after griddepcontrol.wait.
#pragma once
#include "assign_body.cuh"
#include "block_utils.cuh"
#include "loss_terms.cuh"
#include "tma.cuh"

namespace b200det {

// bulk copy shared -> global (dst, src 16-byte aligned, bytes a multiple of 16), tracked by the thread's bulk group
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct LossMaps {
  const float* reg[B200DET_MAX_LEVELS];
  const float* reg_scale[B200DET_MAX_LEVELS];   // ScaleExp folded in (common.cuh); NULL = reg holds the distances
  const float* cnt[B200DET_MAX_LEVELS];
  float* greg[B200DET_MAX_LEVELS];
  float* gcnt[B200DET_MAX_LEVELS];
};

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

constexpr int kStreamThreads = 256;
constexpr int kFillWarp = 7;                   // the last warp fills, warps 0-6 are the chain
constexpr int kChain = kFillWarp * 32;         // 224 chain threads
constexpr int kArrivalPolls = 32;              // x (load + 64 ns sleep): a few us at most before the local recount
constexpr int kStageRegs = 2;                  // ground-truth boxes a chain thread keeps in registers
constexpr int kFillChunk = 8192;               // bytes of each constant buffer = largest single bulk copy

__device__ __forceinline__ void chain_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kChain) : "memory"); }

// Exclusive scan of one int per chain thread (7 warps); `sums` is 9 ints of shared scratch.
__device__ __forceinline__ int chain_exclusive_scan(int v, int* sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) sums[warp] = incl;
  chain_sync();
  int before = 0, all = 0;
#pragma unroll
  for (int i = 0; i < kFillWarp; ++i) {
    const int s = sums[i];
    before += i < warp ? s : 0;
    all += s;
  }
  *total = all;
  return before + incl - v;
}

// n floats at p (4-byte aligned) <- v, by ONE thread: the 16-byte aligned body as bulk copies from `src`
// (kFillChunk bytes of the constant in shared memory), the <= 3 floats before / after it as plain stores.
__device__ __forceinline__ void lane_fill_f32(float* __restrict__ p, const int n, const float v, const float* src) {
  const int head = min(n, (int)(((16u - (unsigned)(reinterpret_cast<uintptr_t>(p) & 15u)) & 15u) >> 2));
  const int body = (n - head) & ~3;
  char* dst = reinterpret_cast<char*>(p + head);
  for (int off = 0; off < body * 4; off += kFillChunk) bulk_s2g(dst + off, src, (uint32_t)min(kFillChunk, body * 4 - off));
  for (int i = 0; i < head; ++i) p[i] = v;
  for (int i = head + body; i < n; ++i) p[i] = v;
}

struct StreamArgs {
  int has_cnt, M, mode, grad_mode, n_tiles;
  int arrival_polls;                  // bounded wait for num_pos (kArrivalPolls; 0 = always recount: tests)
  float inv_batch;
  const float* gt_boxes;
  const long long* gt_labels;
  const float* grad_box;
  const float* grad_cnt;
  unsigned long long* counters;       // [B] (arrivals << 32) | positives, zero on entry (kLoss)
  long long* cls_t;
  float* cnt_t;
  float* reg_t;
  int32_t* gt_index;                  // optional (assign only)
  float* partial;                     // [B, n_tiles, 4] (kLoss)
  int batch;
};

template <int kPts, bool kLoss, bool kScaleExp>
__global__ void __launch_bounds__(kStreamThreads, 4)
assign_stream_kernel(const AssignTable at, const LossMaps lm, const StreamArgs a) {
  constexpr int kTilePoints = kStreamThreads * kPts;                     // points of one level of one image per CTA
  constexpr int kRun = (kTilePoints + kChain - 1) / kChain;              // consecutive points a chain thread compacts
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = a.M;
  GtEntry* gts = reinterpret_cast<GtEntry*>(smem_raw);                   // [M] the image's boxes, by GT index
  int* cand = reinterpret_cast<int*>(smem_raw + (size_t)M * sizeof(GtEntry));   // [M] GT indices relevant to this tile
  int* cand_col = cand + M;                                              // [M] their columns in reach: first | count << 16
  int* cand_row = cand_col + M;                                          // [M] their rows in reach
  unsigned* bitmap = reinterpret_cast<unsigned*>(cand_row + M);          // [ceil(P / 32)] recount fallback (kLoss)
  __shared__ unsigned long long keys[kTilePoints];                       // per point: (area bits << 32) | GT index
  __shared__ unsigned short plist[kTilePoints];                          // the tile's positives, ascending
  __shared__ __align__(16) float c_neg[kFillChunk / 4];                  // -1.0f ...   (sources of the bulk fill)
  __shared__ __align__(16) float c_zero[kFillChunk / 4];                 //  0.0f ...
  __shared__ int s_scan[9];
  __shared__ float s_red[32];
  __shared__ int s_n, s_np, s_wv;

  // grid = (tile, image): the tiles of an image are neighbours (they meet in the image's counter); within an image
  // the coarse levels, which carry the most vote work, come first
  const int b = blockIdx.y;
  const int n_tiles = (int)gridDim.x;
  const int tile = n_tiles - 1 - (int)blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int l = 0;
#pragma unroll
  for (int i = 1; i < B200DET_MAX_LEVELS; ++i) l += (i < at.n_levels && tile >= at.tile_off[i]) ? 1 : 0;
  const int hw = at.hw[l], w = at.w[l], h = at.h[l], s = at.stride[l];
  const int t0 = (tile - at.tile_off[l]) * kTilePoints;
  const int n = min(kTilePoints, hw - t0);                               // points of this tile
  const float lo = at.lo[l], hi = at.hi[l], radius = at.radius[l];
  const size_t out0 = (size_t)b * at.num_points + at.point_off[l] + t0;  // first point of the tile in [B, P]
  pdl_launch_dependents();                       // the next kernel of the stream may become resident (it waits itself)
  const bool traced = b == 0 && (tile == 0 || tile == n_tiles - 1);
  const int tslot = tile == 0 ? 0 : 16;
  B200DET_STAMP_IF(traced, tslot + 0);
  B200DET_SPAN_BEGIN();

  float* __restrict__ greg = kLoss ? lm.greg[l] + (size_t)b * 4 * hw + t0 : nullptr;
  float* __restrict__ gcnt = (kLoss && a.has_cnt) ? lm.gcnt[l] + (size_t)b * hw + t0 : nullptr;
  const float* __restrict__ reg_p = kLoss ? lm.reg[l] + (size_t)b * 4 * hw + t0 : nullptr;
  const float* __restrict__ cnt_p = (kLoss && a.has_cnt) ? lm.cnt[l] + (size_t)b * hw + t0 : nullptr;
  const bool scaled = kLoss && kScaleExp && lm.reg_scale[l] != nullptr;

  // one positive (chain threads): the reference's targets for local point `local`
  auto targets_of = [&](const int local, float4* tg, float* ct, int* label, int* idx) {
    const GtEntry g = gts[(unsigned)(keys[local] & 0xffffffffull)];
    const int pos = t0 + local;
    const int row = pos / w, col = pos - row * w;
    positive_targets(g, col, row, s, tg, ct);
    *label = g.label;
    *idx = g.idx;
  };

  constexpr int kHeld = 2;                       // positives per chain thread whose gradients stay in registers
  float4 gr[kHeld];
  float gc[kHeld];
  float acc_box = 0.f, acc_cnt = 0.f, acc_dsc = 0.f;
  float sc = 0.f;
  int n_pos = 0;
  float scale_box = 0.f, scale_cnt = 0.f;

  // loss terms and unscaled gradients of one positive; adds the terms to this thread's partials
  auto loss_of = [&](const int local, float4* g_reg, float* g_cnt) {
    float4 tg;
    float ct;
    int label, idx;
    const float4 x4 = make_float4(reg_p[local], reg_p[hw + local], reg_p[2 * hw + local], reg_p[3 * hw + local]);
    const float px = a.has_cnt ? cnt_p[local] : 0.f;
    targets_of(local, &tg, &ct, &label, &idx);
    float4 gg;
    if (scaled) {
      // raw regression output x: distances d = exp(x * sc) (ScaleExp); dL/dx = dL/dd * d * sc, dL/dsc += dL/dd * d * x
      const float4 d4 = make_float4(scale_exp_f32(x4.x, sc), scale_exp_f32(x4.y, sc), scale_exp_f32(x4.z, sc),
                                    scale_exp_f32(x4.w, sc));
      acc_box += box_term<true>(d4, tg, a.mode, &gg);
      gg = make_float4(gg.x * d4.x, gg.y * d4.y, gg.z * d4.z, gg.w * d4.w);
      acc_dsc += (gg.x * x4.x + gg.y * x4.y) + (gg.z * x4.z + gg.w * x4.w);
      gg = make_float4(gg.x * sc, gg.y * sc, gg.z * sc, gg.w * sc);
    } else {
      acc_box += box_term<true>(x4, tg, a.mode, &gg);
    }
    *g_reg = gg;
    if (a.has_cnt) {
      acc_cnt += bce_term(px, ct);
      *g_cnt = sigmoid_f32(px) - ct;
    }
  };

  // shared-memory set-up by all threads: the two constants, the empty vote
  for (int i = tid; i < kFillChunk / 16; i += kStreamThreads) {
    reinterpret_cast<float4*>(c_neg)[i] = make_float4(-1.f, -1.f, -1.f, -1.f);
    reinterpret_cast<float4*>(c_zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int q = 0; q < kPts; ++q) keys[tid + q * kStreamThreads] = kNoWinner;
  if (tid == 0) {
    s_n = 0;
    s_wv = 1;
  }
  fence_proxy_async_smem();                      // the constants are visible to the copy engine
  __syncthreads();
  pdl_wait();                                    // the previous kernel of the stream is complete: global memory may be touched
  B200DET_STAMP_IF(traced, tslot + 1);

  if (warp == kFillWarp) {
    // ================= FILL WARP ============================================================================
    B200DET_STAMP_ANY(traced && lane == 0, tslot + 10);
    // one segment per lane: reg_t is half of the target bytes and goes out in four parts
    if (lane < 4) lane_fill_f32(a.reg_t + 4 * out0 + lane * n, n, -1.f, c_neg);
    if (lane == 4) lane_fill_f32(reinterpret_cast<float*>(a.cls_t + out0), 2 * n, 0.f, c_zero);   // int64 zeros
    if (lane == 5) lane_fill_f32(a.cnt_t + out0, n, -1.f, c_neg);
    if (kLoss) {
      if (lane >= 8 && lane < 12) lane_fill_f32(greg + (size_t)(lane - 8) * hw, n, 0.f, c_zero);
      if (lane == 12 && a.has_cnt) lane_fill_f32(gcnt, n, 0.f, c_zero);
    }
    if (a.gt_index)                              // (test / debug output: plain stores)
      for (int i = lane; i < n; i += 32) a.gt_index[out0 + i] = -1;
    bulk_commit();
    B200DET_STAMP_ANY(traced && lane == 0, tslot + 11);
    bulk_wait_all();                             // the fill has landed: the patches go on top of it
    B200DET_STAMP_ANY(traced && lane == 0, tslot + 9);
  } else {
    // ================= CHAIN ================================================================================
    const float4* g4 = reinterpret_cast<const float4*>(a.gt_boxes) + (size_t)b * M;
    const long long* lab = a.gt_labels + (size_t)b * M;
    float4 gbox[kStageRegs];
    long long glab[kStageRegs];
    const bool staged_in_regs = M <= kStageRegs * kChain;
    if (staged_in_regs) {
#pragma unroll
      for (int k = 0; k < kStageRegs; ++k) {
        const int m = tid + k * kChain;
        if (m < M) {
          gbox[k] = __ldg(g4 + m);
          glab[k] = __ldg(lab + m);
        }
      }
    }
    if (scaled) sc = __ldg(lm.reg_scale[l]);
    // Stage the boxes and keep those that can be positive somewhere in this tile, each with the columns and rows
    // of the level that pass the reference's centre test for it (head.py:275-283: |x - cx| < radius and
    // |y - cy| < radius, the same fp32 subtraction): at most ceil(2 radius / stride) of each, so the vote below
    // evaluates ~9 points per box instead of a 5 x 5 window with slack, and needs no centre test of its own.
    const int row_first = t0 / w, row_last = (t0 + n - 1) / w;
    const int hwin = window_half(radius, s);
    const float inv_s = 1.0f / (float)s;
    const int half = s / 2;
    auto stage = [&](const GtEntry& g, const int m) {
      gts[m] = g;
      if (!gt_may_hit(g, row_first, row_last, s, lo, hi, radius)) return;
      const int cj = (int)floorf(g.cx * inv_s), ci = (int)floorf(g.cy * inv_s);
      int j0 = 0, nj = 0, i0 = 0, ni = 0;
      for (int c = -hwin; c <= hwin; ++c) {
        const int j = cj + c, i = ci + c;
        if (j >= 0 && j < w && fabsf(__fsub_rn((float)(j * s + half), g.cx)) < radius) {
          if (!nj) j0 = j;
          ++nj;
        }
        if (i >= row_first && i <= row_last && fabsf(__fsub_rn((float)(i * s + half), g.cy)) < radius) {
          if (!ni) i0 = i;
          ++ni;
        }
      }
      if (nj == 0 || ni == 0) return;
      const int slot = atomicAdd(&s_n, 1);
      cand[slot] = m;
      cand_col[slot] = j0 | (nj << 16);
      cand_row[slot] = i0 | (ni << 16);
      atomicMax(&s_wv, max(nj, ni));
    };
    if (staged_in_regs) {
#pragma unroll
      for (int k = 0; k < kStageRegs; ++k) {
        const int m = tid + k * kChain;
        if (m < M) stage(make_gt_entry(gbox[k], m, (int)glab[k]), m);
      }
    } else {
      for (int m = tid; m < M; m += kChain) stage(make_gt_entry(g4[m], m, (int)lab[m]), m);
    }
    chain_sync();
    const int n_list = s_n;
    B200DET_STAMP_IF(traced, tslot + 2);
    // box-centric vote: every (box, column in reach, row in reach) evaluates the reference's exact fp32 expressions;
    // positives race with a 64-bit atomicMin on (area bits, GT index): smallest area, lowest index on ties
    // (torch.min's first index on the masked areas, head.py:285-286)
    const int wv = s_wv, kk = wv * wv;
    for (int pi = tid; pi < n_list * kk; pi += kChain) {
      const int e = pi / kk, k = pi - e * kk;
      const int kr = k / wv, kc = k - kr * wv;
      const int col = cand_col[e], row = cand_row[e];
      if (kc >= (col >> 16) || kr >= (row >> 16)) continue;
      const int j = (col & 0xffff) + kc, i = (row & 0xffff) + kr;
      const int pos = i * w + j;
      if (pos < t0 || pos >= t0 + n) continue;
      const GtEntry g = gts[cand[e]];
      const float x = (float)(j * s + half), y = (float)(i * s + half);
      const float lf = __fsub_rn(x, g.x0), tf = __fsub_rn(y, g.y0);
      const float rf = __fsub_rn(g.x1, x), bf = __fsub_rn(g.y1, y);
      const float omin = fminf(fminf(lf, tf), fminf(rf, bf));
      const float omax = fmaxf(fmaxf(lf, tf), fmaxf(rf, bf));
      if (!((omin > 0.f) && (omax > lo) && (omax <= hi))) continue;
      const float area = __fmul_rn(__fadd_rn(lf, rf), __fadd_rn(tf, bf));
      atomicMin(&keys[pos - t0], ((unsigned long long)__float_as_uint(area) << 32) | (unsigned)g.idx);
    }
    chain_sync();
    B200DET_STAMP_IF(traced, tslot + 3);

    // ordered list of the tile's positives: chain thread t looks at local points [t * kRun, (t + 1) * kRun)
    const int r0 = tid * kRun;
    int mine = 0;
#pragma unroll
    for (int q = 0; q < kRun; ++q) mine += (r0 + q < kTilePoints && keys[r0 + q] != kNoWinner) ? 1 : 0;
    int at_ = chain_exclusive_scan(mine, s_scan, &n_pos);
    if (mine) {
#pragma unroll
      for (int q = 0; q < kRun; ++q)
        if (r0 + q < kTilePoints && keys[r0 + q] != kNoWinner) plist[at_++] = (unsigned short)(r0 + q);
    }
    if (kLoss && tid == 0) atomicAdd(a.counters + b, (1ull << 32) | (unsigned long long)n_pos);   // this tile has arrived
    chain_sync();
    B200DET_STAMP_IF(traced, tslot + 4);
    B200DET_NOTE_IF(traced, tslot + 12, n_pos);
    B200DET_NOTE_IF(traced, tslot + 14, n_list);

    if constexpr (kLoss) {
      // predictions of the positives, loss terms, unscaled gradients: the gradients of up to two positives per
      // thread stay in registers, the (rare) rest of a dense tile is evaluated again after the join
#pragma unroll
      for (int k = 0; k < kHeld; ++k) {
        const int e = tid + k * kChain;
        gr[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        gc[k] = 0.f;
        if (e < n_pos) loss_of(plist[e], &gr[k], &gc[k]);
      }
      B200DET_STAMP_IF(traced, tslot + 5);
      // num_pos of the image (only the gradients of positives need it): a BOUNDED wait for the other tiles
      if (n_pos > 0) {
        if (tid == 0) {
          unsigned long long v = 0ull;
          for (int spin = 0; spin < a.arrival_polls; ++spin) {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a.counters + b) : "memory");
            if ((int)(v >> 32) >= n_tiles) break;
            __nanosleep(64);
          }
          s_np = ((int)(v >> 32) >= n_tiles) ? (int)(v & 0xffffffffull) : -1;
        }
        chain_sync();
        int np_i = s_np;
        B200DET_NOTE_IF(traced, tslot + 13, np_i);
        if (np_i < 0) {
          // the other tiles are late (a grid of several waves): local recount — every (box, level) pair that passes
          // the level / padding filter, lanes = window points, into a bitmap
          const int words = (at.num_points + 31) / 32;
          for (int i = tid; i < words; i += kChain) bitmap[i] = 0u;
          chain_sync();
          for (int pair = warp; pair < M * at.n_levels; pair += kFillWarp) {
            const int pl = pair / M, m = pair - pl * M;
            const GtEntry g = gts[m];
            const float side = fmaxf(g.x1 - g.x0, g.y1 - g.y0);
            if (!(side > 0.f && side > at.lo[pl] - 1.0f && 0.5f * side <= at.hi[pl] + 1.0f)) continue;
            const int ps = at.stride[pl];
            const float pr = at.radius[pl];
            const int phw = window_half(pr, ps);
            const int pcount = (2 * phw + 1) * (2 * phw + 1);
            for (int k = lane; k < pcount; k += 32) {
              int pos;
              float area;
              if (window_point_positive(g, k, phw, ps, at.w[pl], at.h[pl], at.lo[pl], at.hi[pl], pr, &pos, &area)) {
                const int p = at.point_off[pl] + pos;
                atomicOr(bitmap + (p >> 5), 1u << (p & 31));
              }
            }
          }
          chain_sync();
          int c = 0;
          for (int i = tid; i < words; i += kChain) c += __popc(bitmap[i]);
          c = __reduce_add_sync(0xffffffffu, c);
          if (lane == 0) s_scan[warp] = c;
          chain_sync();
          np_i = 0;
#pragma unroll
          for (int i = 0; i < kFillWarp; ++i) np_i += s_scan[i];
        }
        const float np = fmaxf((float)np_i, 1.f);
        scale_box = upstream_of(a.grad_box, a.grad_mode, b, a.inv_batch) / np;
        scale_cnt = a.has_cnt ? upstream_of(a.grad_cnt, a.grad_mode, b, a.inv_batch) / np : 0.f;
      }
      B200DET_STAMP_IF(traced, tslot + 6);
    }
  }

  __syncthreads();                               // join: the fill is complete, the patches may be stored
  B200DET_STAMP_IF(traced, tslot + 7);
  if (warp != kFillWarp) {
    auto store_targets = [&](const int local) {
      float4 tg;
      float ct;
      int label, idx;
      targets_of(local, &tg, &ct, &label, &idx);
      const size_t o = out0 + local;
      a.cls_t[o] = (long long)label;
      a.cnt_t[o] = ct;
      *reinterpret_cast<float4*>(a.reg_t + 4 * o) = tg;
      if (a.gt_index) a.gt_index[o] = idx;
    };
    auto store_grads = [&](const int local, const float4 g_reg, const float g_cnt) {
      greg[local] = g_reg.x * scale_box;
      greg[hw + local] = g_reg.y * scale_box;
      greg[2 * hw + local] = g_reg.z * scale_box;
      greg[3 * hw + local] = g_reg.w * scale_box;
      if (a.has_cnt) gcnt[local] = g_cnt * scale_cnt;
    };
    for (int e = tid; e < n_pos; e += kChain) store_targets(plist[e]);
    if constexpr (kLoss) {
#pragma unroll
      for (int k = 0; k < kHeld; ++k) {
        const int e = tid + k * kChain;
        if (e < n_pos) store_grads(plist[e], gr[k], gc[k]);
      }
      for (int e = tid + kHeld * kChain; e < n_pos; e += kChain) {       // dense tiles only
        float4 g_reg = make_float4(0.f, 0.f, 0.f, 0.f);
        float g_cnt = 0.f;
        loss_of(plist[e], &g_reg, &g_cnt);
        store_grads(plist[e], g_reg, g_cnt);
      }
      // this tile's loss partials: fixed shuffle tree per warp, then the warps in order
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        acc_box += __shfl_xor_sync(0xffffffffu, acc_box, d);
        acc_cnt += __shfl_xor_sync(0xffffffffu, acc_cnt, d);
        acc_dsc += __shfl_xor_sync(0xffffffffu, acc_dsc, d);
      }
      if (lane == 0) {
        s_red[3 * warp] = acc_box;
        s_red[3 * warp + 1] = acc_cnt;
        s_red[3 * warp + 2] = acc_dsc;
      }
      chain_sync();
      if (tid == 0) {
        float tb = 0.f, tc = 0.f, td = 0.f;
#pragma unroll
        for (int wi = 0; wi < kFillWarp; ++wi) {
          tb += s_red[3 * wi];
          tc += s_red[3 * wi + 1];
          td += s_red[3 * wi + 2];
        }
        *reinterpret_cast<float4*>(a.partial + ((size_t)b * n_tiles + tile) * 4) = make_float4(tb, tc, td, 0.f);
      }
    }
  }
  B200DET_STAMP_IF(traced, tslot + 8);
  B200DET_SPAN_END();
}

// Tile shape against WAVE QUANTISATION (small problems are a handful of waves): points per thread in {4, 6, 8} that
// minimises waves * (fixed CTA latency + per-point time).  Large problems take 8.
inline int stream_points_per_thread(const int32_t* level_hw, int n_levels, int batch) {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      sm_count = 148;
  }
  const long long slots = (long long)sm_count * 4;                       // 256 threads, <= 64 registers: 4 CTAs per SM
  double best = 1e30;
  int pts = 8;
  for (int c = 4; c <= 8; c += 2) {
    long long tiles = 0;
    for (int l = 0; l < n_levels; ++l)
      tiles += ((long long)level_hw[2 * l] * level_hw[2 * l + 1] + kStreamThreads * c - 1) / (kStreamThreads * c);
    const long long waves = (tiles * batch + slots - 1) / slots;
    const double cost = (double)waves * (1.2 + 0.21 * c);
    if (cost < best) {
      best = cost;
      pts = c;
    }
  }
  return pts;
}

// Launch as a programmatic dependent of the stream's previous kernel (the kernel waits itself before it touches
// global memory); B200DET_NO_PDL=1 switches the attribute off.
template <typename Kernel>
inline cudaError_t launch_stream_kernel(Kernel kernel, dim3 grid, size_t smem, cudaStream_t st, const AssignTable& at,
                                        const LossMaps& lm, const StreamArgs& a, bool pdl) {
  if (smem > 8 * 1024) {                         // static shared memory (keys, list, constants) counts against 48 KB
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kStreamThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, at, lm, a);
}

}  // namespace b200det
