// K4 fused — target assignment + box (IoU / GIoU) loss + centerness BCE loss, forward AND backward,
// in ONE launch.  Replaces, for one training step,
//   FCOSGenTargets.forward            model/modules/head.py:218-316
//   compute_reg_loss (+ giou / iou)   model/loss.py:116-177      and its autograd backward
//   compute_cnt_loss                  model/loss.py:29-57        and its autograd backward
//   the two `.mean()` of FCOSLoss     model/loss.py:210-213
//
// One thread-block CLUSTER per image; CTA `rank` owns a contiguous slice of the image's level-major
// points.  Phases (all data dependent state stays in shared memory, nothing is re-read from HBM):
//   1. stage the image's GT boxes; list the (box, level) pairs that can be positive in the slice;
//   2. box-centric vote (assign_body.cuh): per point a 64-bit atomicMin on (area, GT index);
//   3. count the slice's positives, exchange the counts through distributed shared memory,
//      cluster barrier -> every CTA knows num_pos of the image, hence the gradient scale;
//   4. stream pass A: every point's targets (28 B) and the zero gradients of negatives (20 B) — the
//      write stream starts here and keeps HBM busy while
//      pass B fetches the predictions at the positives only, evaluates the loss terms and writes
//      their gradients, already scaled by grad_loss[b] / num_pos[b];
//   5. loss partials meet in CTA 0 through DSMEM and are added in rank order (deterministic);
//      the last cluster to finish (self-resetting ticket) adds the per-image losses in image order
//      and writes the two batch means.
// HBM traffic: 48 B written per point + ~36 B read per positive; GT boxes are read once per CTA.
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>

#include "assign_body.cuh"
#include "loss_terms.cuh"

namespace cg = cooperative_groups;

namespace b200det {
namespace {

struct FusedTable {
  const float* reg[B200DET_MAX_LEVELS];
  const float* cnt[B200DET_MAX_LEVELS];
  float* greg[B200DET_MAX_LEVELS];
  float* gcnt[B200DET_MAX_LEVELS];
  int h[B200DET_MAX_LEVELS], w[B200DET_MAX_LEVELS], stride[B200DET_MAX_LEVELS], hw[B200DET_MAX_LEVELS];
  int point_off[B200DET_MAX_LEVELS + 1];
  float lo[B200DET_MAX_LEVELS], hi[B200DET_MAX_LEVELS], radius[B200DET_MAX_LEVELS];
  int n_levels, num_points;
  int chunk;             // points per CTA = ceil(P / cluster size)
  int has_cnt;           // centerness maps + gradients present
};

__device__ __forceinline__ int level_of(const FusedTable& t, const int p) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < B200DET_MAX_LEVELS; ++i) l += (i < t.n_levels && p >= t.point_off[i]) ? 1 : 0;
  return l;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

constexpr int kMaxCluster = 16;

template <int kThreads>
__global__ void __launch_bounds__(kThreads)
assign_loss_fused_kernel(const FusedTable ft, const int M, const float* __restrict__ gt_boxes,
                         const long long* __restrict__ gt_labels, const int mode,
                         const float* __restrict__ grad_box, const float* __restrict__ grad_cnt,
                         const float inv_batch, long long* __restrict__ cls_t, float* __restrict__ cnt_t,
                         float* __restrict__ reg_t, float* __restrict__ box_loss, float* __restrict__ cnt_loss,
                         float* __restrict__ num_pos, float* __restrict__ mean_out, unsigned* __restrict__ ticket) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);          // [chunk] per own point
  GtEntry* gts = reinterpret_cast<GtEntry*>(smem_raw + (size_t)ft.chunk * 8);          // [M] by GT index
  int* cand = reinterpret_cast<int*>(gts + M);                                         // [M * levels] (level << 24) | m
  __shared__ float s_red[32];
  __shared__ int s_count[kMaxCluster];          // positives per CTA of the cluster (every CTA holds a copy)
  __shared__ float s_part[2 * kMaxCluster];     // loss partials (meaningful in CTA 0)
  __shared__ int s_n;

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int csize = (int)cluster.num_blocks();
  const int b = blockIdx.y;
  const int P = ft.num_points;
  const int p_lo = min(P, rank * ft.chunk), p_hi = min(P, p_lo + ft.chunk);
  const int n_own = p_hi - p_lo;
  const int tid = threadIdx.x;

  cluster.barrier_arrive();      // matched by barrier_wait() before the first distributed-shared-memory store:
                                 // by then every CTA of the cluster has started

  // ---- 1. stage the boxes, list the (box, level) pairs relevant to the slice -----------------------
  if (tid == 0) s_n = 0;
  for (int i = tid; i < n_own; i += kThreads) keys[i] = kNoWinner;
  {
    const float4* g4 = reinterpret_cast<const float4*>(gt_boxes) + (size_t)b * M;
    const long long* lab = gt_labels + (size_t)b * M;
    for (int m = tid; m < M; m += kThreads) gts[m] = make_gt_entry(g4[m], m, (int)lab[m]);
  }
  __syncthreads();
  const int l_first = n_own > 0 ? level_of(ft, p_lo) : 0;
  const int l_last = n_own > 0 ? level_of(ft, p_hi - 1) : -1;
  int wmax = 1;                                   // window points per box, max over the slice's levels
  for (int l = l_first; l <= l_last; ++l) {
    const int side = 2 * window_half(ft.radius[l], ft.stride[l]) + 1;
    wmax = max(wmax, side * side);
  }
  for (int i = tid; i < M * (l_last - l_first + 1); i += kThreads) {
    const int l = l_first + i / M, m = i - (l - l_first) * M;
    const int t0 = max(p_lo, ft.point_off[l]) - ft.point_off[l];
    const int t1 = min(p_hi, ft.point_off[l + 1]) - 1 - ft.point_off[l];
    const int w = ft.w[l];
    if (gt_may_hit(gts[m], t0 / w, t1 / w, ft.stride[l], ft.lo[l], ft.hi[l], ft.radius[l]))
      cand[atomicAdd(&s_n, 1)] = (l << 24) | m;
  }
  __syncthreads();
  const int n_list = s_n;

  // ---- 2. box-centric vote ------------------------------------------------------------------------
  for (int pi = tid; pi < n_list * wmax; pi += kThreads) {
    const int e = pi / wmax, k = pi - e * wmax;
    const int c = cand[e];
    const int l = c >> 24, m = c & 0xffffff;
    const int s = ft.stride[l];
    const int hwin = window_half(ft.radius[l], s);
    if (k >= (2 * hwin + 1) * (2 * hwin + 1)) continue;
    const int off = ft.point_off[l];
    const int t0 = max(p_lo, off) - off;
    const int t1 = min(p_hi, ft.point_off[l + 1]) - 1 - off;
    // keys is indexed by the slice-local point index: (off + pos) - p_lo = pos - t0 + (off + t0 - p_lo)
    window_vote(gts[m], k, hwin, s, ft.w[l], ft.h[l], t0, t1, ft.lo[l], ft.hi[l], ft.radius[l],
                keys + (off + t0 - p_lo));
  }
  __syncthreads();

  // ---- 3. num_pos of the image ---------------------------------------------------------------------
  {
    int c = 0;
    for (int i = tid; i < n_own; i += kThreads) c += keys[i] != kNoWinner ? 1 : 0;
    const int total = (int)block_sum_f((float)c, s_red);        // <= chunk < 2^24: exact in fp32
    cluster.barrier_wait();
    if (tid < csize) cluster.map_shared_rank(s_count, tid)[rank] = total;
  }
  cluster.sync();
  int npos_i = 0;
  for (int r = 0; r < csize; ++r) npos_i += s_count[r];
  const float np = fmaxf((float)npos_i, 1.f);
  const float scale_box = (grad_box ? grad_box[b] : inv_batch) / np;
  const float scale_cnt = (grad_cnt ? grad_cnt[b] : inv_batch) / np;
  const bool has_cnt = ft.has_cnt != 0;

  // ---- 4a. targets of every point, zero gradients of the negatives ---------------------------------
  const size_t out0 = (size_t)b * P;
  for (int i = tid; i < n_own; i += kThreads) {
    const int p = p_lo + i;
    const int l = level_of(ft, p);
    const int pos = p - ft.point_off[l];
    const int hw = ft.hw[l];
    const unsigned long long key = keys[i];
    long long label = 0;
    float cnt = -1.f;
    float4 reg = make_float4(-1.f, -1.f, -1.f, -1.f);
    const size_t base = (size_t)b * hw + pos;                  // index into a 1-channel map of the level
    if (key != kNoWinner) {
      const GtEntry g = gts[(unsigned)(key & 0xffffffffull)];
      const int w = ft.w[l];
      const int row = pos / w, col = pos - row * w;
      positive_targets(g, col, row, ft.stride[l], &reg, &cnt);
      label = (long long)g.label;
      const float* rg = ft.reg[l] + (size_t)b * 4 * hw + pos;  // the predictions pass B will need
      prefetch_l2(rg);
      prefetch_l2(rg + hw);
      prefetch_l2(rg + 2 * hw);
      prefetch_l2(rg + 3 * hw);
      if (has_cnt) prefetch_l2(ft.cnt[l] + base);
    } else {
      float* go = ft.greg[l] + (size_t)b * 4 * hw + pos;
      stg_stream_f1(go, 0.f);
      stg_stream_f1(go + hw, 0.f);
      stg_stream_f1(go + 2 * hw, 0.f);
      stg_stream_f1(go + 3 * hw, 0.f);
      if (has_cnt) stg_stream_f1(ft.gcnt[l] + base, 0.f);
    }
    const size_t o = out0 + p;
    stg_stream_s64(cls_t + o, label);
    stg_stream_f1(cnt_t + o, cnt);
    stg_stream_f4(reg_t + 4 * o, reg);
  }

  // ---- 4b. positives: loss terms and scaled gradients ----------------------------------------------
  float acc_box = 0.f, acc_cnt = 0.f;
  for (int i = tid; i < n_own; i += kThreads) {
    const unsigned long long key = keys[i];
    if (key == kNoWinner) continue;
    const int p = p_lo + i;
    const int l = level_of(ft, p);
    const int pos = p - ft.point_off[l];
    const int hw = ft.hw[l], w = ft.w[l];
    const int row = pos / w, col = pos - row * w;
    float4 tg;
    float ct;
    positive_targets(gts[(unsigned)(key & 0xffffffffull)], col, row, ft.stride[l], &tg, &ct);
    const size_t rbase = (size_t)b * 4 * hw + pos;
    const float* rg = ft.reg[l] + rbase;
    const float4 pr = make_float4(rg[0], rg[hw], rg[2 * hw], rg[3 * hw]);
    float4 g;
    acc_box += box_term<true>(pr, tg, mode, &g);
    float* go = ft.greg[l] + rbase;
    stg_stream_f1(go, g.x * scale_box);
    stg_stream_f1(go + hw, g.y * scale_box);
    stg_stream_f1(go + 2 * hw, g.z * scale_box);
    stg_stream_f1(go + 3 * hw, g.w * scale_box);
    if (has_cnt) {
      const size_t base = (size_t)b * hw + pos;
      const float x = ft.cnt[l][base];
      acc_cnt += bce_term(x, ct);
      stg_stream_f1(ft.gcnt[l] + base, scale_cnt * (sigmoid_f32(x) - ct));
    }
  }

  // ---- 5. per-image losses (rank order), batch means (image order) ---------------------------------
  const float tot_box = block_sum_f(acc_box, s_red);
  const float tot_cnt = block_sum_f(acc_cnt, s_red);
  if (tid == 0) {
    float* dst = cluster.map_shared_rank(s_part, 0);
    dst[2 * rank] = tot_box;
    dst[2 * rank + 1] = tot_cnt;
  }
  cluster.sync();
  if (rank == 0 && tid == 0) {
    float tb = 0.f, tc = 0.f;
    for (int r = 0; r < csize; ++r) {
      tb += s_part[2 * r];
      tc += s_part[2 * r + 1];
    }
    box_loss[b] = tb / np;
    if (cnt_loss) cnt_loss[b] = tc / np;
    num_pos[b] = np;
    if (mean_out) {
      __threadfence();
      const unsigned done = atomicAdd(ticket, 1u);
      if (done == gridDim.y - 1) {                              // every other image's losses are visible
        __threadfence();
        float mb = 0.f, mc = 0.f;
        for (unsigned i = 0; i < gridDim.y; ++i) {
          mb += __ldcg(box_loss + i);
          if (cnt_loss) mc += __ldcg(cnt_loss + i);
        }
        mean_out[0] = mb / (float)gridDim.y;
        mean_out[1] = mc / (float)gridDim.y;
        *ticket = 0u;                                           // ready for the next launch
      }
    }
  }
}

// In-place multiply of up to kMaxScaleMaps arrays, each by its own device scalar; a map whose scalar is
// exactly 1 costs nothing (the usual case: the fused kernel already applied d(mean)/d(loss[b]) = 1/B
// and backward() feeds 1).
constexpr int kMaxScaleMaps = 2 * B200DET_MAX_LEVELS;
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

}  // namespace
}  // namespace b200det

using namespace b200det;

extern "C" size_t b200det_assign_loss_workspace_bytes(void) { return 256; }

extern "C" int b200det_assign_loss_fused(const b200det_level* levels, float* const* reg_grads, float* const* cnt_grads,
                                         int n_levels, const float* limit_lo, const float* limit_hi,
                                         const float* radius_px, int batch, int max_gt, const float* gt_boxes,
                                         const int64_t* gt_labels, int mode, const float* grad_box,
                                         const float* grad_cnt, int64_t* cls_t, float* cnt_t, float* reg_t,
                                         float* box_loss, float* cnt_loss, float* num_pos, float* mean_out,
                                         void* workspace, void* stream) {
  if (!levels || n_levels <= 0 || n_levels > B200DET_MAX_LEVELS || !limit_lo || !limit_hi || !radius_px ||
      batch <= 0 || batch > 65535 || max_gt < 0 || max_gt >= (1 << 24) || !reg_grads || !cls_t || !cnt_t || !reg_t ||
      !box_loss || !num_pos)
    return B200DET_ERR_ARG;
  if (max_gt > 0 && (!gt_boxes || !gt_labels)) return B200DET_ERR_ARG;
  if (!aligned16(gt_boxes) || !aligned16(reg_t)) return B200DET_ERR_ARG;
  if (mode != 0 && mode != 1) return B200DET_ERR_UNSUPPORTED;
  if (mean_out && !workspace) return B200DET_ERR_ARG;
  const bool has_cnt = cnt_grads != nullptr;
  if (has_cnt != (cnt_loss != nullptr)) return B200DET_ERR_ARG;
  FusedTable ft;
  long long off = 0;
  for (int l = 0; l < B200DET_MAX_LEVELS; ++l) {
    const bool on = l < n_levels;
    if (on) {
      if (levels[l].h <= 0 || levels[l].w <= 0 || levels[l].stride <= 0 || !levels[l].reg || !reg_grads[l])
        return B200DET_ERR_ARG;
      if (has_cnt && (!levels[l].cnt || !cnt_grads[l])) return B200DET_ERR_ARG;
    }
    ft.reg[l] = on ? static_cast<const float*>(levels[l].reg) : nullptr;
    ft.cnt[l] = on && has_cnt ? static_cast<const float*>(levels[l].cnt) : nullptr;
    ft.greg[l] = on ? reg_grads[l] : nullptr;
    ft.gcnt[l] = on && has_cnt ? cnt_grads[l] : nullptr;
    ft.h[l] = on ? levels[l].h : 0;
    ft.w[l] = on ? levels[l].w : 0;
    ft.stride[l] = on ? levels[l].stride : 0;
    ft.hw[l] = ft.h[l] * ft.w[l];
    ft.lo[l] = on ? limit_lo[l] : 0.f;
    ft.hi[l] = on ? limit_hi[l] : 0.f;
    ft.radius[l] = on ? radius_px[l] : 0.f;
    ft.point_off[l] = (int)off;
    off += ft.hw[l];
    if (off > (1ll << 30)) return B200DET_ERR_ARG;
  }
  ft.point_off[B200DET_MAX_LEVELS] = (int)off;
  ft.n_levels = n_levels;
  ft.num_points = (int)off;
  ft.has_cnt = has_cnt ? 1 : 0;

  // Cluster size / CTA width.  B200DET_FUSED_CFG="<cluster>x<threads>" overrides (tuning knob).
  int csize = 8, threads = 256;
  static const char* cfg = getenv("B200DET_FUSED_CFG");
  if (cfg) {
    int c = 0, t = 0;
    if (sscanf(cfg, "%dx%d", &c, &t) == 2 && (c == 1 || c == 2 || c == 4 || c == 8 || c == 16) &&
        (t == 128 || t == 256 || t == 512)) {
      csize = c;
      threads = t;
    }
  }
  ft.chunk = (int)((off + csize - 1) / csize);
  const size_t smem = (size_t)ft.chunk * 8 + (size_t)max_gt * (sizeof(GtEntry) + sizeof(int) * n_levels) + 16;
  if (smem > 200 * 1024) return B200DET_ERR_UNSUPPORTED;

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto launch = [&](auto kernel) -> int {
    cudaError_t e = cudaSuccess;
    if (smem > 40 * 1024) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && csize > 8) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
    cudaLaunchConfig_t cfgl = {};
    cfgl.gridDim = dim3(csize, batch);
    cfgl.blockDim = dim3(threads);
    cfgl.dynamicSmemBytes = smem;
    cfgl.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfgl.attrs = attr;
    cfgl.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfgl, kernel, ft, max_gt, gt_boxes, reinterpret_cast<const long long*>(gt_labels), mode,
                           grad_box, grad_cnt, 1.0f / (float)batch, reinterpret_cast<long long*>(cls_t), cnt_t, reg_t,
                           box_loss, cnt_loss, num_pos, mean_out, static_cast<unsigned*>(workspace));
    if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
    return B200DET_OK;
  };
  const int rc = threads == 128 ? launch(assign_loss_fused_kernel<128>)
               : threads == 512 ? launch(assign_loss_fused_kernel<512>)
                                : launch(assign_loss_fused_kernel<256>);
  if (rc) return rc;
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
