// K4 fused — target assignment + box (IoU / GIoU) loss + centerness BCE loss, forward AND backward.
// Replaces, for one training step,
//   FCOSGenTargets.forward            model/modules/head.py:218-316
//   compute_reg_loss (+ giou / iou)   model/loss.py:116-177      and its autograd backward
//   compute_cnt_loss                  model/loss.py:29-57        and its autograd backward
//   the two `.mean()` of FCOSLoss     model/loss.py:210-213
// assign_stream_kernel<kLoss = true> (assign_stream.cuh, which describes the design) + the one-CTA finalize kernel as
// its programmatic dependent; this file holds the C entry point and the small kernels behind the eager-gradient autograd functions (scale / rescale of maps).
// HBM traffic: 48 B written per point + ~36 B read per positive; GT boxes are read once per CTA.
// History (B200, config 3: B=32, P=23265, M<=100, kernels only): separate assign + loss forward + backward kernels
// 26.6 us; count kernel -> tile kernel -> finalize kernel chained by programmatic dependent launch 18.2 us; this
// design see DESIGN.md.
#include <stdio.h>
#include <stdlib.h>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

B200DET_TRACE_BUFFER(train)

#include "assign_stream.cuh"



namespace b200det {
namespace {


// ---- per-image losses (tile partials added in tile order), batch means (image order) ------------------
// One CTA, launched as a programmatic dependent of the streaming kernel: it is resident before that kernel ends and
// proceeds as soon as its partials are complete and visible.  Publishes num_pos[] (the image's counter holds its
// positives: every tile has arrived) and clears the counters for the next call.
constexpr int kFinalThreads = 256;
constexpr int kFinalStage = 1024;            // tile partials (float4: box, cnt, d/d scale, -) staged per round
constexpr int kFinalImages = 128;            // images per round at most

__global__ void __launch_bounds__(kFinalThreads)
finalize_losses_kernel(const int batch, const int n_tiles, const AssignTable at, const float* __restrict__ partial,
                       unsigned long long* __restrict__ counters, const float* __restrict__ grad_box, const int grad_mode,
                       const float* __restrict__ grad_cnt, const float inv_batch, float* __restrict__ num_pos,
                       float* __restrict__ box_loss, float* __restrict__ cnt_loss, float* __restrict__ mean_out,
                       float* __restrict__ reg_scale_grad) {
  __shared__ float4 stage[kFinalStage];
  __shared__ float s_np[kFinalImages];
  __shared__ float2 img[kFinalImages];
  __shared__ float img_dsc[kFinalImages][B200DET_MAX_LEVELS];    // per image, per level: scaled d/d scale
  pdl_launch_dependents();                       // the stream's next kernel may become resident; it waits for this one
  pdl_wait();
  const int tid = threadIdx.x;
  const int per_round = min(kFinalImages, kFinalStage / n_tiles);   // images per round (n_tiles <= kFinalStage)
  float mb = 0.f, mc = 0.f, my_dsc = 0.f;
  for (int i0 = 0; i0 < batch; i0 += per_round) {
    const int n = min(per_round, batch - i0);
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(partial) + (size_t)i0 * n_tiles;
    for (int t = tid; t < n * n_tiles; t += kFinalThreads) stage[t] = __ldcg(src + t);
    for (int i = tid; i < n; i += kFinalThreads) {               // (in flight together with the partials)
      s_np[i] = fmaxf((float)(unsigned)(__ldcg(counters + i0 + i) & 0xffffffffull), 1.f);
      counters[i0 + i] = 0ull;                                   // cleared for the next call
    }
    __syncthreads();
    for (int i = tid; i < n; i += kFinalThreads) {
      const float np = s_np[i];
      num_pos[i0 + i] = np;
      const float up = upstream_of(grad_box, grad_mode, i0 + i, inv_batch) / np;
      float tb = 0.f, tc = 0.f;
      for (int l = 0; l < at.n_levels; ++l) {
        float td = 0.f;
        for (int t = at.tile_off[l]; t < at.tile_off[l + 1]; ++t) {
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
    if (tid < at.n_levels)                                     // thread l: level l's scale gradient, image order
      for (int i = 0; i < n; ++i) my_dsc += img_dsc[i][tid];
  }
  if (tid == 0 && mean_out) {
    mean_out[0] = mb / (float)batch;
    mean_out[1] = mc / (float)batch;
    // the upstream gradients of the two means this call's gradients were written for (grad_mode 1), for its backward
    mean_out[2] = (grad_mode && grad_box) ? __ldcg(grad_box) : 1.f;
    mean_out[3] = (grad_mode && grad_cnt) ? __ldcg(grad_cnt) : 1.f;
  }
  if (reg_scale_grad && tid < at.n_levels) reg_scale_grad[tid] = my_dsc;
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
  const float* assumed[kMaxStates];          // the value THIS forward wrote its gradients for (its own copy), or NULL = *state
  float* state[kMaxStates];                  // shared {upstream gradient the next forward will assume, ticket (an unsigned)}
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
  pdl_launch_dependents();                       // (launched as a programmatic dependent: the launch latency of this
  pdl_wait();                                    //  usually empty kernel hides behind its predecessor)
  float f[kMaxStates];
  bool any = false, any_update = false;
#pragma unroll
  for (int s = 0; s < kMaxStates; ++s) {
    f[s] = 1.0f;
    if (s < n_states) {
      const float got = __ldcg(t.got[s]), shared = __ldcg(t.state[s]);
      const float assumed = t.assumed[s] ? __ldcg(t.assumed[s]) : shared;
      if (got != assumed) {
        f[s] = got / assumed;
        any = true;
      }
      any_update |= got != shared;
    }
  }
  if (!any && !any_update) return;
  if (any) {
    const long long step = (long long)gridDim.x * 256;
    const long long i0 = (long long)blockIdx.x * 256 + threadIdx.x;
    for (int i = 0; i < n_maps; ++i) {
      const int so = t.state_of[i];
      const float fi = so == 0 ? f[0] : so == 1 ? f[1] : so == 2 ? f[2] : f[3];
      if (fi != 1.0f) rescale_map(static_cast<T*>(t.map[i]), t.numel[i], fi, i0, step);
    }
  }
  // remember the upstream gradient that arrived for the NEXT forward (a zero / non-finite one is not a usable
  // assumption: its gradients could not be rescaled).  The shared word changes only after every CTA has read it.
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
// workspace: [per-image arrival / positive counters, B x u64: zero between calls (finalize_losses_kernel clears them)]
// [tile partials B x tiles x 4 floats]
size_t state_bytes(int batch) { return align_up((size_t)batch * sizeof(unsigned long long), 256); }
int train_tiles(int num_points) { return (num_points + 4 * kStreamThreads - 1) / (4 * kStreamThreads) + B200DET_MAX_LEVELS; }

template <int kPts>
cudaError_t launch_stream_loss(bool scale_exp, dim3 grid, size_t smem, cudaStream_t st, const AssignTable& at,
                               const LossMaps& lm, const StreamArgs& a, bool pdl) {
  return scale_exp ? launch_stream_kernel(assign_stream_kernel<kPts, true, true>, grid, smem, st, at, lm, a, pdl)
                   : launch_stream_kernel(assign_stream_kernel<kPts, true, false>, grid, smem, st, at, lm, a, pdl);
}
}  // namespace

extern "C" size_t b200det_assign_loss_workspace_bytes(int batch, int num_points) {
  if (batch <= 0 || num_points <= 0) return 0;
  return state_bytes(batch) + align_up((size_t)batch * train_tiles(num_points) * 4 * sizeof(float), 256);
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
  if ((mode != 0 && mode != 1) || (grad_mode != 0 && grad_mode != 1) || !fp32_levels(levels, n_levels))
    return B200DET_ERR_UNSUPPORTED;
  const bool has_cnt = cnt_grads != nullptr;
  if (has_cnt != (cnt_loss != nullptr)) return B200DET_ERR_ARG;
  int32_t level_hw[2 * B200DET_MAX_LEVELS], strides[B200DET_MAX_LEVELS];
  LossMaps lm = {};
  bool scale_exp = false;
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
    scale_exp |= levels[l].reg_scale != nullptr;
  }
  const int pts = stream_points_per_thread(level_hw, n_levels, batch);
  AssignTable at;
  if (!make_assign_table(level_hw, strides, limit_lo, limit_hi, radius_px, n_levels, kStreamThreads * pts, &at))
    return B200DET_ERR_ARG;
  const int n_tiles = at.tile_off[B200DET_MAX_LEVELS];
  if (n_tiles > 65535 || n_tiles > train_tiles(at.num_points) || n_tiles > kFinalStage) return B200DET_ERR_UNSUPPORTED;
  if (workspace_bytes < b200det_assign_loss_workspace_bytes(batch, at.num_points)) return B200DET_ERR_WORKSPACE;
  const size_t smem = (size_t)max_gt * (sizeof(GtEntry) + 3 * sizeof(int)) + (size_t)((at.num_points + 31) / 32) * 4;
  if (smem > 160 * 1024) return B200DET_ERR_UNSUPPORTED;
  static const bool no_pdl = getenv("B200DET_NO_PDL") && getenv("B200DET_NO_PDL")[0] == '1';
  StreamArgs a = {};
  a.has_cnt = has_cnt ? 1 : 0;
  a.M = max_gt;
  a.mode = mode;
  a.grad_mode = grad_mode;
  a.n_tiles = n_tiles;
  // B200DET_FORCE_RECOUNT=1 (tests): no wait for the arrival counter, every CTA with positives recounts its image
  const bool force_recount = getenv("B200DET_FORCE_RECOUNT") && getenv("B200DET_FORCE_RECOUNT")[0] == '1';
  a.arrival_polls = force_recount ? 0 : kArrivalPolls;
  a.inv_batch = 1.0f / (float)batch;
  a.gt_boxes = gt_boxes;
  a.gt_labels = reinterpret_cast<const long long*>(gt_labels);
  a.grad_box = grad_box;
  a.grad_cnt = grad_cnt;
  char* ws = static_cast<char*>(workspace);
  a.counters = reinterpret_cast<unsigned long long*>(ws);
  a.partial = reinterpret_cast<float*>(ws + state_bytes(batch));
  a.cls_t = reinterpret_cast<long long*>(cls_t);
  a.cnt_t = cnt_t;
  a.reg_t = reg_t;
  a.gt_index = nullptr;
  a.batch = batch;
  const dim3 grid(n_tiles, batch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = pts == 4   ? launch_stream_loss<4>(scale_exp, grid, smem, st, at, lm, a, !no_pdl)
                        : pts == 6 ? launch_stream_loss<6>(scale_exp, grid, smem, st, at, lm, a, !no_pdl)
                                   : launch_stream_loss<8>(scale_exp, grid, smem, st, at, lm, a, !no_pdl);
  if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
  int rc = check_launch();
  if (rc) return rc;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(kFinalThreads);
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  const cudaError_t e2 = cudaLaunchKernelEx(&cfg, finalize_losses_kernel, batch, n_tiles, at, (const float*)a.partial,
                                            a.counters, grad_box, grad_mode, grad_cnt, 1.0f / (float)batch, num_pos,
                                            box_loss, cnt_loss, mean_out, reg_scale_grad);
  if (e2 != cudaSuccess) { set_cuda_error(e2); return B200DET_ERR_CUDA; }
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
                                    int n_maps, const float* const* got, const float* const* assumed,
                                    float* const* state, int n_states, void* stream) {
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
    t.assumed[s] = assumed ? assumed[s] : nullptr;
    t.state[s] = state[s];
  }
  static const bool no_pdl = getenv("B200DET_NO_PDL") && getenv("B200DET_NO_PDL")[0] == '1';
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.gridDim = dim3(kRescaleCtas);
  cfg.blockDim = dim3(256);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  cudaError_t e;
  if (dtype == B200DET_F32) e = cudaLaunchKernelEx(&cfg, rescale_maps_kernel<float>, t, n_maps, n_states);
  else if (dtype == B200DET_F16) e = cudaLaunchKernelEx(&cfg, rescale_maps_kernel<__half>, t, n_maps, n_states);
  else e = cudaLaunchKernelEx(&cfg, rescale_maps_kernel<__nv_bfloat16>, t, n_maps, n_states);
  if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
  return check_launch();
}
