// N3 — the step after the path: VOC-style average precision (test.py:15-162, eval_ap_2d) on the device, fed by
// the padded detections FCOSHead.detect() leaves in HBM, so an evaluation epoch needs no per-image D2H copy and
// none of the reference's O(N) np.append loops.
//
//   ap_match_kernel   one CTA per image, one warp per class at a time.  The image's detections are walked in
//                     score order (they arrive sorted: head.py:69-80); for a detection of class c the lanes
//                     evaluate the reference's fp32 IoU (test.py:24-55: GT first, overlap / (area_gt + area_det -
//                     overlap)) against the image's GT boxes of class c, the warp takes the arg-max (first index
//                     on ties, np.argmax), and the detection is a true positive iff IoU >= threshold and that GT
//                     box is still free (test.py:121-140).  Also counts GT and detections per class.
//   ap_class_kernel   one CTA per class: gathers the class's detections from all images, sorts them by score
//                     (descending, stable by image/rank where np.argsort(-scores) leaves ties unspecified), runs
//                     the cumulative TP count, precision = tp / (tp + fp), recall = tp / total_gt in fp64, the
//                     precision envelope and the area under the curve exactly as _compute_ap (test.py:58-83):
//                     AP = sum over true positives of (recall_i - recall_{i-1}) * max_{j >= i} precision_j.
//                     total_gt = 0 with detections gives NaN (0/0 recall), no detections gives 0, as numpy does.
#include "block_utils.cuh"

namespace b200det {
namespace {

constexpr int kMatchThreads = 256;
constexpr int kMatchWarps = kMatchThreads / 32;

__device__ __forceinline__ float iou_ref_f32(const float4 g, const float4 d) {        // test.py:35-55
  const float w = fmaxf(0.f, __fsub_rn(fminf(g.z, d.z), fmaxf(g.x, d.x)));
  const float h = fmaxf(0.f, __fsub_rn(fminf(g.w, d.w), fmaxf(g.y, d.y)));
  const float overlap = __fmul_rn(w, h);
  const float area_g = __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
  const float area_d = __fmul_rn(__fsub_rn(d.z, d.x), __fsub_rn(d.w, d.y));
  return __fdiv_rn(overlap, __fsub_rn(__fadd_rn(area_g, area_d), overlap));
}

__global__ void __launch_bounds__(kMatchThreads)
ap_match_kernel(const int K, const int M, const int num_cls, const long long* __restrict__ det_cls,
                const float4* __restrict__ det_box, const int32_t* __restrict__ det_count,
                const float4* __restrict__ gt_box, const long long* __restrict__ gt_label, const double iou_thr,
                unsigned char* __restrict__ tp, int32_t* __restrict__ gt_total, int32_t* __restrict__ det_total) {
  extern __shared__ unsigned assigned_all[];                 // [kMatchWarps][ceil(M / 32)] one bit per GT row
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int words = (M + 31) / 32;
  unsigned* assigned = assigned_all + warp * words;
  const int n = min(max(det_count[b], 0), K);
  const long long* dc = det_cls + (size_t)b * K;
  const float4* db = det_box + (size_t)b * K;
  const float4* gb = gt_box + (size_t)b * M;
  const long long* gl = gt_label + (size_t)b * M;
  for (int c = 1 + warp; c < num_cls; c += kMatchWarps) {
    int g = 0;
    for (int m = lane; m < M; m += 32) g += gl[m] == c ? 1 : 0;
    for (int wd = lane; wd < words; wd += 32) assigned[wd] = 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) g += __shfl_xor_sync(0xffffffffu, g, d);
    __syncwarp();
    int nd = 0;
    for (int k0 = 0; k0 < n; k0 += 32) {
      const int k = k0 + lane;
      unsigned todo = __ballot_sync(0xffffffffu, k < n && dc[k] == c);
      nd += __popc(todo);
      while (todo) {
        const int kk = k0 + __ffs(todo) - 1;
        todo &= todo - 1;
        bool hit = false;
        if (g > 0) {
          const float4 d4 = db[kk];
          float best = -1.f;                                  // IoU >= 0; NaN (degenerate 0/0) never wins
          int best_m = 0x7fffffff;
          for (int m = lane; m < M; m += 32) {
            if (gl[m] != c) continue;
            const float v = iou_ref_f32(gb[m], d4);
            if (v > best) { best = v; best_m = m; }           // ascending m: first index of the lane's maximum
          }
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, d);
            const int om = __shfl_xor_sync(0xffffffffu, best_m, d);
            if (ov > best || (ov == best && om < best_m)) { best = ov; best_m = om; }
          }
          if (lane == 0 && best_m != 0x7fffffff && (double)best >= iou_thr) {
            const unsigned bit = 1u << (best_m & 31);
            if (!(assigned[best_m >> 5] & bit)) {             // one GT box is given to one detection only
              assigned[best_m >> 5] |= bit;
              hit = true;
            }
          }
          __syncwarp();
        }
        if (lane == 0) tp[(size_t)b * K + kk] = hit ? 1 : 0;
      }
    }
    if (lane == 0) {
      if (g) atomicAdd(gt_total + c, g);
      if (nd) atomicAdd(det_total + c, nd);
    }
    __syncwarp();
  }
}

constexpr int kApThreads = 1024;

__device__ __forceinline__ long long next_pow2_ll(long long v) {
  long long p = 1;
  while (p < v) p <<= 1;
  return p;
}

__global__ void __launch_bounds__(kApThreads)
ap_class_kernel(const long long total, const int K, const long long* __restrict__ det_cls,
                const float* __restrict__ det_score, const int32_t* __restrict__ det_count,
                const unsigned char* __restrict__ tp, const int32_t* __restrict__ gt_total,
                const int32_t* __restrict__ det_total, unsigned long long* __restrict__ scratch,
                double* __restrict__ ap) {
  __shared__ int s_scan[33];
  __shared__ int s_n;
  __shared__ int s_cum[kApThreads];
  __shared__ double s_ap, s_runmax;
  const int c = blockIdx.x + 1, tid = threadIdx.x;
  const int n = det_total[c];
  const int G = gt_total[c];
  if (n == 0) {                                              // recall = [] -> (1 - 0) * 0
    if (tid == 0) ap[c] = 0.0;
    return;
  }
  long long off = 0;
  for (int q = 1; q < c; ++q) off += next_pow2_ll(det_total[q]);
  const int n2 = (int)next_pow2_ll(n);
  unsigned long long* keys = scratch + off;

  // gather the class's detections: key = (score order, ~flat index): descending sort = score desc, index asc
  if (tid == 0) s_n = 0;
  __syncthreads();
  for (long long i = tid; i < total; i += kApThreads) {
    const int b = (int)(i / K), k = (int)(i - (long long)b * K);
    if (k < det_count[b] && det_cls[i] == c) {
      const int slot = atomicAdd(&s_n, 1);
      keys[slot] = ((unsigned long long)order_key(det_score[i]) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
    }
  }
  for (int i = n + tid; i < n2; i += kApThreads) keys[i] = 0ull;
  __syncthreads();
  bitonic_sort_desc(keys, n2);

  // cumulative true positives in sorted order (the key slot is reused for the count)
  int carry = 0;
  for (int i0 = 0; i0 < n; i0 += kApThreads) {
    const int i = i0 + tid;
    int flag = 0;
    if (i < n) flag = tp[0xffffffffu - (unsigned)(keys[i] & 0xffffffffull)];
    int tot;
    const int excl = block_exclusive_scan(flag, s_scan, &tot);
    if (i < n) keys[i] = (unsigned long long)(unsigned)(carry + excl + flag);
    carry += tot;
  }
  __syncthreads();

  // precision envelope (running maximum from the end) and the area under the curve
  if (tid == 0) { s_ap = 0.0; s_runmax = 0.0; }              // mpre's trailing sentinel is 0
  for (int i1 = n; i1 > 0; i1 -= kApThreads) {
    const int i0 = max(0, i1 - kApThreads);
    __syncthreads();
    if (i0 + tid < i1) s_cum[tid] = (int)keys[i0 + tid];
    const int before = i0 > 0 ? (int)keys[i0 - 1] : 0;
    __syncthreads();
    if (tid == 0) {
      double run = s_runmax, acc = s_ap;
      for (int j = i1 - i0 - 1; j >= 0; --j) {
        const int cum = s_cum[j];
        const int prev = j > 0 ? s_cum[j - 1] : before;
        const double prec = (double)cum / fmax((double)(i0 + j + 1), 2.220446049250313e-16);   // tp + fp = rank
        run = fmax(run, prec);
        if (G == 0) {                                         // recall = 0/0: every mrec entry is NaN and "changes"
          acc += (double)CUDART_NAN;
        } else if (cum != prev) {                             // recall changes only at a true positive
          acc += ((double)cum / (double)G - (double)prev / (double)G) * run;
        }
      }
      s_runmax = run;
      s_ap = acc;
    }
  }
  __syncthreads();
  if (tid == 0) ap[c] = s_ap;
}

// COCO result rows (Test_coco.py:144-168): boxes / scale, then (x, y, w, h) from the scaled corners, and per image the
// number of leading detections with score >= threshold (the reference breaks at the first one below it).
__global__ void __launch_bounds__(256)
coco_boxes_kernel(const int K, const float4* __restrict__ det_box, const float* __restrict__ det_score,
                  const int32_t* __restrict__ det_count, const float* __restrict__ scale, const float threshold,
                  float4* __restrict__ out_xywh, int32_t* __restrict__ out_count) {
  __shared__ int s_first;
  const int b = blockIdx.x;
  const int n = min(max(det_count[b], 0), K);
  const float sc = scale[b];
  if (threadIdx.x == 0) s_first = n;
  __syncthreads();
  for (int k = threadIdx.x; k < n; k += 256) {
    const float4 v = det_box[(size_t)b * K + k];
    const float x1 = __fdiv_rn(v.x, sc), y1 = __fdiv_rn(v.y, sc), x2 = __fdiv_rn(v.z, sc), y2 = __fdiv_rn(v.w, sc);
    out_xywh[(size_t)b * K + k] = make_float4(x1, y1, __fsub_rn(x2, x1), __fsub_rn(y2, y1));
    if (det_score[(size_t)b * K + k] < threshold) atomicMin(&s_first, k);
  }
  __syncthreads();
  if (threadIdx.x == 0) out_count[b] = s_first;
}

}  // namespace
}  // namespace b200det

using namespace b200det;

extern "C" size_t b200det_eval_ap_workspace_bytes(int batch, int max_det, int num_cls) {
  if (batch <= 0 || max_det <= 0 || num_cls <= 1) return 0;
  const size_t total = (size_t)batch * max_det;
  // tp flags | gt_total, det_total [num_cls] | sort scratch: sum of next_pow2(n_c) <= 2 * total + num_cls keys
  return align_up(total, 256) + align_up((size_t)2 * num_cls * 4, 256) + align_up((2 * total + num_cls) * 8, 256);
}

extern "C" int b200det_eval_ap(int batch, int max_det, int max_gt, int num_cls, const float* det_score,
                               const int64_t* det_cls, const float* det_box, const int32_t* det_count,
                               const float* gt_boxes, const int64_t* gt_labels, double iou_thr, void* workspace,
                               size_t workspace_bytes, double* ap, void* stream) {
  if (batch <= 0 || max_det <= 0 || max_gt < 0 || num_cls <= 1 || num_cls > 65535 || !det_score || !det_cls ||
      !det_box || !det_count || !workspace || !ap || !aligned16(det_box) || !aligned16(gt_boxes) ||
      !aligned16(workspace))
    return B200DET_ERR_ARG;
  if (max_gt > 0 && (!gt_boxes || !gt_labels)) return B200DET_ERR_ARG;
  if ((long long)batch * max_det >= (1ll << 31)) return B200DET_ERR_UNSUPPORTED;
  if (workspace_bytes < b200det_eval_ap_workspace_bytes(batch, max_det, num_cls)) return B200DET_ERR_WORKSPACE;
  const size_t total = (size_t)batch * max_det;
  char* base = static_cast<char*>(workspace);
  unsigned char* tp = reinterpret_cast<unsigned char*>(base);
  int32_t* gt_total = reinterpret_cast<int32_t*>(base + align_up(total, 256));
  int32_t* det_total = gt_total + num_cls;
  unsigned long long* scratch =
      reinterpret_cast<unsigned long long*>(base + align_up(total, 256) + align_up((size_t)2 * num_cls * 4, 256));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(gt_total, 0, (size_t)2 * num_cls * 4, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(ap, 0, sizeof(double), st);          // ap[0]: the background slot
  if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
  const size_t smem = (size_t)kMatchWarps * ((max_gt + 31) / 32) * 4;
  if (smem > 160 * 1024) return B200DET_ERR_UNSUPPORTED;
  if (smem > 40 * 1024) {
    e = cudaFuncSetAttribute(ap_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
  }
  ap_match_kernel<<<batch, kMatchThreads, smem, st>>>(
      max_det, max_gt, num_cls, reinterpret_cast<const long long*>(det_cls), reinterpret_cast<const float4*>(det_box),
      det_count, reinterpret_cast<const float4*>(gt_boxes), reinterpret_cast<const long long*>(gt_labels), iou_thr, tp,
      gt_total, det_total);
  int rc = check_launch();
  if (rc) return rc;
  ap_class_kernel<<<num_cls - 1, kApThreads, 0, st>>>((long long)total, max_det, reinterpret_cast<const long long*>(det_cls),
                                                     det_score, det_count, tp, gt_total, det_total, scratch, ap);
  return check_launch();
}

extern "C" int b200det_coco_boxes(int batch, int max_det, const float* det_box, const float* det_score,
                                  const int32_t* det_count, const float* scale, float threshold, float* out_xywh,
                                  int32_t* out_count, void* stream) {
  if (batch <= 0 || batch > 65535 || max_det <= 0 || !det_box || !det_score || !det_count || !scale || !out_xywh ||
      !out_count || !aligned16(det_box) || !aligned16(out_xywh))
    return B200DET_ERR_ARG;
  coco_boxes_kernel<<<batch, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      max_det, reinterpret_cast<const float4*>(det_box), det_score, det_count, scale, threshold,
      reinterpret_cast<float4*>(out_xywh), out_count);
  return check_launch();
}
