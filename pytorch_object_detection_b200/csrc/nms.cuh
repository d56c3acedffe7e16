// Sorted candidate set shared by K2 (producer) and K3 (consumer).
#pragma once
#include "common.cuh"

namespace b200det {

constexpr int kModeTrick = 0;     // torchvision _batched_nms_coordinate_trick (numel <= 4000 on CPU)
constexpr int kModeVanilla = 1;   // torchvision _batched_nms_vanilla (per class, raw boxes)
constexpr int kModeDone = 2;      // vanilla image already finished by nms_class_kernel: the dense kernels skip it
constexpr int kTrickMaxNumel = 4000;

// Per image `count[b]` candidates in stable descending score order, row stride `cap`.
struct CandSet {
  float* score;      // [B, cap]
  int32_t* cls;      // [B, cap]   1-based class
  float* box;        // [B, cap, 4] raw boxes (what the caller gets back)
  int32_t* src;      // [B, cap]   index reported as the keep index
  float* nms_box;    // [B, cap, 4] boxes the IoU is evaluated on (offset by class in trick mode)
  int32_t* count;    // [B]
  int32_t* mode;     // [B]
  int cap;
};

// torchvision/ops/boxes.py::batched_nms as run on CPU: with m candidates,
//   m*4 <= 4000: boxes_for_nms = boxes + float(cls) * (boxes.max() + 1)      (fp32, one rounding each)
//   otherwise  : per-class NMS on the raw boxes.
// Thread `tid` must have written rows i = tid, tid + nthreads, ... of set.box / set.cls itself.
__device__ __forceinline__ void nms_prepare_boxes(const CandSet& set, int b, int m, float max_coord, int tid,
                                                  int nthreads) {
  const int mode = (m * 4 <= kTrickMaxNumel) ? kModeTrick : kModeVanilla;
  const float span = __fadd_rn(max_coord, 1.0f);
  const size_t o0 = (size_t)b * set.cap;
  for (int i = tid; i < m; i += nthreads) {
    float4 bx = reinterpret_cast<const float4*>(set.box)[o0 + i];
    if (mode == kModeTrick) {
      const float off = __fmul_rn((float)set.cls[o0 + i], span);
      bx.x = __fadd_rn(bx.x, off);
      bx.y = __fadd_rn(bx.y, off);
      bx.z = __fadd_rn(bx.z, off);
      bx.w = __fadd_rn(bx.w, off);
    }
    reinterpret_cast<float4*>(set.nms_box)[o0 + i] = bx;
  }
  if (tid == 0) set.mode[b] = mode;
}

struct NmsOut {
  float* score;        // [B, stride]
  long long* cls;      // [B, stride]
  float* box;          // [B, stride, 4]
  long long* keep;     // [B, stride]
  int32_t* count;      // [B]
  int stride;
};

// (double)iou > thr for a non-NaN fp32 iou  <=>  iou >= thr_up, the smallest fp32 strictly above thr
// (torchvision's CPU kernel compares the fp32 IoU with a double threshold).  zero_sup: thr < 0 (or
// NaN), where even a zero IoU suppresses and the overlap shortcut must not be taken.
void nms_threshold_params(double nms_thr, float* thr_up, bool* zero_sup);

// K2 + K3 as one kernel (fused.cu), for max_box <= 1024 and nms_thr >= 0
bool fused_supported(const LevelTable& lt, int max_box, double nms_thr);
int launch_fused_select_nms(const LevelTable& lt, int batch, const float* score, const int16_t* cls0, float thr,
                            int max_box, const CandSet& set, double nms_thr, int clip_h, int clip_w,
                            const NmsOut& out, cudaStream_t stream);

// the NMS half of that kernel alone, on a candidate set made by nms_prepare_kernel (cap <= 1024 and nms_thr >= 0)
bool fused_nms_supported(int cap, double nms_thr);
int launch_fused_nms_from_set(const CandSet& set, int batch, double nms_thr, int clip_h, int clip_w, const NmsOut& out,
                              cudaStream_t stream);

// candidate set + suppression mask carved out of one caller-owned workspace
size_t nms_set_workspace_bytes(int batch, int cap);
void nms_set_carve(void* base, int batch, int cap, CandSet* set, unsigned long long** mask);

// per-class path for vanilla images (nms_class.cu); marks the images it finishes kModeDone
int launch_nms_class(const CandSet& set, int batch, float thr_up, bool zero_sup, int clip_h, int clip_w,
                     const NmsOut& out, cudaStream_t stream);

int launch_nms(const CandSet& set, int batch, double nms_thr, int clip_h, int clip_w,
               unsigned long long* mask, const NmsOut& out, cudaStream_t stream);

}  // namespace b200det
