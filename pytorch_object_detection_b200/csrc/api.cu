// C-ABI glue: status strings, the fused post-process entry (K1 -> K2 -> K3) and ClipBoxes.
#include "nms.cuh"

#include <stdlib.h>
#include <string.h>

namespace b200det {

int launch_score_points(const LevelTable& lt, int batch, int num_classes, float* score, int16_t* cls0,
                        cudaStream_t stream);
int launch_select_topk(const LevelTable& lt, int batch, const float* score, const int16_t* cls0, float thr,
                       int max_box, const CandSet& out, int32_t* cand_point, cudaStream_t stream);

namespace {
thread_local char g_cuda_error[256] = "";

__global__ void clip_boxes_kernel(float4* __restrict__ boxes, const long long n, const float xmax, const float ymax) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 v = boxes[i];
  v.x = fminf(fmaxf(v.x, 0.f), xmax);
  v.y = fminf(fmaxf(v.y, 0.f), ymax);
  v.z = fminf(fmaxf(v.z, 0.f), xmax);
  v.w = fminf(fmaxf(v.w, 0.f), ymax);
  boxes[i] = v;
}

struct PostWorkspace {
  float* score;
  int16_t* cls0;
  void* nms_base;
  size_t bytes;
};

PostWorkspace carve_post(void* base, int batch, int num_points, int cap) {
  PostWorkspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? static_cast<char*>(base) + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  w.score = reinterpret_cast<float*>(take((size_t)batch * num_points * 4));
  w.cls0 = reinterpret_cast<int16_t*>(take((size_t)batch * num_points * 2));
  w.nms_base = take(nms_set_workspace_bytes(batch, cap));
  w.bytes = off;
  return w;
}
}  // namespace

void set_cuda_error(cudaError_t e) {
  strncpy(g_cuda_error, cudaGetErrorString(e), sizeof(g_cuda_error) - 1);
  g_cuda_error[sizeof(g_cuda_error) - 1] = 0;
}

}  // namespace b200det

using namespace b200det;

extern "C" int b200det_abi_version(void) { return B200DET_ABI_VERSION; }

extern "C" const char* b200det_status_string(int status) {
  switch (status) {
    case B200DET_OK: return "ok";
    case B200DET_ERR_ARG: return "bad argument (null/misaligned pointer, non-positive size, too many levels)";
    case B200DET_ERR_UNSUPPORTED: return "size or mode outside what the kernels cover";
    case B200DET_ERR_WORKSPACE: return "workspace too small";
    case B200DET_ERR_CUDA: return "CUDA launch failed";
    default: return "unknown status";
  }
}

extern "C" const char* b200det_last_cuda_error(void) { return g_cuda_error; }

extern "C" int b200det_select_topk(const b200det_level* levels, int n_levels, int batch, const float* score,
                                   const int16_t* cls0, float score_thr, int max_box, float* cand_score,
                                   int32_t* cand_cls, float* cand_box, int32_t* cand_point, int32_t* cand_count,
                                   void* stream) {
  LevelTable lt;
  if (!make_level_table(levels, n_levels, &lt) || batch <= 0 || !score || !cls0 || max_box <= 0 || !cand_score ||
      !cand_cls || !cand_box || !cand_point || !cand_count || !aligned16(cand_box))
    return B200DET_ERR_ARG;
  if (max_box > B200DET_MAX_BOX) return B200DET_ERR_UNSUPPORTED;
  for (int l = 0; l < n_levels; ++l)
    if (!levels[l].reg) return B200DET_ERR_ARG;
  // cand_point doubles as the (identity) keep-source column; no NMS boxes, no mode needed
  CandSet out{cand_score, cand_cls, cand_box, cand_point, nullptr, cand_count, cand_count, max_box};
  return launch_select_topk(lt, batch, score, cls0, score_thr, max_box, out, cand_point,
                            static_cast<cudaStream_t>(stream));
}

extern "C" size_t b200det_postprocess_workspace_bytes(int batch, int num_points, int max_box) {
  if (batch <= 0 || num_points <= 0 || max_box <= 0 || max_box > B200DET_MAX_BOX) return 0;
  return carve_post(nullptr, batch, num_points, max_box).bytes;
}

extern "C" int b200det_postprocess(const b200det_level* levels, int n_levels, int batch, int num_classes,
                                   float score_thr, double nms_thr, int max_box, int clip_h, int clip_w,
                                   void* workspace, size_t workspace_bytes, float* out_score, int64_t* out_cls,
                                   float* out_box, int64_t* out_keep, int32_t* out_count, void* stream) {
  LevelTable lt;
  if (!make_level_table(levels, n_levels, &lt) || batch <= 0 || batch > 65535 || num_classes <= 0 ||
      num_classes > 32767 || max_box <= 0 || !workspace || !out_score || !out_cls || !out_box || !out_keep ||
      !out_count || !aligned16(workspace) || !aligned16(out_box))
    return B200DET_ERR_ARG;
  if (max_box > B200DET_MAX_BOX) return B200DET_ERR_UNSUPPORTED;
  for (int l = 0; l < n_levels; ++l)
    if (!levels[l].cls || !levels[l].cnt || !levels[l].reg) return B200DET_ERR_ARG;
  const PostWorkspace w = carve_post(workspace, batch, lt.num_points, max_box);
  if (workspace_bytes < w.bytes) return B200DET_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CandSet set;
  unsigned long long* mask;
  nms_set_carve(w.nms_base, batch, max_box, &set, &mask);

  int rc = launch_score_points(lt, batch, num_classes, w.score, w.cls0, st);
  if (rc) return rc;
  NmsOut out{out_score, reinterpret_cast<long long*>(out_cls), out_box, reinterpret_cast<long long*>(out_keep),
             out_count, max_box};
  // K2 + K3: one fused kernel (one CTA per image) when the candidate set fits (<= 1024), else three kernels.
  // B200DET_NO_FUSED=1 forces the three-kernel chain (A/B testing; results are bit-identical).
  static const bool no_fused = getenv("B200DET_NO_FUSED") && getenv("B200DET_NO_FUSED")[0] == '1';
  if (!no_fused && fused_supported(lt, max_box, nms_thr))
    return launch_fused_select_nms(lt, batch, w.score, w.cls0, score_thr, max_box, set, nms_thr, clip_h, clip_w, out, st);
  rc = launch_select_topk(lt, batch, w.score, w.cls0, score_thr, max_box, set, nullptr, st);
  if (rc) return rc;
  return launch_nms(set, batch, nms_thr, clip_h, clip_w, mask, out, st);
}

extern "C" int b200det_clip_boxes(float* boxes, int64_t n_boxes, int img_h, int img_w, void* stream) {
  if (n_boxes < 0 || (n_boxes > 0 && !boxes) || !aligned16(boxes)) return B200DET_ERR_ARG;
  if (n_boxes == 0) return B200DET_OK;
  const int threads = 256;
  const long long blocks = (n_boxes + threads - 1) / threads;
  clip_boxes_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(boxes), n_boxes, (float)(img_w - 1), (float)(img_h - 1));
  return check_launch();
}
