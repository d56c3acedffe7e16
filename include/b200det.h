/*
 * b200det — C ABI of the B200-native (sm_100a) FCOS detection hot path.
 *
 * The reference (hby1320/pytorch_object_detection) is pure Python and has no FFI layer;
 * its boundary is the call signatures of four nn.Modules.  Every entry point below names
 * the reference interface it replaces (paths relative to the reference tree).  The Python
 * host side (pytorch_object_detection_b200/head.py, loss.py) binds these with ctypes and
 * mirrors those modules; INTEGRATION.md shows the stub a maintainer of the reference adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; the library never allocates,
 *     frees or synchronises; all work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - tensors are dense row-major ("contiguous") in the shapes given; level maps are NCHW fp32.
 *   - return value: 0 = ok, otherwise a b200det_status; b200det_status_string() names it.
 *   - no CPU implementation exists behind this ABI.
 */
#ifndef B200DET_H_
#define B200DET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DET_ABI_VERSION 6
#define B200DET_MAX_LEVELS 8
#define B200DET_MAX_BOX 8192      /* largest max_detection_box / NMS candidate count per image */

typedef enum {
  B200DET_OK = 0,
  B200DET_ERR_ARG = 1,            /* null pointer, non-positive size, too many levels ... */
  B200DET_ERR_UNSUPPORTED = 2,    /* size or dtype outside what the kernels cover */
  B200DET_ERR_WORKSPACE = 3,      /* workspace smaller than *_workspace_bytes() */
  B200DET_ERR_CUDA = 4            /* a launch failed; see b200det_last_cuda_error() */
} b200det_status;

/* element type of a map that may be something else than fp32 */
typedef enum {
  B200DET_F32 = 0,
  B200DET_F16 = 1,
  B200DET_BF16 = 2
} b200det_dtype;

/* One FPN level of head outputs: cls [B,C,h,w], cnt [B,1,h,w], reg [B,4,h,w], fp32 (any may
 * be NULL where an entry point does not read it).  reg holds the (l, t, r, b) distances, or — when
 * reg_scale is not NULL — the raw output x of the regression convolution, the distances being
 * exp(x * *reg_scale): the reference's ScaleExp (model/modules/modules.py:170-176, applied at
 * model/od/HISFcos.py:228) folded into the consumer, so the exp'd copy of reg is never written.
 * reg_scale points to ONE fp32 on the device (the level's ScaleExp.scale parameter).  It is honoured by
 * b200det_select_topk / b200det_postprocess and b200det_assign_loss_fused; the stand-alone box-loss
 * entry points return B200DET_ERR_UNSUPPORTED for it.  `stride` is the level's pixel stride; the
 * point grid (j*stride + stride/2, i*stride + stride/2) of utill/utills.py:58-73 is computed
 * in-kernel, never materialised.  Points are numbered level-major, row-major inside a level
 * (the order reshape_cat_out produces, head.py:8-26); P = sum(h*w). */
typedef struct {
  const void* cls;
  const void* cnt;
  const void* reg;
  int32_t h, w, stride;
  int32_t dtypes;                 /* B200DET_LEVEL_DTYPES(cls_cnt, reg); 0 = everything fp32 */
  const void* reg_scale;
} b200det_level;

/* Element types of a level's maps: `cls_cnt` for the cls and cnt maps, `reg` for the reg map (b200det_dtype each).
 * Honoured by b200det_score_points / b200det_select_topk / b200det_postprocess, which read the fp16 / bf16 head
 * outputs of an autocast forward (train.py:175) as they are and evaluate in fp32 — bit-identical to the fp32 kernels
 * on up-cast inputs, without the up-cast pass.  All levels of a call must agree.  Every other entry point takes
 * fp32 maps (dtypes == 0) unless it has a dtype argument of its own. */
#define B200DET_LEVEL_DTYPES(cls_cnt, reg) ((int32_t)(cls_cnt) | ((int32_t)(reg) << 4))

int         b200det_abi_version(void);
const char* b200det_status_string(int status);
const char* b200det_last_cuda_error(void);

/* ---------------------------------------------------------------------------------------
 * Inference post-processing.  Replaces FCOSHead.forward + ClipBoxes.forward
 * (model/modules/head.py:52-102, 152-162; call sites test.py:206-207, Test_coco.py:141-142).
 * ------------------------------------------------------------------------------------- */

/* K1 — head.py:8-26,57-63 without the NHWC copy: per point
 *   score = sqrt(max_c sigmoid(cls) * sigmoid(cnt)),  cls0 = argmax_c sigmoid(cls) with torch.max's first index
 * among EQUAL fp32 sigmoid values (distinct logits can share one: saturated or a few ulps apart, head.py:57-62),
 * written level-major as score[B,P] f32 and cls0[B,P] int16 (0-based). */
int b200det_score_points(const b200det_level* levels, int n_levels, int batch, int num_classes,
                         float* score, int16_t* cls0, void* stream);

/* K2 — head.py:69-80 + the threshold of head.py:90: per image the k = min(max_box, P) highest
 * scores that are >= score_thr, sorted (score desc, point index asc), gathered to
 *   cand_score[B,max_box] f32, cand_cls[B,max_box] i32 (1-based), cand_box[B,max_box,4] f32
 *   (x-l, y-t, x+r, y+b; head.py:29-38), cand_point[B,max_box] i32, cand_count[B] i32. */
int b200det_select_topk(const b200det_level* levels, int n_levels, int batch,
                        const float* score, const int16_t* cls0, float score_thr, int max_box,
                        float* cand_score, int32_t* cand_cls, float* cand_box,
                        int32_t* cand_point, int32_t* cand_count, void* stream);

/* K3 — torchvision.ops.batched_nms as called at head.py:94, CPU semantics of torchvision
 * 0.26.0: count*4 <= 4000 -> coordinate trick (boxes + cls*(max_coord+1), all pairs);
 * otherwise per-class NMS on raw boxes.  Greedy, stable descending score order, fp32
 * IoU = inter / (area_i + area_j - inter) without FMA, suppressed when (double)IoU > nms_thr.
 * Input: per image `in_count[b]` (or `n` when in_count is NULL) candidates, row stride `n`;
 * only those with score >= score_thr take part (head.py:90-93).  n <= B200DET_MAX_BOX.
 * Output, in keep order: out_score[B,n] f32, out_cls[B,n] i64, out_box[B,n,4] f32,
 * out_keep[B,n] i64 (index into the thresholded candidate list, as batched_nms returns),
 * out_count[B] i32.  If clip_h > 0 the boxes are clamped to [0,clip_w-1]x[0,clip_h-1]
 * (ClipBoxes, head.py:152-162) while they are written. */
size_t b200det_nms_workspace_bytes(int batch, int n);
int b200det_batched_nms(int batch, int n, const float* boxes, const float* scores,
                        const int64_t* classes, const int32_t* in_count,
                        float score_thr, double nms_thr, int clip_h, int clip_w,
                        void* workspace, size_t workspace_bytes,
                        float* out_score, int64_t* out_cls, float* out_box, int64_t* out_keep,
                        int32_t* out_count, void* stream);

/* K1 + K2 + K3 on one stream: the whole of FCOSHead.forward (+ ClipBoxes when clip_h > 0).
 * Outputs have row stride max_box.  out_keep holds indices into the image's top-k list. */
size_t b200det_postprocess_workspace_bytes(int batch, int num_points, int max_box);
int b200det_postprocess(const b200det_level* levels, int n_levels, int batch, int num_classes,
                        float score_thr, double nms_thr, int max_box,
                        int clip_h, int clip_w, void* workspace, size_t workspace_bytes,
                        float* out_score, int64_t* out_cls, float* out_box, int64_t* out_keep,
                        int32_t* out_count, void* stream);

/* ClipBoxes.forward (head.py:152-162): in-place clamp of boxes[n_boxes,4]. */
int b200det_clip_boxes(float* boxes, int64_t n_boxes, int img_h, int img_w, void* stream);

/* ---------------------------------------------------------------------------------------
 * Training targets.  Replaces FCOSGenTargets.forward / generate_target
 * (model/modules/head.py:218-316; call site train.py:177).
 * ------------------------------------------------------------------------------------- */

/* level_hw[2*l] = h, level_hw[2*l+1] = w; limit_lo/hi[l] = lim_range of level l (as fp32);
 * radius_px[l] = stride * sample_radio_ratio (as fp32);
 * gt_boxes [B,M,4] f32 (x0,y0,x1,y1; padding rows -1), gt_labels [B,M] i64.
 * Outputs, level-major over P = sum(h*w): cls_t [B,P] i64 (0 = background),
 * cnt_t [B,P] f32 (-1 = negative), reg_t [B,P,4] f32 (l,t,r,b; -1 = negative),
 * gt_index [B,P] i32 (assigned GT row, -1 = negative; may be NULL). */
int b200det_assign_targets(const int32_t* level_hw, const int32_t* strides,
                           const float* limit_lo, const float* limit_hi, const float* radius_px,
                           int n_levels, int batch, int max_gt,
                           const float* gt_boxes, const int64_t* gt_labels,
                           int64_t* cls_t, float* cnt_t, float* reg_t, int32_t* gt_index,
                           void* stream);

/* ---------------------------------------------------------------------------------------
 * Losses.  Replace compute_reg_loss / compute_cnt_loss / compute_cls_loss and their
 * autograd backward (model/loss.py:6-57, 116-193; call site train.py:178-180).
 * `cnt_t` [B,P] f32 is the positive mask source: a point is positive iff cnt_t > -1
 * (loss.py:205); FCOSLoss passes the centerness target itself.
 * Every forward writes loss[B] f32 = (sum over the image) / num_pos[b] and
 * num_pos[B] f32 = clamp(count(cnt_t > -1), 1), as loss.py:26,57,139.  Every backward takes
 * grad_loss[B] f32 = dL/d(loss[b]) and writes the gradient of every level map in its own
 * NCHW layout (zeros where the reference's gradient is zero).
 * ------------------------------------------------------------------------------------- */

/* mode: 0 = 'iou' (loss.py:142-152), 1 = 'giou' (loss.py:155-177).  Reads levels[].reg. */
int b200det_box_loss_fwd(const b200det_level* levels, int n_levels, int batch,
                         const float* cnt_t, const float* reg_t, int mode,
                         float* loss, float* num_pos, void* stream);
/* grads[l] = d/d(reg level l) [B,4,h,w]. */
int b200det_box_loss_bwd(const b200det_level* levels, float* const* grads, int n_levels, int batch,
                         const float* cnt_t, const float* reg_t, int mode,
                         const float* grad_loss, const float* num_pos, void* stream);

/* Centerness BCE-with-logits over positives (loss.py:29-57).  Reads levels[].cnt;
 * cnt_target [B,P] f32 holds the BCE targets (normally the same pointer as cnt_t). */
int b200det_cnt_loss_fwd(const b200det_level* levels, int n_levels, int batch,
                         const float* cnt_t, const float* cnt_target,
                         float* loss, float* num_pos, void* stream);
int b200det_cnt_loss_bwd(const b200det_level* levels, float* const* grads, int n_levels, int batch,
                         const float* cnt_t, const float* cnt_target, const float* grad_loss,
                         const float* num_pos, void* stream);

/* Focal loss over ALL points (loss.py:6-26,180-193; alpha 0.25, gamma 2).  Reads levels[].cls.
 * workspace: b200det_cls_loss_workspace_bytes(). */
size_t b200det_cls_loss_workspace_bytes(int batch, int num_points, int num_classes);
int b200det_cls_loss_fwd(const b200det_level* levels, int n_levels, int batch, int num_classes,
                         const int64_t* cls_t, const float* cnt_t,
                         void* workspace, size_t workspace_bytes,
                         float* loss, float* num_pos, void* stream);
int b200det_cls_loss_bwd(const b200det_level* levels, float* const* grads, int n_levels, int batch,
                         int num_classes, const int64_t* cls_t,
                         const float* grad_loss, const float* num_pos, void* stream);

/* The focal loss of a TRAINING STEP: loss and gradient from ONE read of the class logits
 * (compute_cls_loss forward, loss.py:6-26, and its autograd backward; the `.mean()` of loss.py:210).
 *   cls_dtype (b200det_dtype): element type of levels[].cls AND of grads[] — fp32, or fp16 / bf16 as the
 *     class convolution leaves them under torch.cuda.amp.autocast (train.py:175); arithmetic is fp32 either way
 *     and a half gradient is rounded once (rn).  Use grad_mode 1 with a loss scale for fp16 gradients;
 *   grads[l] [B,C,h,w] receive d(sum_b grad_loss[b] * loss[b]) / d(cls level l);
 *   grad_loss: NULL for 1/B each (the gradient of the batch mean); grad_mode 0: [B] f32 device, dL/d(loss[b]);
 *     grad_mode 1: ONE f32 device value, the upstream gradient of the batch MEAN (a loss scale) -> /B each;
 *   num_pos [B] f32: read when num_pos_ready != 0 (as b200det_assign_loss_fused / *_loss_fwd wrote it),
 *     otherwise computed first from cnt_t (> -1 marks a positive) and written;
 *   loss [B] f32 = focal sum / num_pos; mean_out [2] f32 or NULL = {batch mean, added in image order; the upstream
 *   gradient the gradients were written for (grad_mode 1: *grad_loss as read by this call, else 1)}.
 * workspace: b200det_cls_loss_workspace_bytes(). */
int b200det_cls_loss_step(const b200det_level* levels, void* const* grads, int cls_dtype, int n_levels, int batch,
                          int num_classes, const int64_t* cls_t, const float* cnt_t,
                          const float* grad_loss, int grad_mode, int num_pos_ready,
                          void* workspace, size_t workspace_bytes,
                          float* loss, float* num_pos, float* mean_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * K4 fused: one launch per training step for
 *   FCOSGenTargets.forward (head.py:218-316)
 *   + compute_reg_loss forward AND its autograd backward (loss.py:116-177)
 *   + compute_cnt_loss forward AND backward (loss.py:29-57; optional)
 *   + the `.mean()` over the batch of FCOSLoss.forward (loss.py:210-213).
 * One streaming kernel (csrc/assign_stream.cuh: a fill warp writes the dense negatives with bulk copies while
 * the other warps assign in shared memory, predictions are fetched at positives only, num_pos — which scales
 * every gradient — travels through per-image arrival counters) and its programmatic dependent, a one-CTA
 * deterministic reduction of the loss partials.
 *   levels[l].reg (+ .cnt when cnt_grads != NULL), h, w, stride : head outputs (read at positives)
 *   reg_grads[l] [B,4,h,w], cnt_grads[l] [B,1,h,w] (host arrays of device pointers; cnt_grads may be
 *     NULL together with cnt_loss): receive d(sum_b grad_*[b] * loss[b]) / d(map), zeros off positives
 *   grad_box / grad_cnt: NULL for 1/B each (the gradient of the batch mean); grad_mode 0: [B] f32 device;
 *     grad_mode 1: ONE f32 device value each, the upstream gradient of the batch mean (divided by B inside)
 *   cls_t / cnt_t / reg_t : the targets, as b200det_assign_targets writes them (bit-identical)
 *   box_loss / cnt_loss / num_pos [B] f32 : as b200det_box_loss_fwd / b200det_cnt_loss_fwd
 *   mean_out [4] f32 or NULL : batch means of box_loss and cnt_loss, added in image order; then the upstream
 *     gradients of the two means the gradients were written for (grad_mode 1: *grad_box, *grad_cnt as read by THIS
 *     call; else 1) — what its backward hands to b200det_rescale_maps as `assumed`
 *   reg_scale_grad [n_levels] f32 or NULL : d(sum_b grad_box[b] * box_loss[b]) / d(*levels[l].reg_scale)
 *     (0 for levels without reg_scale); with reg_scale set, reg_grads are gradients w.r.t. the raw x
 *   workspace : b200det_assign_loss_workspace_bytes(batch, P) bytes (tile partials); one workspace per
 *     stream that may run this concurrently.
 * ------------------------------------------------------------------------------------- */
size_t b200det_assign_loss_workspace_bytes(int batch, int num_points);
int b200det_assign_loss_fused(const b200det_level* levels, float* const* reg_grads, float* const* cnt_grads,
                              int n_levels, const float* limit_lo, const float* limit_hi,
                              const float* radius_px, int batch, int max_gt,
                              const float* gt_boxes, const int64_t* gt_labels, int mode,
                              const float* grad_box, const float* grad_cnt, int grad_mode,
                              int64_t* cls_t, float* cnt_t, float* reg_t,
                              float* box_loss, float* cnt_loss, float* num_pos, float* mean_out,
                              float* reg_scale_grad, void* workspace, size_t workspace_bytes, void* stream);

/* maps[i] (device, numel[i] floats) *= *factors[i] (device scalar) for i < n_maps <= 16; the arrays
 * themselves are HOST arrays.  A map whose factor is exactly 1 is not touched.  This is the autograd
 * backward of the fused step: its gradients are final unless the upstream gradient differs from 1. */
int b200det_scale_maps(float* const* maps, const int64_t* numel, const float* const* factors, int n_maps,
                       void* stream);

/* The autograd backward of b200det_assign_loss_fused / b200det_cls_loss_step when their gradients were
 * written for an ASSUMED upstream gradient (grad_mode 1).  A `state` is TWO consecutive fp32 words on the
 * device: {upstream gradient the next forward will assume, 0} (the second word is a ticket the kernel uses and
 * resets); the forward entry points take a pointer to its first word and report the value they read in mean_out.
 * got[s] is the upstream gradient that arrived for state s, assumed[s] the value the forward THAT IS BEING
 * DIFFERENTIATED was given (its mean_out copy; NULL = the state's current word).  Map i (numel[i] elements of
 * `dtype`) belongs to state state_of[i]: if *got != *assumed the map is multiplied by *got / *assumed in place,
 * otherwise it is not touched — also when several forwards were outstanding and the shared word has moved on.
 * Afterwards the state's word = *got (unless that is 0 or not finite), so that the next forward assumes the
 * upstream gradient this backward received — under torch.cuda.amp.GradScaler (train.py:127,180) that is the loss
 * scale, constant for thousands of steps.  One launch.  maps / numel / state_of (n_maps <= 16) and got / assumed /
 * state (n_states <= 4, each state listed once; `assumed` itself may be NULL) are HOST arrays. */
int b200det_rescale_maps(void* const* maps, const int64_t* numel, const int32_t* state_of, int dtype, int n_maps,
                         const float* const* got, const float* const* assumed, float* const* state, int n_states,
                         void* stream);

/* ---------------------------------------------------------------------------------------
 * N4 — the datasets' collate_fn on the device (dataset/voc.py:141-173, dataset/coco.py:135-165).
 * ------------------------------------------------------------------------------------- */

/* Ragged GT lists -> padded batch.  flat_boxes [N,4] f32 and flat_labels [N] i64 hold the images' rows back
 * to back, offsets [batch+1] i32 (device) the row range of each image.  Writes gt_boxes [batch,max_gt,4]
 * and gt_labels [batch,max_gt], rows beyond an image's count = -1 (F.pad(..., value=-1) + torch.stack). */
int b200det_pack_gt(const float* flat_boxes, const int64_t* flat_labels, const int32_t* offsets,
                    int batch, int max_gt, float* gt_boxes, int64_t* gt_labels, void* stream);

/* images[b] (HOST array of device pointers) = image b, [channels,h_b,w_b] f32 with image_hw[2b] = h_b,
 * image_hw[2b+1] = w_b (host).  out [batch,channels,out_h,out_w] = Normalize(mean, std)(zero-pad(image)):
 * a padded pixel is (0 - mean[c]) / std[c], as in the reference, which normalises AFTER padding.
 * mean / std are host arrays of `channels` floats. */
int b200det_collate_images(const void* const* images, const int32_t* image_hw, int batch, int channels,
                           int out_h, int out_w, const float* mean, const float* std, float* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * N3 — VOC average precision on the device.  Replaces sort_by_score + eval_ap_2d + _compute_ap
 * (test.py:15-162; call site test.py:225-226) on the padded detections FCOSHead.detect() produces.
 *   det_score [batch,max_det] f32 (descending per image), det_cls [batch,max_det] i64 (1-based),
 *   det_box [batch,max_det,4] f32, det_count [batch] i32; gt_boxes [batch,max_gt,4] f32, gt_labels
 *   [batch,max_gt] i64 (padding rows: label outside 1..num_cls-1, e.g. -1).
 * `batch` is the whole evaluation set (or a rank's shard of it).  num_cls counts the background class 0,
 * as in the reference.  Writes ap [num_cls] f64: ap[0] = 0, ap[c] = average precision of class c
 * (NaN when the class has detections but no ground truth, as numpy's 0/0 gives; 0 without detections).
 * A detection is a true positive iff its best-IoU ground-truth box of the same class (first index on
 * ties) has IoU >= iou_thr and is not taken by a higher-scoring detection of the image.
 * ------------------------------------------------------------------------------------- */
size_t b200det_eval_ap_workspace_bytes(int batch, int max_det, int num_cls);
int b200det_eval_ap(int batch, int max_det, int max_gt, int num_cls,
                    const float* det_score, const int64_t* det_cls, const float* det_box,
                    const int32_t* det_count, const float* gt_boxes, const int64_t* gt_labels,
                    double iou_thr, void* workspace, size_t workspace_bytes, double* ap, void* stream);

/* COCO result rows (Test_coco.py:144-168): out_xywh [batch,max_det,4] = (x1/s, y1/s, x2/s - x1/s, y2/s - y1/s)
 * with s = scale[b] (the resize factor of image b, fp32, device), each op rounded in fp32 like numpy's
 * in-place float32 arithmetic; out_count [batch] = number of leading detections with score >= threshold
 * (the reference stops at the first score below it). */
int b200det_coco_boxes(int batch, int max_det, const float* det_box, const float* det_score,
                       const int32_t* det_count, const float* scale, float threshold,
                       float* out_xywh, int32_t* out_count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DET_H_ */
