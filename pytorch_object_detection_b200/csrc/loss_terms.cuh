// Per-element loss terms and the block reduction shared by loss.cu and train_fused.cu.
//   box  : iou_loss / giou_loss on (l, t, r, b) offsets     model/loss.py:142-177
//   cnt  : BCE-with-logits                                  model/loss.py:29-57
#pragma once
#include "common.cuh"

namespace b200det {

__device__ __forceinline__ float block_sum_f(float v, float* scratch /*32*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = (lane < nwarps) ? scratch[lane] : 0.f;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) r += __shfl_xor_sync(0xffffffffu, r, d);
  return r;
}

// ---- box regression term -------------------------------------------------------------------
// p, t = (l, t, r, b) offsets.  mode 0: -log(clamp(iou, 1e-6)); mode 1: 1 - giou.
// min_g(a, b): d min(a,b)/da with torch's tie rule.
__device__ __forceinline__ float dmin(float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); }
__device__ __forceinline__ float dmax(float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); }

template <bool GRAD>
__device__ __forceinline__ float box_term(const float4 p, const float4 t, const int mode, float4* grad) {
  const float w_pre = fminf(p.z, t.z) + fminf(p.x, t.x);
  const float h_pre = fminf(p.w, t.w) + fminf(p.y, t.y);
  const float wm = fmaxf(w_pre, 0.f), hm = fmaxf(h_pre, 0.f);
  const float O = wm * hm;
  const float pw = p.z + p.x, ph = p.w + p.y;
  const float a1 = pw * ph;
  const float a2 = (t.z + t.x) * (t.w + t.y);
  const float U = a1 + a2 - O;
  const float iou = O / U;
  float loss, g_O, g_a1;            // g_* = d loss / d *
  float g_wM = 0.f, g_hM = 0.f, W_pre = 0.f, H_pre = 0.f;
  if (mode == 0) {
    const float c = fmaxf(iou, 1e-6f);
    loss = -logf(c);
    if (GRAD) {
      const float g_iou = (iou >= 1e-6f) ? -1.f / iou : 0.f;
      const float g_U = -g_iou * O / (U * U);
      g_O = g_iou / U - g_U;
      g_a1 = g_U;
    }
  } else {
    W_pre = fmaxf(p.z, t.z) + fmaxf(p.x, t.x);
    H_pre = fmaxf(p.w, t.w) + fmaxf(p.y, t.y);
    const float wM = fmaxf(W_pre, 0.f), hM = fmaxf(H_pre, 0.f);
    const float G = wM * hM;
    const float Gc = fmaxf(G, 1e-10f);
    const float giou = iou - (G - U) / Gc;
    loss = 1.f - giou;
    if (GRAD) {
      // loss = 1 - iou + (G - U)/Gc
      const float g_iou = -1.f;
      const float g_num = 1.f / Gc;
      const float g_Gc = -(G - U) / (Gc * Gc);
      const float g_G = g_num + ((G >= 1e-10f) ? g_Gc : 0.f);
      // d loss/dU = d(-iou)/dU + d((G-U)/Gc)/dU = O/U^2 - 1/Gc
      const float gU = O / (U * U) - g_num;
      g_O = g_iou / U - gU;
      g_a1 = gU;
      g_wM = g_G * hM;
      g_hM = g_G * wM;
    }
  }
  if (GRAD) {
    const float g_wm = (w_pre >= 0.f) ? g_O * hm : 0.f;
    const float g_hm = (h_pre >= 0.f) ? g_O * wm : 0.f;
    const float g_wMp = (W_pre >= 0.f) ? g_wM : 0.f;
    const float g_hMp = (H_pre >= 0.f) ? g_hM : 0.f;
    grad->x = g_wm * dmin(p.x, t.x) + g_wMp * dmax(p.x, t.x) + g_a1 * ph;   // l
    grad->z = g_wm * dmin(p.z, t.z) + g_wMp * dmax(p.z, t.z) + g_a1 * ph;   // r
    grad->y = g_hm * dmin(p.y, t.y) + g_hMp * dmax(p.y, t.y) + g_a1 * pw;   // t
    grad->w = g_hm * dmin(p.w, t.w) + g_hMp * dmax(p.w, t.w) + g_a1 * pw;   // b
  }
  return loss;
}

// BCE-with-logits: (1 - z) * x + softplus(-x)
__device__ __forceinline__ float bce_term(float x, float z) {
  return (1.f - z) * x + (fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x))));
}

}  // namespace b200det
