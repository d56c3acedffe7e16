// NMS device functions shared by nms.cu (three-kernel path) and fused.cu (cluster kernel).
#pragma once
#include "nms.cuh"

namespace b200det {

// One row of a 64x64 mask tile: box `a` against the 64 column boxes staged in shared memory
// (cbox / carea / ccls, 64 entries).  Pass 1 is branch-free and fully unrolled (one broadcast
// LDS.128, 4 min/max, 2 compares per pair): w > 0 <=> min(x2) > max(x1) exactly in IEEE
// arithmetic, so only overlapping pairs become candidates.  Pass 2 evaluates the reference's exact
// IoU expression (torchvision nms_kernel_impl: fp32, one rounding per operation, no FMA) for the
// candidates.  ZERO_SUP (thr < 0, where a zero IoU suppresses) makes every pair a candidate.
template <bool ZERO_SUP>
__device__ __forceinline__ unsigned long long mask_row_bits(const float4 a, const float aarea, const int acls,
                                                            const float4* cbox, const float* carea, const int* ccls,
                                                            const float thr_up, const bool same_class_only) {
  // pass 1: candidate bit j <=> the boxes overlap with positive area
  const unsigned cbase = (unsigned)__cvta_generic_to_shared(cbox);
  unsigned lo = 0u, hi = 0u;
#pragma unroll
  for (int j = 0; j < kNmsTile; ++j) {
    float4 c;                                       // same address in every lane: broadcast LDS.128
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "r"(cbase + j * 16));
    const bool overlap = ZERO_SUP || (fminf(a.z, c.z) > fmaxf(a.x, c.x) && fminf(a.w, c.w) > fmaxf(a.y, c.y));
    if (j < 32) lo |= overlap ? (1u << j) : 0u;
    else hi |= overlap ? (1u << (j - 32)) : 0u;
  }
  unsigned long long bits = ((unsigned long long)hi << 32) | lo;
  // pass 2 (rare): the reference's exact IoU expression for the candidates only
  for (unsigned long long m = bits; m; m &= m - 1ull) {
    const int j = __ffsll((long long)m) - 1;
    const float4 c = cbox[j];
    const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, c.z), fmaxf(a.x, c.x)));
    const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, c.w), fmaxf(a.y, c.y)));
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, carea[j]), inter));
    bool sup = ovr >= thr_up;                        // == (double)ovr > thr, see launch_nms
    if (same_class_only) sup = sup && (ccls[j] == acls);
    if (!sup) bits &= ~(1ull << j);
  }
  return bits;
}

// The same for the first NC (<= 32) staged columns only: the work unit of the per-class kernel's cooperative tiles
// (a 64 x 64 tile split into 64 x 16 parts, so that a class of a few hundred boxes keeps all 32 warps of the CTA busy).
template <bool ZERO_SUP, int NC>
__device__ __forceinline__ unsigned mask_row_bits_part(const float4 a, const float aarea, const float4* cbox,
                                                       const float* carea, const float thr_up) {
  const unsigned cbase = (unsigned)__cvta_generic_to_shared(cbox);
  unsigned bits = 0u;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    float4 c;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "r"(cbase + j * 16));
    const bool overlap = ZERO_SUP || (fminf(a.z, c.z) > fmaxf(a.x, c.x) && fminf(a.w, c.w) > fmaxf(a.y, c.y));
    bits |= overlap ? (1u << j) : 0u;
  }
  for (unsigned m = bits; m; m &= m - 1u) {
    const int j = __ffs((int)m) - 1;
    const float4 c = cbox[j];
    const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, c.z), fmaxf(a.x, c.x)));
    const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, c.w), fmaxf(a.y, c.y)));
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, carea[j]), inter));
    if (!(ovr >= thr_up)) bits &= ~(1u << j);            // == !((double)ovr > thr), see launch_nms
  }
  return bits;
}

// column-block-major packed upper triangle: column block w holds the words of rows [0, (w+1)*64)
__device__ __forceinline__ int col_off(int w) { return kNmsTile * (w * (w + 1) / 2); }   // words before column w

__device__ __forceinline__ float clip1(float v, float hi) { return fminf(fmaxf(v, 0.f), hi); }

__device__ __forceinline__ unsigned long long warp_or64(unsigned long long v) {
  const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
  const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) {
  const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src);
  const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
  return ((unsigned long long)hi << 32) | lo;
}

// Resolve one 64-row block given the already-removed bits `cur` and the block's diagonal words
// held lane-per-row (d0: row lane, d1: row lane+32).  If no still-alive row suppresses another
// still-alive row (one warp OR-reduction) all alive rows are kept at once; otherwise the 64 rows
// are walked serially out of registers (shuffles), the greedy rule of torchvision's nms kernel.
__device__ __forceinline__ unsigned long long resolve_block(const unsigned long long cur, const unsigned long long valid,
                                                            const unsigned long long d0, const unsigned long long d1,
                                                            const int lane) {
  const bool a0 = !((cur >> lane) & 1ull), a1 = !((cur >> (lane + 32)) & 1ull);
  const unsigned long long S = warp_or64((a0 ? d0 : 0ull) | (a1 ? d1 : 0ull));
  if (((S & ~cur) & valid) == 0ull) return ~cur & valid;
  // (Walking only the kept rows — next = lowest row neither removed nor visited — was measured and is SLOWER: its
  // shuffle depends on the previous step, ~40 cycles each, while here the 64 row fetches are independent of the
  // chain and pipeline; 1 000 crowded candidates x 16 images: 80.6 vs 68.6 us.)
  unsigned long long c = cur, keep = 0ull;
#pragma unroll 8
  for (int i = 0; i < kNmsTile; ++i) {
    const unsigned long long di = shfl64(i < 32 ? d0 : d1, i & 31);
    const bool alive = ((valid >> i) & 1ull) && !((c >> i) & 1ull);
    keep |= alive ? (1ull << i) : 0ull;
    c |= alive ? di : 0ull;
  }
  return keep;
}

// gather + (optional) clip + store of one kept row
__device__ __forceinline__ void store_kept(const CandSet& set, const NmsOut& out, const size_t o0, const size_t q0,
                                           const int q, const int o, const int clip_h, const int clip_w,
                                           const float4 bx, const float sc, const int cl, const int sr) {
  float4 v = bx;
  if (clip_h > 0) {   // ClipBoxes: clamp_(min=0), then x <= w-1, y <= h-1   (head.py:156-162)
    v.x = clip1(v.x, (float)(clip_w - 1));
    v.y = clip1(v.y, (float)(clip_h - 1));
    v.z = clip1(v.z, (float)(clip_w - 1));
    v.w = clip1(v.w, (float)(clip_h - 1));
  }
  out.score[q0 + o] = sc;
  out.cls[q0 + o] = (long long)cl;
  out.keep[q0 + o] = (long long)sr;
  reinterpret_cast<float4*>(out.box)[q0 + o] = v;
}


}  // namespace b200det
