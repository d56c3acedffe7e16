// NMS device functions shared by nms.cu (three-kernel path) and fused.cu (cluster kernel).
#pragma once
#include "nms.cuh"

namespace b200det {

// The reference's test `fl(inter / uni) >= thr_up` without the division (an IEEE fp32 divide is ~18 instructions
// with a slow path; this is two conversions, one DMUL and one DSETP).  Rounding is monotone, so the rounded quotient
// reaches thr_up exactly when the real quotient lies above the midpoint `mid` of thr_up and the float below it
// (25 significant bits, exact as a double), i.e. when inter > mid * uni — both sides exact in double arithmetic
// (24 + 25 <= 53 bits).  The tie inter == mid * uni cannot happen for a normal thr_up: mid is an odd 25-bit number, so
// the product has at least 25 significant bits and is no float; 0 / 0, infinities and NaN compare false on both sides.
// (Thresholds >= 0 only: ZERO_SUP keeps the division.  thr_up = the smallest float above the double threshold,
// nms_threshold_params.)
__device__ __forceinline__ double iou_midpoint(const float thr_up) {
  const float below = __int_as_float(__float_as_int(thr_up) - 1);      // thr_up > 0
  return 0.5 * ((double)below + (double)thr_up);
}
__device__ __forceinline__ bool iou_reaches(const float inter, const float uni, const double mid) {
  return (double)inter > mid * (double)uni;
}

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
  const double mid = ZERO_SUP ? 0.0 : iou_midpoint(thr_up);
  for (unsigned long long m = bits; m; m &= m - 1ull) {
    const int j = __ffsll((long long)m) - 1;
    const float4 c = cbox[j];
    const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, c.z), fmaxf(a.x, c.x)));
    const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, c.w), fmaxf(a.y, c.y)));
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(aarea, carea[j]), inter);
    bool sup = ZERO_SUP ? __fdiv_rn(inter, uni) >= thr_up : iou_reaches(inter, uni, mid);   // == (double)ovr > thr, see launch_nms
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
  const double mid = ZERO_SUP ? 0.0 : iou_midpoint(thr_up);
  auto reaches = [&](const int j) {
    const float4 c = cbox[j];
    const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, c.z), fmaxf(a.x, c.x)));
    const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, c.w), fmaxf(a.y, c.y)));
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(aarea, carea[j]), inter);
    return ZERO_SUP ? __fdiv_rn(inter, uni) >= thr_up : iou_reaches(inter, uni, mid);   // == (double)ovr > thr, see launch_nms
  };
  // A crowd (most pairs of a class overlap): every column, branch-free and unrolled — the candidate loop below would
  // run its serial chain (find bit, load, test) as many times as the busiest lane has candidates.
  // (Called by whole warps: the vote is over all 32 lanes.)
  if (__any_sync(0xffffffffu, __popc(bits) > NC / 4)) {
    unsigned sup = 0u;
#pragma unroll
    for (int j = 0; j < NC; ++j) sup |= reaches(j) ? (1u << j) : 0u;
    return bits & sup;
  }
  for (unsigned m = bits; m; m &= m - 1u) {
    const int j = __ffs((int)m) - 1;
    if (!reaches(j)) bits &= ~(1u << j);
  }
  return bits;
}

// Two rows (a0, a1) against the first NC staged columns in ONE branch-free pass, for crowds where most pairs of a
// class overlap (the per-class kernel's work unit).  The division is bracketed instead of evaluated: with uni > 0,
//   inter >  fl(thr_up * uni)  =>  inter >= thr_up * uni  =>  fl(inter / uni) >= thr_up        (suppressed)
//   inter <  fl(below * uni)   =>  inter <= below * uni   =>  fl(inter / uni) <= below < thr_up (not suppressed)
// (a float above / below a rounded product is not below / above the real product; `below` = the float under thr_up),
// which leaves a band one or two ulps wide — about one pair in 10^7 — to the exact test iou_reaches().
// 19 instructions per pair (the column loads are shared by the two rows) against ~30 for candidate pass + exact pass.
// Rows that overlap nothing (kNoBox) yield no bits; !TWO_ROWS skips a1.  Thresholds >= 0 only.
template <int NC, bool TWO_ROWS>
__device__ __forceinline__ void mask_rows2_part(const float4 a0, const float4 a1, const float4* cbox, const float* carea,
                                                const float thr_up, unsigned& bits0, unsigned& bits1) {
  const unsigned cbase = (unsigned)__cvta_generic_to_shared(cbox);
  const float below = __int_as_float(__float_as_int(thr_up) - 1);
  const float area0 = __fmul_rn(__fsub_rn(a0.z, a0.x), __fsub_rn(a0.w, a0.y));
  const float area1 = __fmul_rn(__fsub_rn(a1.z, a1.x), __fsub_rn(a1.w, a1.y));
  unsigned s0 = 0u, s1 = 0u, m0 = 0u, m1 = 0u;                 // sure / maybe bits of the two rows
  auto pair = [&](const float4 a, const float aarea, const float4 c, const float ca, const unsigned bit, unsigned& sure,
                  unsigned& maybe) {
    const float xx1 = fmaxf(a.x, c.x), yy1 = fmaxf(a.y, c.y), xx2 = fminf(a.z, c.z), yy2 = fminf(a.w, c.w);
    const bool overlap = xx2 > xx1 && yy2 > yy1;               // both boxes proper, uni > 0
    const float inter = __fmul_rn(__fsub_rn(xx2, xx1), __fsub_rn(yy2, yy1));   // (= the reference's clamped w * h here)
    const float uni = __fsub_rn(__fadd_rn(aarea, ca), inter);
    const bool yes = overlap && inter > __fmul_rn(thr_up, uni);
    const bool open = overlap && inter >= __fmul_rn(below, uni);
    sure |= yes ? bit : 0u;
    maybe |= open ? bit : 0u;
  };
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    float4 c;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "r"(cbase + j * 16));
    const float ca = carea[j];
    pair(a0, area0, c, ca, 1u << j, s0, m0);
    if (TWO_ROWS) pair(a1, area1, c, ca, 1u << j, s1, m1);
  }
  m0 &= ~s0;
  m1 &= ~s1;
  if (__any_sync(0xffffffffu, (m0 | m1) != 0u)) {              // the undecided band: the exact test
    const double mid = iou_midpoint(thr_up);
    auto settle = [&](const float4 a, const float aarea, unsigned m, unsigned& sure) {
      for (; m; m &= m - 1u) {
        const int j = __ffs((int)m) - 1;
        const float4 c = cbox[j];
        const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, c.z), fmaxf(a.x, c.x)));
        const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, c.w), fmaxf(a.y, c.y)));
        const float inter = __fmul_rn(w, h);
        if (iou_reaches(inter, __fsub_rn(__fadd_rn(aarea, carea[j]), inter), mid)) sure |= 1u << j;
      }
    };
    settle(a0, area0, m0, s0);
    settle(a1, area1, m1, s1);
  }
  bits0 = s0;
  bits1 = s1;
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
//
// The walk is a dependency chain, so it is written for the shortest one: row i only matters through the boxes AFTER
// it (bits > i; lower bits are decided already), so bit i of the removed set is final when step i reads it and the
// kept rows are simply the valid rows not removed at the end; rows past `valid` are zeroed before the walk; rows
// 32-63 only touch the high word.  What is left per step is one bit test, one select and one OR (~15 cycles; the
// shuffles that fetch the rows do not depend on the chain and run ahead) — the straightforward loop (alive flag from
// valid and removed, select, OR, keep mask) was four dependent instructions and ~39 cycles per row, 1.3 us per block.
// (Walking only the kept rows — next = lowest row neither removed nor visited — was measured and is SLOWER: its
// shuffle depends on the previous step; 1 000 crowded candidates x 16 images: 80.6 vs 68.6 us.)
__device__ __forceinline__ unsigned long long resolve_block(const unsigned long long cur, const unsigned long long valid,
                                                            unsigned long long d0, unsigned long long d1, const int lane) {
  d0 = ((valid >> lane) & 1ull) ? d0 & ~((2ull << lane) - 1ull) : 0ull;           // later boxes of valid rows only
  d1 = ((valid >> (lane + 32)) & 1ull) ? d1 & ~((2ull << (lane + 32)) - 1ull) : 0ull;
  const bool a0 = !((cur >> lane) & 1ull), a1 = !((cur >> (lane + 32)) & 1ull);
  const unsigned long long S = warp_or64((a0 ? d0 : 0ull) | (a1 ? d1 : 0ull));
  if (((S & ~cur) & valid) == 0ull) return ~cur & valid;
  unsigned c_lo = (unsigned)cur, c_hi = (unsigned)(cur >> 32);
  const unsigned d0_lo = (unsigned)d0, d0_hi = (unsigned)(d0 >> 32), d1_hi = (unsigned)(d1 >> 32);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const unsigned r_lo = __shfl_sync(0xffffffffu, d0_lo, i), r_hi = __shfl_sync(0xffffffffu, d0_hi, i);
    if (!(c_lo & (1u << i))) {
      c_lo |= r_lo;
      c_hi |= r_hi;
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const unsigned r_hi = __shfl_sync(0xffffffffu, d1_hi, i);
    if (!(c_hi & (1u << i))) c_hi |= r_hi;
  }
  return ~(((unsigned long long)c_hi << 32) | c_lo) & valid;
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
