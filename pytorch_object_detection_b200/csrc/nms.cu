// K3 — batched NMS with the arithmetic of torchvision 0.26.0's CPU path (the op the reference
// calls at model/modules/head.py:94), plus ClipBoxes (head.py:152-162) fused into the writer.
//
//   nms_prepare_kernel  (stand-alone entry only; K2 does this for the fused path)
//       score >= thr compaction (order kept), stable descending sort, class-offset boxes.
//   nms_mask_kernel     one CTA of 64 threads per 64x64 tile of the upper triangle: bit j of
//       word (i, cb) says "box i suppresses box cb*64+j".  IoU in the reference's operation
//       order with explicitly rounded fp32 intrinsics (no FMA contraction is possible):
//         inter = max(0, min(x2)-max(x1)) * max(0, min(y2)-max(y1))
//         iou   = inter / (area_i + area_j - inter),  suppressed iff (double)iou > thr.
//   nms_scan_kernel     one CTA per image: greedy pass over the mask, 64 rows at a time staged
//       into shared memory with cp.async (double-buffered); warp 0 resolves the 64x64 diagonal
//       serially, all warps OR the kept rows into the removed-bitmap, and the kept boxes are
//       written (clipped when asked) in keep order.
// This stage is latency-bound (the greedy dependency chain), not bandwidth-bound.
#include "block_utils.cuh"
#include "nms.cuh"

B200DET_TRACE_BUFFER(nms)

#include <math.h>

namespace b200det {
namespace {

constexpr int kPrepThreads = 1024;
constexpr int kScanThreads = 256;

// ------------------------------------------------------------------------------------------
// prepare (stand-alone batched_nms entry)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPrepThreads, 1)
nms_prepare_kernel(const int n, const float* __restrict__ boxes, const float* __restrict__ scores,
                   const long long* __restrict__ classes, const int32_t* __restrict__ in_count,
                   const float thr, const CandSet set) {
  extern __shared__ __align__(16) unsigned long long sortbuf[];   // [n2] keys, then [n] int source map
  __shared__ int s_scan[33];
  __shared__ float s_fmax[32];
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  int n2 = next_pow2(n);
  int* srcmap = reinterpret_cast<int*>(sortbuf + n2);
  const int n_in = in_count ? min(max(in_count[b], 0), n) : n;
  const float* sc = scores + (size_t)b * n;

  // order-preserving threshold compaction: rank = index into the thresholded list (head.py:90-93)
  int m = 0;
  for (int base = 0; base < n_in; base += kPrepThreads) {
    const int i = base + tid;
    float s = 0.f;
    const bool ok = (i < n_in) && ((s = sc[i]) >= thr);
    int total;
    const int rank = m + block_exclusive_scan(ok ? 1 : 0, s_scan, &total);
    if (ok) {
      sortbuf[rank] = ((unsigned long long)order_key(s) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)rank);
      srcmap[rank] = i;
    }
    m += total;
  }
  if (m == 0) {
    if (tid == 0) { set.count[b] = 0; set.mode[b] = 0; }
    return;
  }
  n2 = next_pow2(m);
  for (int i = m + tid; i < n2; i += kPrepThreads) sortbuf[i] = 0ull;
  __syncthreads();
  bitonic_sort_desc(sortbuf, n2);   // equal scores: lower rank first = stable descending

  const size_t o0 = (size_t)b * set.cap;
  float vmax = -CUDART_INF_F;
  for (int i = tid; i < m; i += kPrepThreads) {
    const unsigned long long e = sortbuf[i];
    const int rank = (int)(0xffffffffu - (uint32_t)(e & 0xffffffffull));
    const int src = srcmap[rank];
    const float4 bx = reinterpret_cast<const float4*>(boxes)[(size_t)b * n + src];
    set.score[o0 + i] = sc[src];
    set.cls[o0 + i] = (int)classes[(size_t)b * n + src];
    set.src[o0 + i] = rank;
    reinterpret_cast<float4*>(set.box)[o0 + i] = bx;
    vmax = fmaxf(vmax, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
  }
  vmax = block_max(vmax, s_fmax);
  nms_prepare_boxes(set, b, m, vmax, tid, kPrepThreads);
  if (tid == 0) set.count[b] = m;
}

// ------------------------------------------------------------------------------------------
// suppression mask
// ------------------------------------------------------------------------------------------
// One CTA of 64 threads per 64x64 tile: thread t owns row rb*64+t and walks the 64 staged column
// boxes.  Fast path per pair: one broadcast LDS.128, 4 min/max, 2 compares.  w > 0 <=> min(x2) >
// max(x1) exactly in IEEE arithmetic, so the reference's max(0, .) products and the division are
// only evaluated when some lane of the warp has an overlapping pair.  ZERO_SUP (thr < 0, where a
// zero IoU suppresses) takes the full expression for every pair.
template <bool ZERO_SUP>
__global__ void __launch_bounds__(kNmsTile)
nms_mask_kernel(const CandSet set, const int wcap, const float thr_up, unsigned long long* __restrict__ mask) {
  const int cb = blockIdx.x, rb = blockIdx.y, b = blockIdx.z;
  if (cb < rb) return;
  const int n = set.count[b];
  if (rb * kNmsTile >= n || cb * kNmsTile >= n) return;
  const bool same_class_only = set.mode[b] == kModeVanilla;

  __shared__ float4 cbox[kNmsTile];
  __shared__ float carea[kNmsTile];
  __shared__ int ccls[kNmsTile];
  const int t = threadIdx.x;
  const size_t o0 = (size_t)b * set.cap;
  const int cj = cb * kNmsTile + t;
  if (cj < n) {
    const float4 v = reinterpret_cast<const float4*>(set.nms_box)[o0 + cj];
    cbox[t] = v;
    carea[t] = __fmul_rn(__fsub_rn(v.z, v.x), __fsub_rn(v.w, v.y));
    ccls[t] = set.cls[o0 + cj];
  } else {
    // beyond the image's candidates: an inverted infinite box overlaps nothing (min(x2) = -inf is never
    // > max(x1) = +inf) and its NaN area keeps the IoU unordered on the thr < 0 path -> bit stays 0
    cbox[t] = make_float4(CUDART_INF_F, CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    carea[t] = __int_as_float(0x7fc00000);
    ccls[t] = -1;
  }
  __syncthreads();

  const int i = rb * kNmsTile + t;
  const bool row_ok = i < n;                       // keep whole warps in the loop (warp votes below)
  const float4 a = row_ok ? reinterpret_cast<const float4*>(set.nms_box)[o0 + i]
                          : make_float4(CUDART_INF_F, CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
  const float aarea = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const int acls = row_ok ? set.cls[o0 + i] : -2;
  // pass 1 (branch-free, fully unrolled): candidate bit j <=> the boxes overlap with positive area
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
  if (cb == rb) bits &= ~((2ull << t) - 1ull);     // diagonal tile: only later boxes (j > t)
  if (row_ok) mask[(o0 + i) * wcap + cb] = bits;
}

// ------------------------------------------------------------------------------------------
// greedy scan + output
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ float clip1(float v, float hi) { return fminf(fmaxf(v, 0.f), hi); }

__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const CandSet set, const int wcap, const unsigned long long* __restrict__ mask,
                const int clip_h, const int clip_w, const NmsOut out) {
  extern __shared__ __align__(16) unsigned long long sm[];   // [2][64*wcap] row chunks, [wcap] removed
  __shared__ unsigned long long s_keep;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int n = set.count[b];
  const int W = (n + kNmsTile - 1) / kNmsTile;
  const int Wr2 = ((W + 1) & ~1) / 2;                        // 16-byte units per row to stage
  unsigned long long* removed = sm + 2 * kNmsTile * wcap;
  const size_t o0 = (size_t)b * set.cap;
  const size_t q0 = (size_t)b * out.stride;

  for (int w = tid; w < wcap; w += kScanThreads) removed[w] = 0ull;

  auto stage = [&](int rb) {
    unsigned long long* dst = sm + (rb & 1) * kNmsTile * wcap;
    const int rows = min(kNmsTile, n - rb * kNmsTile);
    for (int e = tid; e < rows * Wr2; e += kScanThreads) {
      const int r = e / Wr2, c = e - r * Wr2;
      cp_async16(dst + r * wcap + 2 * c, mask + (o0 + rb * kNmsTile + r) * wcap + 2 * c);
    }
    cp_async_commit();
  };

  int total = 0;
  if (W > 0) stage(0);
  for (int rb = 0; rb < W; ++rb) {
    if (rb + 1 < W) {
      stage(rb + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();                                         // chunk rb (and removed[]) visible
    const unsigned long long* chunk = sm + (rb & 1) * kNmsTile * wcap;
    const int rows = min(kNmsTile, n - rb * kNmsTile);
    if (warp == 0) {
      unsigned long long cur = removed[rb], keep = 0ull;
#pragma unroll 16
      for (int i = 0; i < kNmsTile; ++i) {
        const unsigned long long d = chunk[i * wcap + rb];   // broadcast read, off the dependency chain
        const bool alive = (i < rows) && !((cur >> i) & 1ull);
        keep |= alive ? (1ull << i) : 0ull;
        cur |= alive ? d : 0ull;
      }
      if (lane == 0) s_keep = keep;
    }
    __syncthreads();
    const unsigned long long keep = s_keep;
    // removed[w] |= OR of the kept rows, for the column blocks still ahead
    for (int w = rb + 1 + lane; w < W; w += 32) {
      unsigned long long acc = 0ull;
      for (int i = warp; i < rows; i += kScanThreads / 32)
        if ((keep >> i) & 1ull) acc |= chunk[i * wcap + w];
      if (acc) atomicOr(&removed[w], acc);
    }
    // kept boxes of this block, in order
    if (tid < kNmsTile && ((keep >> tid) & 1ull)) {
      const int q = rb * kNmsTile + tid;
      const int o = total + __popcll(keep & ((1ull << tid) - 1ull));
      float4 bx = reinterpret_cast<const float4*>(set.box)[o0 + q];
      if (clip_h > 0) {   // ClipBoxes: clamp_(min=0), then x <= w-1, y <= h-1   (head.py:156-162)
        bx.x = clip1(bx.x, (float)(clip_w - 1));
        bx.y = clip1(bx.y, (float)(clip_h - 1));
        bx.z = clip1(bx.z, (float)(clip_w - 1));
        bx.w = clip1(bx.w, (float)(clip_h - 1));
      }
      out.score[q0 + o] = set.score[o0 + q];
      out.cls[q0 + o] = (long long)set.cls[o0 + q];
      out.keep[q0 + o] = (long long)set.src[o0 + q];
      reinterpret_cast<float4*>(out.box)[q0 + o] = bx;
    }
    total += __popcll(keep);
    __syncthreads();                                         // removed[] complete; chunk buffer reusable
  }
  if (tid == 0) out.count[b] = total;
}

// ---- small-n variant: the image's whole upper-triangular mask resident in shared memory ------
// (cap <= kSmemScanMaxCap, i.e. every reference configuration with max_detection_box = 1000).
// All warps stage the rows with cp.async into a padded layout (odd row stride: a lane-per-row
// column read is bank-conflict-free); then ONE warp runs the greedy pass with no block barrier
// on its critical path:
//   per 64-row block: lane i holds diagonal words of rows i and i+32.  If no still-alive row
//   suppresses another still-alive row of the block (one warp OR-reduction) every alive row is
//   kept at once; otherwise the 64 rows are resolved serially out of registers (shuffles).
//   The kept rows are then OR-reduced column by column into the removed-bitmap, which lives in
//   registers (lane w owns word w).
constexpr int kSmemScanThreads = 256;
constexpr int kSmemScanMaxCap = 1280;    // 1280 * 21 * 8 B = 210 KB

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
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

__global__ void __launch_bounds__(kSmemScanThreads, 1)
nms_scan_smem_kernel(const CandSet set, const int wcap, const unsigned long long* __restrict__ mask,
                     const int clip_h, const int clip_w, const NmsOut out) {
  extern __shared__ __align__(16) unsigned long long sm[];   // [cap][stride] mask rows, [32] keep words
  __shared__ int s_pre[33];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kSmemScanThreads / 32;
  const int b = blockIdx.x;
  B200DET_STAMP_NOSYNC(16);
  const int n = set.count[b];
  const int W = (n + kNmsTile - 1) / kNmsTile;              // <= 20
  const int stride = wcap | 1;
  unsigned long long* M = sm;
  unsigned long long* keepw = sm + (size_t)set.cap * stride;
  const size_t o0 = (size_t)b * set.cap;
  const size_t q0 = (size_t)b * out.stride;

  // stage rows [r0, r1): row i needs the words of its own block and the blocks ahead
  // A warp covers 32 / wp2 rows per step (wp2 = W rounded up to a power of two <= 32), one 8-byte
  // word per lane, so the copy loop is a handful of instructions per row.
  int wp2 = 1;
  while (wp2 < W) wp2 <<= 1;
  const int sub = lane / wp2, wl = lane & (wp2 - 1), rows_per_step = 32 / wp2;
  auto stage_rows = [&](int r0, int r1) {
    for (int i = r0 + warp * rows_per_step + sub; i < r1; i += kWarps * rows_per_step)
      if (wl >= (i >> 6) && wl < W) cp_async8(M + (size_t)i * stride + wl, mask + (o0 + i) * wcap + wl);
    cp_async_commit();
  };
  const int split = min(n, 4 * kNmsTile);
  stage_rows(0, split);
  stage_rows(split, n);
  if (tid < 32) keepw[tid] = 0ull;
  cp_async_wait<1>();
  __syncthreads();                                           // rows [0, split) resident
  B200DET_STAMP_NOSYNC(17);

  if (warp != 0) {
    cp_async_wait<0>();
    asm volatile("bar.arrive 1, %0;" ::"n"(kSmemScanThreads) : "memory");
  } else {
    unsigned long long myrem = 0ull;                         // lane w: removed-bitmap word w
    bool rest_ready = false;
    for (int rb = 0; rb < W; ++rb) {
      if (rb == 4) {
        cp_async_wait<0>();
        asm volatile("bar.sync 1, %0;" ::"n"(kSmemScanThreads) : "memory");
        rest_ready = true;
      }
      const int base = rb * kNmsTile;
      const int rows = min(kNmsTile, n - base);
      const unsigned long long valid = rows == kNmsTile ? ~0ull : ((1ull << rows) - 1ull);
      const unsigned long long cur = shfl64(myrem, rb);
      const size_t r0 = (size_t)(base + lane) * stride, r1 = (size_t)(base + lane + 32) * stride;
      const unsigned long long d0 = (lane < rows) ? M[r0 + rb] : 0ull;
      const unsigned long long d1 = (lane + 32 < rows) ? M[r1 + rb] : 0ull;
      const bool a0 = !((cur >> lane) & 1ull), a1 = !((cur >> (lane + 32)) & 1ull);
      const unsigned long long S = warp_or64((a0 ? d0 : 0ull) | (a1 ? d1 : 0ull));
      unsigned long long keep;
      if (((S & ~cur) & valid) == 0ull) {
        keep = ~cur & valid;                                 // no alive row touches another alive row
      } else {
        unsigned long long c = cur;
        keep = 0ull;
#pragma unroll 8
        for (int i = 0; i < kNmsTile; ++i) {
          const unsigned long long di = shfl64(i < 32 ? d0 : d1, i & 31);
          const bool alive = ((valid >> i) & 1ull) && !((c >> i) & 1ull);
          keep |= alive ? (1ull << i) : 0ull;
          c |= alive ? di : 0ull;
        }
      }
      if (lane == 0) keepw[rb] = keep;
      const bool k0 = (keep >> lane) & 1ull, k1 = (keep >> (lane + 32)) & 1ull;
      for (int w = rb + 1; w < W; ++w) {
        const unsigned long long v = warp_or64((k0 ? M[r0 + w] : 0ull) | (k1 ? M[r1 + w] : 0ull));
        if (lane == w) myrem |= v;
      }
    }
    if (!rest_ready) {
      cp_async_wait<0>();
      asm volatile("bar.sync 1, %0;" ::"n"(kSmemScanThreads) : "memory");
    }
  }
  __syncthreads();

  B200DET_STAMP_NOSYNC(18);
  if (warp == 0) {                                           // exclusive prefix of kept counts per block
    const int cnt = (lane < W) ? __popcll(keepw[lane]) : 0;
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    s_pre[lane] = incl - cnt;
    if (lane == 31) s_pre[32] = incl;
  }
  __syncthreads();
  // kept boxes in keep order; the gathers of all of a thread's rows are issued before any store
  constexpr int kRowsPerThread = (kSmemScanMaxCap + kSmemScanThreads - 1) / kSmemScanThreads;
  int oidx[kRowsPerThread];
  float4 bx[kRowsPerThread];
  float sc[kRowsPerThread];
  int cl[kRowsPerThread], sr[kRowsPerThread];
#pragma unroll
  for (int u = 0; u < kRowsPerThread; ++u) {
    const int q = tid + u * kSmemScanThreads;
    oidx[u] = -1;
    if (q < n) {
      const unsigned long long kw = keepw[q >> 6];
      if ((kw >> (q & 63)) & 1ull) {
        oidx[u] = s_pre[q >> 6] + __popcll(kw & ((1ull << (q & 63)) - 1ull));
        bx[u] = reinterpret_cast<const float4*>(set.box)[o0 + q];
        sc[u] = set.score[o0 + q];
        cl[u] = set.cls[o0 + q];
        sr[u] = set.src[o0 + q];
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kRowsPerThread; ++u) {
    if (oidx[u] < 0) continue;
    float4 v = bx[u];
    if (clip_h > 0) {   // ClipBoxes: clamp_(min=0), then x <= w-1, y <= h-1   (head.py:156-162)
      v.x = clip1(v.x, (float)(clip_w - 1));
      v.y = clip1(v.y, (float)(clip_h - 1));
      v.z = clip1(v.z, (float)(clip_w - 1));
      v.w = clip1(v.w, (float)(clip_h - 1));
    }
    const size_t o = q0 + oidx[u];
    out.score[o] = sc[u];
    out.cls[o] = (long long)cl[u];
    out.keep[o] = (long long)sr[u];
    reinterpret_cast<float4*>(out.box)[o] = v;
  }
  B200DET_STAMP(19);
  if (tid == 0) out.count[b] = s_pre[32];
}

struct NmsWorkspace {
  CandSet set;
  unsigned long long* mask;
  size_t bytes;
};

// carve the candidate set + mask out of a workspace (base may be null to size only)
NmsWorkspace carve_nms(void* base, int batch, int cap) {
  NmsWorkspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? static_cast<char*>(base) + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const size_t bc = (size_t)batch * cap;
  w.set.score = reinterpret_cast<float*>(take(bc * 4));
  w.set.cls = reinterpret_cast<int32_t*>(take(bc * 4));
  w.set.box = reinterpret_cast<float*>(take(bc * 16));
  w.set.src = reinterpret_cast<int32_t*>(take(bc * 4));
  w.set.nms_box = reinterpret_cast<float*>(take(bc * 16));
  w.set.count = reinterpret_cast<int32_t*>(take((size_t)batch * 4));
  w.set.mode = reinterpret_cast<int32_t*>(take((size_t)batch * 4));
  w.set.cap = cap;
  w.mask = reinterpret_cast<unsigned long long*>(take(bc * nms_mask_words(cap) * 8));
  w.bytes = off;
  return w;
}

}  // namespace

size_t nms_set_workspace_bytes(int batch, int cap) { return carve_nms(nullptr, batch, cap).bytes; }

void nms_set_carve(void* base, int batch, int cap, CandSet* set, unsigned long long** mask) {
  NmsWorkspace w = carve_nms(base, batch, cap);
  *set = w.set;
  *mask = w.mask;
}

int launch_nms(const CandSet& set, int batch, double nms_thr, int clip_h, int clip_w,
               unsigned long long* mask, const NmsOut& out, cudaStream_t stream) {
  const int wcap = nms_mask_words(set.cap);
  // smallest fp32 F with (double)F > thr: for a non-NaN fp32 iou, (double)iou > thr  <=>  iou >= F
  float thr_up = (float)nms_thr;
  if (!((double)thr_up > nms_thr)) thr_up = nextafterf(thr_up, INFINITY);
  const bool zero_suppresses = !(nms_thr >= 0.0);
  const int wblocks = (set.cap + kNmsTile - 1) / kNmsTile;
  if (zero_suppresses)
    nms_mask_kernel<true><<<dim3(wblocks, wblocks, batch), kNmsTile, 0, stream>>>(set, wcap, thr_up, mask);
  else
    nms_mask_kernel<false><<<dim3(wblocks, wblocks, batch), kNmsTile, 0, stream>>>(set, wcap, thr_up, mask);
  int rc = check_launch();
  if (rc) return rc;
  if (set.cap <= kSmemScanMaxCap) {
    const size_t smem = ((size_t)set.cap * (wcap | 1) + 32) * sizeof(unsigned long long);
    if (smem > 48 * 1024) {
      cudaError_t e =
          cudaFuncSetAttribute(nms_scan_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
    }
    nms_scan_smem_kernel<<<batch, kSmemScanThreads, smem, stream>>>(set, wcap, mask, clip_h, clip_w, out);
    return check_launch();
  }
  const size_t smem = ((size_t)2 * kNmsTile * wcap + wcap) * sizeof(unsigned long long);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
  }
  nms_scan_kernel<<<batch, kScanThreads, smem, stream>>>(set, wcap, mask, clip_h, clip_w, out);
  return check_launch();
}

}  // namespace b200det

extern "C" size_t b200det_nms_workspace_bytes(int batch, int n) {
  if (batch <= 0 || n <= 0 || n > B200DET_MAX_BOX) return 0;
  return b200det::nms_set_workspace_bytes(batch, n);
}

extern "C" int b200det_batched_nms(int batch, int n, const float* boxes, const float* scores,
                                   const int64_t* classes, const int32_t* in_count, float score_thr,
                                   double nms_thr, int clip_h, int clip_w, void* workspace,
                                   size_t workspace_bytes, float* out_score, int64_t* out_cls, float* out_box,
                                   int64_t* out_keep, int32_t* out_count, void* stream) {
  using namespace b200det;
  if (batch <= 0 || batch > 65535 || n <= 0 || !boxes || !scores || !classes || !workspace || !out_score ||
      !out_cls || !out_box || !out_keep || !out_count)
    return B200DET_ERR_ARG;
  if (n > B200DET_MAX_BOX) return B200DET_ERR_UNSUPPORTED;
  if (!aligned16(boxes) || !aligned16(out_box) || !aligned16(workspace)) return B200DET_ERR_ARG;
  if (workspace_bytes < nms_set_workspace_bytes(batch, n)) return B200DET_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CandSet set;
  unsigned long long* mask;
  nms_set_carve(workspace, batch, n, &set, &mask);
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  const size_t smem = (size_t)n2 * 8 + (size_t)n * 4;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(nms_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
  }
  nms_prepare_kernel<<<batch, kPrepThreads, smem, st>>>(n, boxes, scores, reinterpret_cast<const long long*>(classes),
                                                       in_count, score_thr, set);
  int rc = check_launch();
  if (rc) return rc;
  NmsOut out{out_score, reinterpret_cast<long long*>(out_cls), out_box, reinterpret_cast<long long*>(out_keep),
             out_count, n};
  return launch_nms(set, batch, nms_thr, clip_h, clip_w, mask, out, st);
}
