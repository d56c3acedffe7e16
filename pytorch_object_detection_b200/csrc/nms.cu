// K3 — batched NMS with the arithmetic of torchvision 0.26.0's CPU path (the op the reference
// calls at model/modules/head.py:94), plus ClipBoxes (head.py:152-162) fused into the writer.
//
//   nms_prepare_kernel  (stand-alone entry only; K2 does this for the fused path)
//       score >= thr compaction (order kept), stable descending sort, class-offset boxes.
//   nms_mask_kernel     one CTA of 64 threads per 64x64 tile of the upper triangle: bit j of
//       word (cb, i) says "box i suppresses box cb*64+j".  IoU in the reference's operation
//       order with explicitly rounded fp32 intrinsics (no FMA contraction is possible):
//         inter = max(0, min(x2)-max(x1)) * max(0, min(y2)-max(y1))
//         iou   = inter / (area_i + area_j - inter),  suppressed iff (double)iou > thr.
//       The mask is stored COLUMN-BLOCK major, maskT[b][cb][i]: a tile's 64 words are one
//       contiguous 512-byte store, and the scan reads 64 rows of a column as consecutive words.
//   nms_scan_*_kernel   one CTA per image: the greedy pass.  Mask columns are brought into shared
//       memory by the bulk-copy engine (cp.async.bulk + mbarrier transaction counts); one warp
//       runs the dependency chain with lane = row, warp OR-reductions and the removed-bitmap in
//       registers / shared memory; kept boxes are written (clipped when asked) in keep order.
// This stage is latency-bound (the greedy dependency chain), not bandwidth-bound.
#include "block_utils.cuh"
#include "nms_body.cuh"
#include "tma.cuh"

B200DET_TRACE_BUFFER(nms)

#include <math.h>
#include <stdlib.h>

namespace b200det {
namespace {

constexpr int kPrepThreads = 1024;
static_assert(kPrepThreads == kHistThreads, "the score histogram is scanned by 1024 threads");
constexpr int kPrepParts = 8;          // CTAs per image: each gathers a slice of the sorted order
constexpr int kPrepMaxBin = 64;        // more entries than this in one score bin: the bitonic network sorts instead
// byte offset of the score histogram behind the keys [n2] and the source map [n] (16-byte aligned)
__host__ __device__ inline size_t prep_hist_offset(int n, int n2) { return ((size_t)n2 * 8 + (size_t)n * 4 + 15) / 16 * 16; }

// ------------------------------------------------------------------------------------------
// prepare (stand-alone batched_nms entry)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPrepThreads, 1)
nms_prepare_kernel(const int n, const float* __restrict__ boxes, const float* __restrict__ scores,
                   const long long* __restrict__ classes, const int32_t* __restrict__ in_count,
                   const float thr, const CandSet set) {
  // shared memory: [n2] 64-bit keys | [n] int source map | [kHistBins] score histogram | [n] keys (radix scratch)
  extern __shared__ __align__(16) unsigned long long sortbuf[];
  __shared__ int s_scan[33];
  __shared__ float s_fmax[32];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  int n2 = next_pow2(n);
  int* srcmap = reinterpret_cast<int*>(sortbuf + n2);
  unsigned* hist = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(sortbuf) + prep_hist_offset(n, n2));
  unsigned long long* tmp = reinterpret_cast<unsigned long long*>(hist + kHistBins);
  const int n_in = in_count ? min(max(in_count[b], 0), n) : n;
  const float* sc = scores + (size_t)b * n;
  B200DET_STAMP(30);

  {
    uint4* h4 = reinterpret_cast<uint4*>(hist);
#pragma unroll
    for (int q = 0; q < kHistPerThread / 4; ++q) h4[tid + q * kPrepThreads] = make_uint4(0u, 0u, 0u, 0u);
  }
  // order-preserving threshold compaction: rank = index into the thresholded list (head.py:90-93).  Warp w owns a
  // contiguous run of 32-score groups, so the rank is (candidates of the warps before) + (of the warp's groups before)
  // + (of the lanes before): ballots and ONE block scan, with all the score loads in flight together.  The scores are
  // counted into the histogram on the way (the barriers of the scan order the zero-fill before the first count).
  constexpr int kMaxGroups = B200DET_MAX_BOX / kPrepThreads;     // 8 groups of 32 scores per warp
  const int lane = tid & 31, warp = tid >> 5;
  const int groups = (n_in + kPrepThreads - 1) / kPrepThreads;
  const int i0 = warp * groups * 32 + lane;
  float sv[kMaxGroups];
#pragma unroll
  for (int q = 0; q < kMaxGroups; ++q) sv[q] = (q < groups && i0 + q * 32 < n_in) ? sc[i0 + q * 32] : 0.f;
  unsigned votes[kMaxGroups];
  int mine = 0;
#pragma unroll
  for (int q = 0; q < kMaxGroups; ++q) {
    votes[q] = __ballot_sync(0xffffffffu, q < groups && i0 + q * 32 < n_in && sv[q] >= thr);
    mine += __popc(votes[q]);
  }
  int m;
  int run = __shfl_sync(0xffffffffu, block_exclusive_scan(lane == 0 ? mine : 0, s_scan, &m), 0);
#pragma unroll
  for (int q = 0; q < kMaxGroups; ++q) {
    if ((votes[q] >> lane) & 1u) {
      const int rank = run + __popc(votes[q] & ((1u << lane) - 1u));
      const uint32_t key = order_key(sv[q]);
      sortbuf[rank] = ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - (uint32_t)rank);
      srcmap[rank] = i0 + q * 32;
      atomicAdd(&hist[hist_slot(hist_bin(key))], 1u);
    }
    run += __popc(votes[q]);
  }
  if (m == 0) {
    if (tid == 0 && blockIdx.x == 0) { set.count[b] = 0; set.mode[b] = 0; }
    return;
  }
  __syncthreads();
  B200DET_STAMP_NOSYNC(31);
  // Stable descending sort (equal scores: lower rank first).  One-pass radix sort on the score bin: the scan gives every
  // bin's slot range in descending bin order, the keys take a slot of their bin with one atomic, the few keys that share
  // a bin are ranked against each other.  5 000 candidates: ~5 us instead of 88 us for the 8 192-key bitonic network,
  // which stays as the fallback for overfull bins (many equal scores).
  bool sorted = false;
  {
    const int owner = kPrepThreads - 1 - tid;                    // bins [16 owner, 16 owner + 16): thread 0 the highest
    unsigned hb[kHistPerThread];
    int mine = 0, biggest = 0;
#pragma unroll
    for (int q = 0; q < kHistPerThread; ++q) {
      hb[q] = hist[q * kPrepThreads + owner];
      mine += (int)hb[q];
      biggest = max(biggest, (int)hb[q]);
    }
    int total;
    unsigned start = (unsigned)block_exclusive_scan(mine, s_scan, &total);
    if (block_max((float)biggest, s_fmax) <= (float)kPrepMaxBin) {
#pragma unroll
      for (int q = kHistPerThread - 1; q >= 0; --q) {              // counts -> first slot of the bin
        const unsigned c = hb[q];
        hist[q * kPrepThreads + owner] = start;
        start += c;
      }
      __syncthreads();
      B200DET_STAMP_NOSYNC(32);
      for (int i = tid; i < m; i += kPrepThreads) {
        const unsigned long long e = sortbuf[i];
        tmp[atomicAdd(&hist[hist_slot(hist_bin((uint32_t)(e >> 32)))], 1u)] = e;
      }
      __syncthreads();
      B200DET_STAMP_NOSYNC(33);
      for (int i = tid; i < m; i += kPrepThreads) {
        const unsigned long long e = tmp[i];
        const int bin = hist_bin((uint32_t)(e >> 32));
        const unsigned first = bin == kHistBins - 1 ? 0u : hist[hist_slot(bin + 1)];   // = slots of all higher bins
        const unsigned end = hist[hist_slot(bin)];
        unsigned rank = 0;
        for (unsigned j = first; j < end; ++j) rank += tmp[j] > e ? 1u : 0u;
        sortbuf[first + rank] = e;
      }
      __syncthreads();
      B200DET_STAMP_NOSYNC(34);
      sorted = true;
    }
  }
  if (!sorted) {
    n2 = next_pow2(m);
    for (int i = m + tid; i < n2; i += kPrepThreads) sortbuf[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc(sortbuf, n2);   // equal scores: lower rank first = stable descending
  }

  // Gather in score order.  Every candidate costs three scattered reads (box, score, class), and one SM's L1 takes a
  // scattered warp load a line per cycle: 15 000 lines = 7.6 us of the 21 us this kernel took on 5 000 candidates with
  // one CTA per image.  So an image has kPrepParts CTAs; each sorts its own copy of the keys (like the per-class
  // kernel: not worth distributing) and gathers one slice of the order.  Four candidates per thread at a time, so
  // that their loads overlap.  The NMS copy of the boxes (nms_prepare_boxes) is written from the registers: as they
  // are for a vanilla image; class-offset for a coordinate-trick image, which needs the largest coordinate of the
  // whole image and has at most 1 000 candidates — CTA 0 gathers those alone, one per thread.
  const size_t o0 = (size_t)b * set.cap;
  const int mode = (m * 4 <= kTrickMaxNumel) ? kModeTrick : kModeVanilla;
  if (mode == kModeTrick && blockIdx.x != 0) return;
  const int slice = mode == kModeTrick ? m : (m + (int)gridDim.x - 1) / (int)gridDim.x;
  const int lo = mode == kModeTrick ? 0 : (int)blockIdx.x * slice, hi = min(m, lo + slice);
  float vmax = -CUDART_INF_F;
  constexpr int kGather = 4;
  float4 bx[kGather];
  int cl[kGather];
  for (int base = lo; base < hi; base += kPrepThreads * kGather) {
    float s[kGather];
    int rank[kGather];
#pragma unroll
    for (int u = 0; u < kGather; ++u) {
      const int i = base + u * kPrepThreads + tid;
      if (i < hi) {
        rank[u] = (int)(0xffffffffu - (uint32_t)(sortbuf[i] & 0xffffffffull));
        const int src = srcmap[rank[u]];
        bx[u] = reinterpret_cast<const float4*>(boxes)[(size_t)b * n + src];
        s[u] = sc[src];
        cl[u] = (int)classes[(size_t)b * n + src];
      }
    }
#pragma unroll
    for (int u = 0; u < kGather; ++u) {
      const int i = base + u * kPrepThreads + tid;
      if (i < hi) {
        set.score[o0 + i] = s[u];
        set.cls[o0 + i] = cl[u];
        set.src[o0 + i] = rank[u];
        reinterpret_cast<float4*>(set.box)[o0 + i] = bx[u];
        if (mode == kModeVanilla) reinterpret_cast<float4*>(set.nms_box)[o0 + i] = bx[u];
        vmax = fmaxf(vmax, fmaxf(fmaxf(bx[u].x, bx[u].y), fmaxf(bx[u].z, bx[u].w)));
      }
    }
  }
  B200DET_STAMP_NOSYNC(35);
  if (mode == kModeTrick) {                                      // m <= 1000: the thread's one box is bx[0]
    static_assert(kTrickMaxNumel / 4 <= kPrepThreads, "a coordinate-trick image has at most one candidate per thread");
    vmax = block_max(vmax, s_fmax);
    if (tid < m) {
      const float off = __fmul_rn((float)cl[0], __fadd_rn(vmax, 1.0f));
      reinterpret_cast<float4*>(set.nms_box)[o0 + tid] =
          make_float4(__fadd_rn(bx[0].x, off), __fadd_rn(bx[0].y, off), __fadd_rn(bx[0].z, off), __fadd_rn(bx[0].w, off));
    }
  }
  if (tid == 0 && blockIdx.x == 0) { set.mode[b] = mode; set.count[b] = m; }
  B200DET_STAMP(36);
}

// ------------------------------------------------------------------------------------------
// suppression mask
// ------------------------------------------------------------------------------------------
// One CTA of 64 threads per 64x64 tile: thread t owns row rb*64+t and walks the 64 staged column
// boxes.  Pass 1 is branch-free and fully unrolled (one broadcast LDS.128, 4 min/max, 2 compares
// per pair): w > 0 <=> min(x2) > max(x1) exactly in IEEE arithmetic, so only overlapping pairs
// become candidates.  Pass 2 evaluates the reference's exact IoU expression for the candidates.
// ZERO_SUP (thr < 0, where a zero IoU suppresses) makes every pair a candidate.
template <bool ZERO_SUP>
__global__ void __launch_bounds__(kNmsTile)
nms_mask_kernel(const CandSet set, const int wblocks, const int cap_pad, const float thr_up,
                unsigned long long* __restrict__ maskT) {
  // grid = (CTAs per image, batch); a CTA strides over the image's tiles, so an image that needs no dense mask
  // (finished by nms_class_kernel, or few candidates) costs a handful of CTAs instead of wblocks^2
  const int b = blockIdx.y;
  if (set.mode[b] == kModeDone) return;
  const int n = set.count[b];
  const int W = (n + kNmsTile - 1) / kNmsTile;                 // row / column blocks that hold candidates
  const bool same_class_only = set.mode[b] == kModeVanilla;
  __shared__ float4 cbox[kNmsTile];
  __shared__ float carea[kNmsTile];
  __shared__ int ccls[kNmsTile];
  const int t = threadIdx.x;
  const size_t o0 = (size_t)b * set.cap;
  for (int tile = blockIdx.x; tile < W * W; tile += gridDim.x) {
    const int rb = tile / W, cb = tile - rb * W;
    if (cb < rb) continue;
    __syncthreads();                                           // the previous tile's staging is no longer read
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
    const bool row_ok = i < n;
    const float4 a = row_ok ? reinterpret_cast<const float4*>(set.nms_box)[o0 + i]
                            : make_float4(CUDART_INF_F, CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    const float aarea = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const int acls = row_ok ? set.cls[o0 + i] : -2;
    unsigned long long bits = mask_row_bits<ZERO_SUP>(a, aarea, acls, cbox, carea, ccls, thr_up, same_class_only);
    if (cb == rb) bits &= ~((2ull << t) - 1ull);     // diagonal tile: only later boxes (j > t)
    if (!row_ok) bits = 0ull;
    maskT[((size_t)b * wblocks + cb) * cap_pad + i] = bits;   // 64 consecutive words per tile
  }
}

// ------------------------------------------------------------------------------------------
// greedy pass + output
// ------------------------------------------------------------------------------------------
// ---- small-n variant: the image's whole upper-triangular mask resident in shared memory ------
// cap <= kSmemScanMaxCap (every reference configuration: max_detection_box = 1000).  Column block w
// holds the words of rows [0, (w+1)*64); thread 0 issues one bulk copy per column block, all of
// them completing on one mbarrier.  All warps then flag, per (row block, column block), whether
// any word is non-zero; warp 0 runs the greedy pass touching only flagged pairs.
constexpr int kSmemScanThreads = 256;
constexpr int kSmemScanMaxBlocks = 28;                       // 28*29/2 * 512 B = 203 KB
constexpr int kSmemScanMaxCap = kSmemScanMaxBlocks * kNmsTile;   // 1792


__global__ void __launch_bounds__(kSmemScanThreads, 1)
nms_scan_smem_kernel(const CandSet set, const int wblocks, const int cap_pad,
                     const unsigned long long* __restrict__ maskT, const int clip_h, const int clip_w,
                     const NmsOut out) {
  extern __shared__ __align__(16) unsigned long long sm[];   // packed columns
  __shared__ unsigned long long keepw[32];
  __shared__ unsigned nz[32];
  __shared__ int s_pre[33];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kSmemScanThreads / 32;
  const int b = blockIdx.x;
  B200DET_STAMP_NOSYNC(16);
  if (set.mode[b] == kModeDone) return;
  const int n = set.count[b];
  const int W = (n + kNmsTile - 1) / kNmsTile;               // <= kSmemScanMaxBlocks
  const size_t o0 = (size_t)b * set.cap;
  const size_t q0 = (size_t)b * out.stride;

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
    if (W > 0) {
      mbar_arrive_expect_tx(&bar, (uint32_t)(col_off(W) * 8));
      for (int w = 0; w < W; ++w)
        bulk_g2s(sm + col_off(w), maskT + ((size_t)b * wblocks + w) * cap_pad, (uint32_t)((w + 1) * kNmsTile * 8), &bar);
    }
  }
  if (tid < 32) { keepw[tid] = 0ull; nz[tid] = 0u; }
  __syncthreads();
  if (W > 0) mbar_wait(&bar, 0);
  B200DET_STAMP_NOSYNC(17);

  // non-zero flags: bit w of nz[rb] <=> some row of block rb has a bit in column block w (w >= rb)
  for (int pair = warp; pair < W * W; pair += kWarps) {
    const int rb = pair / W, w = pair - rb * W;
    if (w < rb) continue;
    const unsigned long long* col = sm + col_off(w) + rb * kNmsTile;
    const bool any = __any_sync(0xffffffffu, (col[lane] | col[lane + 32]) != 0ull);
    if (any && lane == 0) atomicOr(&nz[rb], 1u << w);
  }
  __syncthreads();

  if (warp == 0) {
    unsigned long long myrem = 0ull;                         // lane w: removed-bitmap word w
    for (int rb = 0; rb < W; ++rb) {
      const int rows = min(kNmsTile, n - rb * kNmsTile);
      const unsigned long long valid = rows == kNmsTile ? ~0ull : ((1ull << rows) - 1ull);
      const unsigned long long cur = shfl64(myrem, rb);
      const unsigned flags = nz[rb];
      unsigned long long keep;
      if (!((flags >> rb) & 1u)) {
        keep = ~cur & valid;                                 // empty diagonal tile
      } else {
        const unsigned long long* diag = sm + col_off(rb) + rb * kNmsTile;
        keep = resolve_block(cur, valid, diag[lane], diag[lane + 32], lane);
      }
      if (lane == 0) keepw[rb] = keep;
      const bool k0 = (keep >> lane) & 1ull, k1 = (keep >> (lane + 32)) & 1ull;
      // flagged column blocks ahead, four at a time: one at a time the loop was a chain of dependent latencies (find
      // the bit, address, two loads, two warp reductions: ~0.12 us per column block, 0.9 of the 1.4 us a crowded
      // row block cost); four independent chains overlap
      for (unsigned m = flags & ~((2u << rb) - 1u); m;) {
        int w[4];
        unsigned long long v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          w[q] = m ? __ffs((int)m) - 1 : -1;
          m &= m - 1u;                                         // (0 stays 0)
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          v[q] = 0ull;
          if (w[q] >= 0) {
            const unsigned long long* col = sm + col_off(w[q]) + rb * kNmsTile;
            v[q] = (k0 ? col[lane] : 0ull) | (k1 ? col[lane + 32] : 0ull);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const unsigned long long r = warp_or64(v[q]);
          if (lane == w[q]) myrem |= r;
        }
      }
    }
    // exclusive prefix of kept counts per block
    __syncwarp();
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
  B200DET_STAMP_NOSYNC(18);

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
  for (int u = 0; u < kRowsPerThread; ++u)
    if (oidx[u] >= 0)
      store_kept(set, out, o0, q0, tid + u * kSmemScanThreads, oidx[u], clip_h, clip_w, bx[u], sc[u], cl[u], sr[u]);
  B200DET_STAMP(19);
  if (tid == 0) out.count[b] = s_pre[32];
}

// ---- large-n variant: a ring of row-block chunks ------------------------------------------------
// Chunk rb = the 64 rows of block rb for column blocks [rb, W): (W - rb) pieces of 512 bytes, one
// bulk copy each, landing on the slot's mbarrier.  Warp 0 resolves the diagonal, all warps OR the
// kept rows into the removed-bitmap (one column block per warp at a time), 64 threads write the
// kept rows, then thread 0 refills the slot with the chunk kRingSlots blocks ahead.
constexpr int kRingThreads = 256;
constexpr int kRingSlots = 3;

__global__ void __launch_bounds__(kRingThreads, 1)
nms_scan_ring_kernel(const CandSet set, const int wblocks, const int cap_pad,
                     const unsigned long long* __restrict__ maskT, const int clip_h, const int clip_w,
                     const NmsOut out) {
  extern __shared__ __align__(16) unsigned long long sm[];   // [kRingSlots][wblocks*64] chunks, [wblocks] removed
  __shared__ unsigned long long s_keep;
  __shared__ __align__(8) uint64_t bars[kRingSlots];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kRingThreads / 32;
  const int b = blockIdx.x;
  if (set.mode[b] == kModeDone) return;
  const int n = set.count[b];
  const int W = (n + kNmsTile - 1) / kNmsTile;
  const int slot_words = wblocks * kNmsTile;
  unsigned long long* removed = sm + (size_t)kRingSlots * slot_words;
  const size_t o0 = (size_t)b * set.cap;
  const size_t q0 = (size_t)b * out.stride;

  auto issue = [&](int rb) {                                 // thread 0 only
    const int slot = rb % kRingSlots;
    mbar_arrive_expect_tx(&bars[slot], (uint32_t)((W - rb) * kNmsTile * 8));
    for (int w = rb; w < W; ++w)
      bulk_g2s(sm + (size_t)slot * slot_words + (w - rb) * kNmsTile,
               maskT + ((size_t)b * wblocks + w) * cap_pad + rb * kNmsTile, kNmsTile * 8, &bars[slot]);
  };
  if (tid == 0) {
    for (int s = 0; s < kRingSlots; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
    for (int rb = 0; rb < min(kRingSlots, W); ++rb) issue(rb);
  }
  for (int w = tid; w < wblocks; w += kRingThreads) removed[w] = 0ull;
  __syncthreads();

  int total = 0;
  for (int rb = 0; rb < W; ++rb) {
    const int slot = rb % kRingSlots;
    mbar_wait(&bars[slot], (uint32_t)((rb / kRingSlots) & 1));
    const unsigned long long* chunk = sm + (size_t)slot * slot_words;
    const int rows = min(kNmsTile, n - rb * kNmsTile);
    const unsigned long long valid = rows == kNmsTile ? ~0ull : ((1ull << rows) - 1ull);
    if (warp == 0) {
      const unsigned long long keep = resolve_block(removed[rb], valid, chunk[lane], chunk[lane + 32], lane);
      if (lane == 0) s_keep = keep;
    }
    __syncthreads();
    const unsigned long long keep = s_keep;
    const bool k0 = (keep >> lane) & 1ull, k1 = (keep >> (lane + 32)) & 1ull;
    for (int w = rb + 1 + warp; w < W; w += kWarps) {        // one column block per warp: no atomics
      const unsigned long long* col = chunk + (w - rb) * kNmsTile;
      const unsigned long long v = warp_or64((k0 ? col[lane] : 0ull) | (k1 ? col[lane + 32] : 0ull));
      if (lane == 0 && v) removed[w] |= v;
    }
    if (tid < kNmsTile && ((keep >> tid) & 1ull)) {
      const int q = rb * kNmsTile + tid;
      const int o = total + __popcll(keep & ((1ull << tid) - 1ull));
      store_kept(set, out, o0, q0, q, o, clip_h, clip_w, reinterpret_cast<const float4*>(set.box)[o0 + q],
                 set.score[o0 + q], set.cls[o0 + q], set.src[o0 + q]);
    }
    total += __popcll(keep);
    __syncthreads();                                         // removed[] complete; slot free
    if (tid == 0 && rb + kRingSlots < W) {
      fence_proxy_async_smem();                              // generic reads of the slot before the async refill
      issue(rb + kRingSlots);
    }
  }
  if (tid == 0) out.count[b] = total;
}

struct NmsWorkspace {
  CandSet set;
  unsigned long long* mask;
  size_t bytes;
};

inline int nms_blocks(int cap) { return (cap + kNmsTile - 1) / kNmsTile; }

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
  // maskT[batch][blocks][blocks * 64] words
  w.mask = reinterpret_cast<unsigned long long*>(take((size_t)batch * nms_blocks(cap) * nms_blocks(cap) * kNmsTile * 8));
  w.bytes = off;
  return w;
}

template <typename K>
int set_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
  }
  return B200DET_OK;
}

}  // namespace

size_t nms_set_workspace_bytes(int batch, int cap) { return carve_nms(nullptr, batch, cap).bytes; }

void nms_set_carve(void* base, int batch, int cap, CandSet* set, unsigned long long** mask) {
  NmsWorkspace w = carve_nms(base, batch, cap);
  *set = w.set;
  *mask = w.mask;
}

void nms_threshold_params(double nms_thr, float* thr_up, bool* zero_sup) {
  float f = (float)nms_thr;
  if (!((double)f > nms_thr)) f = nextafterf(f, INFINITY);
  *thr_up = f;
  *zero_sup = !(nms_thr >= 0.0);
}

int launch_nms(const CandSet& set, int batch, double nms_thr, int clip_h, int clip_w,
               unsigned long long* mask, const NmsOut& out, cudaStream_t stream) {
  float thr_up;
  bool zero_suppresses;
  nms_threshold_params(nms_thr, &thr_up, &zero_suppresses);
  const int wblocks = nms_blocks(set.cap);
  const int cap_pad = wblocks * kNmsTile;
  int rc;
  // Above 1000 candidates an image is in torchvision's per-class mode: nms_class_kernel finishes those images
  // (unless a class is larger than it handles) and the dense kernels below only see what is left.
  static const bool no_class = getenv("B200DET_NO_CLASS_NMS") && getenv("B200DET_NO_CLASS_NMS")[0] == '1';
  bool class_ran = false;
  if (set.cap * 4 > kTrickMaxNumel && !no_class) {
    rc = launch_nms_class(set, batch, thr_up, zero_suppresses, clip_h, clip_w, out, stream);
    if (rc) return rc;
    class_ran = true;
  }
  // CTAs per image: every upper-triangle tile its own CTA up to a few waves of the machine, strided beyond.  After the
  // per-class kernel most images are already done and their CTAs exit at once: a small strided grid then (3 160 CTAs
  // per image that only return cost 15 us at 8 images), which an image with an oversized class walks a little longer.
  const int tri = wblocks * (wblocks + 1) / 2;
  const int per_image = class_ran ? (tri < 128 ? tri : 128) : (tri < 4096 ? tri : 4096);
  if (zero_suppresses)
    nms_mask_kernel<true><<<dim3(per_image, batch), kNmsTile, 0, stream>>>(set, wblocks, cap_pad, thr_up, mask);
  else
    nms_mask_kernel<false><<<dim3(per_image, batch), kNmsTile, 0, stream>>>(set, wblocks, cap_pad, thr_up, mask);
  rc = check_launch();
  if (rc) return rc;
  if (wblocks <= kSmemScanMaxBlocks) {
    const size_t smem = (size_t)kNmsTile * (wblocks * (wblocks + 1) / 2) * sizeof(unsigned long long);
    if ((rc = set_smem(nms_scan_smem_kernel, smem))) return rc;
    nms_scan_smem_kernel<<<batch, kSmemScanThreads, smem, stream>>>(set, wblocks, cap_pad, mask, clip_h, clip_w, out);
    return check_launch();
  }
  const size_t smem = ((size_t)kRingSlots * wblocks * kNmsTile + wblocks) * sizeof(unsigned long long);
  if ((rc = set_smem(nms_scan_ring_kernel, smem))) return rc;
  nms_scan_ring_kernel<<<batch, kRingThreads, smem, stream>>>(set, wblocks, cap_pad, mask, clip_h, clip_w, out);
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
  const size_t smem = prep_hist_offset(n, n2) + (size_t)kHistBins * 4 + (size_t)n * 8;
  int rc = set_smem(nms_prepare_kernel, smem);
  if (rc) return rc;
  nms_prepare_kernel<<<dim3(kPrepParts, batch), kPrepThreads, smem, st>>>(n, boxes, scores, reinterpret_cast<const long long*>(classes),
                                                       in_count, score_thr, set);
  rc = check_launch();
  if (rc) return rc;
  NmsOut out{out_score, reinterpret_cast<long long*>(out_cls), out_box, reinterpret_cast<long long*>(out_keep),
             out_count, n};
  // Up to 1024 candidates the NMS half of the fused head kernel takes the set (class buckets in shared memory, one CTA
  // per image: 1 000 crowded candidates x 16 images 47 -> ~20 us against the dense mask + scan kernels, same keep set).
  // B200DET_DENSE_NMS=1 (read at every call) forces the dense chain: the tests compare the two.
  const char* dense = getenv("B200DET_DENSE_NMS");
  if (fused_nms_supported(n, nms_thr) && !(dense && dense[0] == '1'))
    return launch_fused_nms_from_set(set, batch, nms_thr, clip_h, clip_w, out, st);
  return launch_nms(set, batch, nms_thr, clip_h, clip_w, mask, out, st);
}
