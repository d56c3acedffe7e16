// K3, per-class path — torchvision's _batched_nms_vanilla (what batched_nms does on CPU above 1000 candidates:
// an independent NMS per class on the raw boxes) without the dense n x n mask.
//
// The dense path evaluates every pair of the image (12.5 M for 5000 candidates) although only same-class pairs
// can suppress, and then walks all 79 row blocks one after the other.  Here one CTA per image
//   1. sorts (class, score rank) keys in shared memory, which cuts the score-ordered candidate list into one
//      contiguous, score-ordered segment per class;
//      (a cluster of 8 CTAs per image; each sorts its own copy and takes an eighth of the segments);
//   2. resolves every class on its own: a class of <= 64 boxes is one tile for one warp; a larger class is taken
//      by a whole CTA (tiles evaluated by all warps into shared memory, then one warp's greedy pass with the
//      removed-bitmap in registers) — the same arithmetic (mask_row_bits: fp32 IoU in torchvision's operation order, threshold
//      compared as a double) on ~n^2 / (2 * classes) pairs, and the classes advance in parallel;
//   3. writes the kept candidates in score order (prefix popcount over a keep bitmap indexed by rank).
// Images it handles are marked kModeDone and skipped by the dense kernels, which stay in the chain for
// coordinate-trick images (<= 1000 candidates) and for classes larger than kClassMaxSeg boxes.
#include <cooperative_groups.h>

#include "block_utils.cuh"
#include "nms_body.cuh"

B200DET_TRACE_BUFFER(nmsclass)

namespace cg = cooperative_groups;

namespace b200det {

namespace {

constexpr int kClassCluster = 8;               // CTAs (SMs) per image; the class segments are dealt to all their warps
constexpr int kClassThreads = 1024;
constexpr int kClassWarps = kClassThreads / 32;
constexpr int kClassMaxSeg = 1024;            // boxes of one class handled by a warp (16 blocks of 64)
constexpr int kRankBits = 13;                 // rank < 8192 = B200DET_MAX_BOX
constexpr int kEarly = 8;                     // greedy passes that may follow their class's pair tests row by row
constexpr int kBoxBudget = 2048;              // boxes of one batch of classes in shared memory
constexpr int kOrderMax = 256;                // classes dealt to the CTAs in order of size up to this many
constexpr int kDenseClasses = 256;            // class ids below this take the table-driven stable split
constexpr int kTileBudget = (kClassMaxSeg / 64) * (kClassMaxSeg / 64 + 1) / 2;   // 136 tiles of 64 x 64 bits in shared memory

// ascending bitonic sort of n (power of two) 32-bit keys in shared memory
__device__ __forceinline__ void bitonic_sort_asc_u32(unsigned* buf, const int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const unsigned a = buf[i], c = buf[p];
        const bool up = (i & k) == 0;
        if ((a > c) == up) { buf[i] = c; buf[p] = a; }
      }
      __syncthreads();
    }
  }
}

template <bool ZERO_SUP>
__global__ void __cluster_dims__(kClassCluster, 1, 1) __launch_bounds__(kClassThreads, 1)
nms_class_kernel(const CandSet set, const float thr_up, const int clip_h, const int clip_w, const NmsOut out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_scan[33];
  __shared__ float s_fmax[32];
  __shared__ int s_pre[B200DET_MAX_BOX / 64 + 1];
  __shared__ unsigned short s_cstart[kDenseClasses];             // first slot of a class (dense-id path)
  __shared__ unsigned short s_toff[kTileBudget + 1];            // first tile of the batch's classes
  __shared__ unsigned short s_boff[kTileBudget + 1];            // first box of the batch's classes in sbox
  __shared__ unsigned short s_s0[kTileBudget];                  // where the class starts in keys[]
  __shared__ unsigned short s_roff[kTileBudget];                // first tile row of the batch's classes in s_done
  __shared__ int s_done[kTileBudget];                           // units of pair tests finished, per tile row of a class
  __shared__ int s_queue[3];                                    // next unit, next class to be resolved, early passes
  __shared__ int s_batch[4];                                    // classes, tiles, boxes, tile rows of the batch
  __shared__ unsigned short s_ord[kOrderMax];                   // classes by size, largest first
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const int n = set.count[b];
  // Every CTA of the cluster takes the same decisions from the same data (each sorts its own copy of the keys:
  // the sort is not worth distributing), so the early exits below are taken by all of them or by none.
  if (set.mode[b] != kModeVanilla || n <= 0) return;
  B200DET_STAMP(0);
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  // shared memory: keys [n2] u32 | segment starts [n2] u16 | keep bitmap [n2 / 32] u32 | boxes and areas of the
  // classes being resolved | their tile bits
  unsigned* keys = reinterpret_cast<unsigned*>(smem_raw);
  unsigned short* seg = reinterpret_cast<unsigned short*>(keys + n2);
  unsigned* keepbits = reinterpret_cast<unsigned*>(seg + n2);
  const int kwords = (n2 + 31) / 32;
  float4* cbox_all = reinterpret_cast<float4*>(smem_raw + (((size_t)n2 * 6 + (size_t)kwords * 4 + 15) / 16) * 16);
  float* carea_all = reinterpret_cast<float*>(cbox_all + kBoxBudget + kNmsTile);
  unsigned long long* tmask = reinterpret_cast<unsigned long long*>(carea_all + kBoxBudget + kNmsTile);   // [136 tiles][64]
  const size_t o0 = (size_t)b * set.cap;
  const size_t q0 = (size_t)b * out.stride;

  // ---- 1. (class, rank) keys, sorted: one contiguous score-ordered segment per class -----------------------
  // The candidate list arrives in score order, so this is a STABLE split by class.
  // (a) Class ids below kDenseClasses (COCO, VOC): warp w takes a contiguous run of 32-key groups; in a group the
  //     lanes of one class find each other with ballots on the id's bits, and a per-warp count table [warp][class] gives every key
  //     its position among the warp's keys of its class.  One pass down the table's columns turns the counts into
  //     offsets, one scan over the classes gives the class starts (= the segments) — ~2 us for 5 000 candidates, and
  //     nothing is ranked or compared.
  // (b) Otherwise a counting sort on the class (ids below kHistBins; the histogram and the scratch copy of the keys
  //     live where the column staging and the tile bits will be) with the keys ranked inside their class afterwards,
  //     and the 8 192-key bitonic network for class ids it cannot bin.
  static_assert(kClassThreads == kHistThreads, "the class histogram is scanned by 1024 threads");
  unsigned* chist = reinterpret_cast<unsigned*>(tmask);                 // [kHistBins] (tmask holds 136 x 64 x 8 bytes)
  unsigned* tmpk = reinterpret_cast<unsigned*>(cbox_all);               // [n] (cbox_all holds 2112 x 16 bytes)
  unsigned short* wcnt = reinterpret_cast<unsigned short*>(tmask);      // [kClassWarps][kDenseClasses]
  static_assert((size_t)(kClassMaxSeg / kNmsTile) * (kClassMaxSeg / kNmsTile + 1) / 2 * kNmsTile * 8 >= (size_t)kHistBins * 4,
                "class histogram aliases the tile bits");
  static_assert((size_t)(kBoxBudget + kNmsTile) * sizeof(float4) >= (size_t)B200DET_MAX_BOX * 4, "key scratch aliases the staged boxes");
  static_assert(kBoxBudget >= kClassMaxSeg && (kBoxBudget + kNmsTile) % 2 == 0, "a class alone fits; the tile bits stay 8-byte aligned");
  static_assert(kClassWarps * kDenseClasses * 2 == kClassThreads * 16, "one uint4 per thread clears the count table");
  reinterpret_cast<uint4*>(wcnt)[tid] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < kwords; i += kClassThreads) keepbits[i] = 0u;
  __syncthreads();
  B200DET_STAMP_NOSYNC(21);
  constexpr int kGroupsPerWarp = B200DET_MAX_BOX / 32 / kClassWarps;   // 8
  const int gpw = ((n + 31) / 32 + kClassWarps - 1) / kClassWarps;     // groups per warp in this image
  int cls_q[kGroupsPerWarp];
  unsigned short loc_q[kGroupsPerWarp];
  bool cmax = false, sparse = false;
#pragma unroll
  for (int q = 0; q < kGroupsPerWarp; ++q) {
    const int i = (warp * gpw + q) * 32 + lane;
    cls_q[q] = (q < gpw && i < n) ? set.cls[o0 + i] : 0;
  }
  unsigned short* wrow = wcnt + warp * kDenseClasses;
#ifdef B200DET_TRACE
  if (cls_q[0] == 0x7fffffff) g_trace[50] = 1;                  // (waits for the first load)
  B200DET_STAMP_NOSYNC(22);
#endif
#pragma unroll
  for (int q = 0; q < kGroupsPerWarp; ++q) {
    if (q >= gpw) { cls_q[q] = -1; continue; }                   // (uniform over the CTA)
    const int i = (warp * gpw + q) * 32 + lane;
    const int c = cls_q[q];
    const bool in = i < n;
    const bool ok = in && c >= 0 && c < kDenseClasses;
    sparse = sparse || (in && !ok);
    cmax = cmax || (in && (c < 0 || c >= (1 << (32 - kRankBits - 1))));
    // lanes of the same class: 8 ballots on the bits of the id (match.any takes one pass per distinct value: the five
    // groups of a warp cost ~4 us of this kernel)
    unsigned peers = __ballot_sync(0xffffffffu, ok);
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
      const unsigned v = __ballot_sync(0xffffffffu, (c >> bit) & 1);
      peers &= ((c >> bit) & 1) ? v : ~v;
    }
    static_assert(kDenseClasses == 256, "eight bits tell the dense class ids apart");
    const int r = __popc(peers & ((1u << lane) - 1u));
    const unsigned short base = ok ? wrow[c] : (unsigned short)0;
    __syncwarp();
    if (ok && r == 0) wrow[c] = (unsigned short)(base + __popc(peers));
    __syncwarp();
    cls_q[q] = ok ? c : -1;
    loc_q[q] = (unsigned short)(base + r);
  }
  B200DET_STAMP_NOSYNC(23);
  const bool bad_class = __syncthreads_or(cmax) != 0;            // also the barrier after the counts
  const bool dense = __syncthreads_or(sparse) == 0;
  B200DET_STAMP(1);
  int n_seg = 0;
  float longest = 0.f;
  if (dense) {
    int tot = 0;
    if (tid < kDenseClasses) {
#pragma unroll 8
      for (int w = 0; w < kClassWarps; ++w) {                      // counts -> offset of the warp inside the class
        const int v = wcnt[w * kDenseClasses + tid];
        wcnt[w * kDenseClasses + tid] = (unsigned short)tot;
        tot += v;
      }
    }
    int total;                                                   // low half: keys, high half: classes present
    const int ex = block_exclusive_scan(tot | (tot > 0 ? 1 << 16 : 0), s_scan, &total);
    if (tid < kDenseClasses) {
      s_cstart[tid] = (unsigned short)(ex & 0xffff);
      if (tot > 0) seg[ex >> 16] = (unsigned short)(ex & 0xffff);
    }
    n_seg = total >> 16;
    longest = block_max((float)tot, s_fmax);                     // also the barrier before the placement
#pragma unroll
    for (int q = 0; q < kGroupsPerWarp; ++q) {
      const int c = cls_q[q];
      if (c < 0) continue;
      const unsigned i = (unsigned)((warp * gpw + q) * 32 + lane);
      keys[(int)s_cstart[c] + (int)wrow[c] + (int)loc_q[q]] = ((unsigned)c << kRankBits) | i;
    }
    __syncthreads();
    B200DET_STAMP(2);
  } else {
    {
      uint4* h4 = reinterpret_cast<uint4*>(chist);
#pragma unroll
      for (int q = 0; q < kHistPerThread / 4; ++q) h4[tid + q * kClassThreads] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    float cbig = 0.f;
    for (int i = tid; i < n2; i += kClassThreads) {
      unsigned key = 0xffffffffu;
      if (i < n) {
        const int c = set.cls[o0 + i];
        cbig = fmaxf(cbig, (c < 0 || c >= kHistBins) ? 1.f : 0.f);
        key = ((unsigned)c << kRankBits) | (unsigned)i;
        if (c >= 0 && c < kHistBins) atomicAdd(&chist[hist_slot(c)], 1u);
      }
      keys[i] = key;
    }
    const bool unbinned = block_max(cbig, s_fmax) > 0.f;         // also the barrier before the sort
    if (!unbinned) {
      unsigned hb[kHistPerThread];                                // thread t: classes [16 t, 16 t + 16), ascending
      int mine = 0;
#pragma unroll
      for (int q = 0; q < kHistPerThread; ++q) {
        hb[q] = chist[q * kClassThreads + tid];
        mine += (int)hb[q];
      }
      int total;
      unsigned start = (unsigned)block_exclusive_scan(mine, s_scan, &total);
#pragma unroll
      for (int q = 0; q < kHistPerThread; ++q) {                 // counts -> first slot of the class
        const unsigned c = hb[q];
        chist[q * kClassThreads + tid] = start;
        start += c;
      }
      __syncthreads();
      for (int i = tid; i < n; i += kClassThreads) {
        const unsigned key = keys[i];
        tmpk[atomicAdd(&chist[hist_slot((int)(key >> kRankBits))], 1u)] = key;
      }
      __syncthreads();
      for (int i = tid; i < n; i += kClassThreads) {
        const unsigned key = tmpk[i];
        const int c = (int)(key >> kRankBits);
        const unsigned first = c == 0 ? 0u : chist[hist_slot(c - 1)];   // = end of the classes below
        const unsigned end = chist[hist_slot(c)];
        unsigned rank = 0;
        for (unsigned j = first; j < end; ++j) rank += tmpk[j] < key ? 1u : 0u;
        keys[first + rank] = key;
      }
      __syncthreads();
    } else {
      bitonic_sort_asc_u32(keys, n2);
    }
    B200DET_STAMP(2);
    // segment starts (order-preserving compaction) and the longest segment
    for (int base = 0; base < n; base += kClassThreads) {
      const int i = base + tid;
      const bool start = i < n && (i == 0 || (keys[i] >> kRankBits) != (keys[i - 1] >> kRankBits));
      int total;
      const int pos = n_seg + block_exclusive_scan(start ? 1 : 0, s_scan, &total);
      if (start) seg[pos] = (unsigned short)i;
      n_seg += total;
    }
    __syncthreads();
    for (int s = tid; s < n_seg; s += kClassThreads)
      longest = fmaxf(longest, (float)((s + 1 < n_seg ? (int)seg[s + 1] : n) - (int)seg[s]));
    longest = block_max(longest, s_fmax);
  }
  if (bad_class || longest > (float)kClassMaxSeg) return;        // left to the dense path (mode stays vanilla)
  cluster.sync();                                                // CTA 0's keep bitmap is clear; every CTA runs
  // kept ranks go to the bitmap of EVERY CTA of the cluster (32-bit reductions through DSMEM), so that after the
  // cluster barrier each of them can write a slice of the output

  B200DET_STAMP(3);
  // ---- 2. every class on its own ---------------------------------------------------------------------------------
  auto rank_of = [&](const int pos) { return keys[pos] & ((1u << kRankBits) - 1u); };
  auto keep_rank = [&](const unsigned r) {                      // bit r of the keep bitmap of every CTA of the cluster
    const unsigned word = (unsigned)__cvta_generic_to_shared(keepbits + (r >> 5));
#pragma unroll
    for (int t = 0; t < kClassCluster; ++t) {
      unsigned remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(word), "r"(t));
      asm volatile("red.relaxed.cluster.shared::cluster.or.b32 [%0], %1;" ::"r"(remote), "r"(1u << (r & 31)) : "memory");
    }
  };

  // Dealing: the classes in order of size, largest first, laid out boustrophedon over the CTAs (0..7, 7..0, ...), so
  // every CTA gets a like share of the pair tests (plain round-robin left the slowest CTA of a cluster ~13 us behind
  // the first on 5 000 candidates of 80 classes).  The order costs one n_seg-long count per class, so it is taken up
  // to kOrderMax classes; beyond, round-robin.
  const bool ordered = n_seg <= kOrderMax;
  if (ordered && tid < n_seg) {
    const int mine = (tid + 1 < n_seg ? (int)seg[tid + 1] : n) - (int)seg[tid];
    int before = 0;                                              // classes ahead of this one: larger, or equal and earlier
    for (int o = 0; o < n_seg; ++o) {
      const int other = (o + 1 < n_seg ? (int)seg[o + 1] : n) - (int)seg[o];
      before += (other > mine || (other == mine && o < tid)) ? 1 : 0;
    }
    s_ord[before] = (unsigned short)tid;
  }
  __syncthreads();
  B200DET_STAMP_NOSYNC(24);
  auto my_segment = [&](const int j) {                          // the j-th class of this CTA
    return ordered ? (int)s_ord[j * kClassCluster + ((j & 1) ? kClassCluster - 1 - rank : rank)] : rank + j * kClassCluster;
  };
  int my_n;
  {
    const int rows = n_seg / kClassCluster, left = n_seg - rows * kClassCluster;
    const int col = (ordered && (rows & 1)) ? kClassCluster - 1 - rank : rank;
    my_n = rows + (col < left ? 1 : 0);
  }

  // A CTA takes its classes in batches whose boxes and tiles (64 x 64 blocks of the upper triangle, one for a class
  // of <= 64 boxes) fit shared memory — for 5 000 candidates of 80 classes all ten classes of a CTA are one batch.
  // Per batch the boxes are gathered into shared memory once, class after class; then the warps serve two queues:
  //   * pair tests, a (tile, 16-column part) unit at a time — a single warp needs ~10 us of dependent issue for a
  //     whole tile, so they are spread as thin as they go; the units of the largest class come first;
  //   * greedy passes, one warp per class over the stored bits (a serial chain of up to ~8 us), taken as soon as the
  //     last unit of the class has been counted in, so the long chains run beside the pair tests of the other classes.
  constexpr int kPart = 16, kParts = kNmsTile / kPart;
  float4* sbox = cbox_all;                                       // [kBoxBudget + 64] the batch's boxes, class after class
  float* sarea = carea_all;                                      // [kBoxBudget + 64]
  for (int j0 = 0; j0 < my_n;) {
    // the batch: classes j0 .. j0 + cnt - 1, as many as fit the tile and the box budget.  Warp 0 builds the tables, a
    // class per lane with warp scans (every thread walking the classes one by one kept all 32 warps busy for 1.6 us).
    if (warp == 0) {
      int cnt = 0, tiles = 0, boxes = 0, rows = 0;
      for (bool more = true; more;) {
        const int j = j0 + cnt + lane;
        const bool have = j < my_n && cnt + lane < kTileBudget;
        int s0 = 0, size = 0;
        if (have) {
          const int sg = my_segment(j);
          s0 = seg[sg];
          size = (sg + 1 < n_seg ? (int)seg[sg + 1] : n) - s0;
        }
        const int W = (size + kNmsTile - 1) / kNmsTile;          // 1 .. 16 blocks (0: no class)
        const int t = W * (W + 1) / 2;
        int it = t, ib = size, ir = W;                           // inclusive scans over the lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int ot = __shfl_up_sync(0xffffffffu, it, d), ob = __shfl_up_sync(0xffffffffu, ib, d);
          const int orw = __shfl_up_sync(0xffffffffu, ir, d);
          if (lane >= d) { it += ot; ib += ob; ir += orw; }
        }
        const bool fits = have && tiles + it <= kTileBudget && boxes + ib <= kBoxBudget;   // (a class alone always fits)
        const unsigned ok = __ballot_sync(0xffffffffu, fits);
        const int take = __ffs((int)~ok) - 1 < 0 ? 32 : __ffs((int)~ok) - 1;               // leading lanes that fit
        if (lane < take) {
          s_toff[cnt + lane] = (unsigned short)(tiles + it - t);
          s_boff[cnt + lane] = (unsigned short)(boxes + ib - size);
          s_s0[cnt + lane] = (unsigned short)s0;
          s_roff[cnt + lane] = (unsigned short)(rows + ir - W);
        }
        const int last = max(take - 1, 0);
        const int at = __shfl_sync(0xffffffffu, it, last), ab = __shfl_sync(0xffffffffu, ib, last);
        const int ar = __shfl_sync(0xffffffffu, ir, last);
        if (take > 0) { tiles += at; boxes += ab; rows += ar; }
        cnt += take;
        more = take == 32 && j0 + cnt < my_n && cnt < kTileBudget;
      }
      if (lane == 0) {
        s_toff[cnt] = (unsigned short)tiles;
        s_boff[cnt] = (unsigned short)boxes;
        s_batch[0] = cnt;
        s_batch[1] = tiles;
        s_batch[2] = boxes;
        s_batch[3] = rows;
        s_queue[0] = 0;                                          // next unit
        s_queue[1] = 0;                                          // next class to be resolved
        s_queue[2] = 0;                                          // greedy passes running ahead of their class's tiles
      }
    }
    __syncthreads();
    const int cnt = s_batch[0], tiles = s_batch[1], boxes = s_batch[2], rows = s_batch[3];
    for (int i = tid; i < rows; i += kClassThreads) s_done[i] = 0;
    for (int i = tid; i < tiles * kNmsTile; i += kClassThreads) tmask[i] = 0ull;
    __syncthreads();
    if (j0 == 0) { B200DET_STAMP_NOSYNC(25); }
    // class of the batch that holds box / tile number v of the batch (offsets in `off`)
    auto class_of = [&](const unsigned short* off, const int v) {
      int lo = 0, hi = cnt;                                      // off[lo] <= v < off[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int)off[mid] <= v) lo = mid; else hi = mid;
      }
      return lo;
    };
    for (int i = tid; i < boxes; i += kClassThreads) {
      const int j = class_of(s_boff, i);
      const float4 v = reinterpret_cast<const float4*>(set.box)[o0 + rank_of((int)s_s0[j] + i - (int)s_boff[j])];
      sbox[i] = v;
      sarea[i] = __fmul_rn(__fsub_rn(v.z, v.x), __fsub_rn(v.w, v.y));
    }
    __syncthreads();
    if (j0 == 0) { B200DET_STAMP_NOSYNC(26); }
    const int units = tiles * kParts;
    volatile int* queue = s_queue;
    volatile int* done = s_done;
    for (;;) {
      // (a) a class whose greedy pass can start and nobody has taken?  (Classes are handed out in order, the order of
      //     their units.)  A class of up to two row blocks starts when all its tiles are complete.  A larger one
      //     starts as soon as its FIRST tile row is and then follows the pair tests row by row — it is most of the
      //     CTA's critical path: 28 tiles, then a 7 us chain — waiting where it catches up; at most kEarly such
      //     passes at a time, so that the other warps keep the pair tests going.
      int g = -1;
      bool early = false;
      if (lane == 0) {
        const int next = queue[1];
        if (next < cnt) {
          const int size = (int)s_boff[next + 1] - (int)s_boff[next];
          const int W = (size + kNmsTile - 1) / kNmsTile, r0 = s_roff[next];
          bool ready = done[r0] == W * kParts;
          early = W > 2;
          if (!early && ready && W == 2) ready = done[r0 + 1] == kParts;
          if (early && ready) ready = atomicAdd(&s_queue[2], 1) < kEarly || (atomicSub(&s_queue[2], 1), false);
          if (ready) {
            if (atomicCAS(&s_queue[1], next, next + 1) == next) g = next;
            else if (early) atomicSub(&s_queue[2], 1);
          }
        }
      }
      g = __shfl_sync(0xffffffffu, g, 0);
      if (g >= 0) {
        __threadfence_block();                                   // the bits counted in by done[g] are visible
        B200DET_STAMP_ANY(lane == 0 && g == 0 && j0 == 0 && blockIdx.x == 0 && blockIdx.y == 0, 15);
        B200DET_STAMP_ANY(lane == 0 && g == cnt - 1 && j0 == 0 && blockIdx.x == 0 && blockIdx.y == 0, 17);
        const int s0 = s_s0[g], size = (int)s_boff[g + 1] - (int)s_boff[g];
        const int W = (size + kNmsTile - 1) / kNmsTile, row0 = s_roff[g];
        const unsigned long long* buf = tmask + (size_t)s_toff[g] * kNmsTile;
        unsigned long long myrem = 0ull;                         // lane w: removed bits of the class's block w
        int tile = 0;
        for (int rb = 0; rb < W; ++rb) {
          if (W > 2 && rb > 0) {                                 // the pair tests of this tile row
            while (done[row0 + rb] != (W - rb) * kParts) __nanosleep(64);
            __threadfence_block();
          }
          const unsigned long long* dg = buf + (size_t)tile * kNmsTile;
          const int rows_here = min(kNmsTile, size - rb * kNmsTile);
          const unsigned long long valid = rows_here == kNmsTile ? ~0ull : ((1ull << rows_here) - 1ull);
          const unsigned long long keep = resolve_block(shfl64(myrem, rb), valid, dg[lane], dg[lane + 32], lane);
          const bool k0 = (keep >> lane) & 1ull, k1 = (keep >> (lane + 32)) & 1ull;
          if (k0) keep_rank(rank_of(s0 + rb * kNmsTile + lane));
          if (k1) keep_rank(rank_of(s0 + rb * kNmsTile + lane + 32));
#pragma unroll 4
          for (int cb = rb + 1; cb < W; ++cb) {                 // (independent chains of loads and warp reductions)
            const unsigned long long* col = buf + (size_t)(tile + cb - rb) * kNmsTile;
            const unsigned long long v = warp_or64((k0 ? col[lane] : 0ull) | (k1 ? col[lane + 32] : 0ull));
            if (lane == cb) myrem |= v;
          }
          tile += W - rb;
        }
        if (W > 2 && lane == 0) atomicSub(&s_queue[2], 1);
        B200DET_STAMP_ANY(lane == 0 && g == 0 && j0 == 0 && blockIdx.x == 0 && blockIdx.y == 0, 16);
        B200DET_STAMP_ANY(lane == 0 && g == cnt - 1 && j0 == 0 && blockIdx.x == 0 && blockIdx.y == 0, 18);
        continue;
      }
      // (b) a unit of pair tests
      int unit = 0;
      if (lane == 0) unit = queue[0] < units ? atomicAdd(&s_queue[0], 1) : units;
      unit = __shfl_sync(0xffffffffu, unit, 0);
      B200DET_STAMP_ANY(lane == 0 && unit == units - 1 && j0 == 0 && blockIdx.x == 0 && blockIdx.y == 0, 19);
      B200DET_STAMP_ANY(lane == 0 && unit == 0 && j0 == 0 && blockIdx.x == 0 && blockIdx.y == 0, 20);
      if (unit < units) {
        const int tile_g = unit / kParts, part = unit - tile_g * kParts;
        const int j = class_of(s_toff, tile_g);
        const int b0 = s_boff[j], size = (int)s_boff[j + 1] - b0;
        const int W = (size + kNmsTile - 1) / kNmsTile;
        int rb = 0, rem = tile_g - (int)s_toff[j];
        while (rem >= W - rb) { rem -= W - rb; ++rb; }            // row-major upper triangle
        const int cb = rb + rem;
        const int col0 = cb * kNmsTile + part * kPart;           // (positions inside the class)
        if (col0 < size) {                                       // else: no column in this part
          const bool diag = cb == rb;
          const bool need1 = !(diag && part < 2);                // diagonal tile: rows 32-63 only see columns > 32
          const int r0 = rb * kNmsTile + lane, r1 = r0 + 32;
          // Rows and columns past the end of the class read the boxes that follow in shared memory (64 entries of
          // slack behind the last class): row bits are cleared below, column bits beyond the class never reach a
          // kept rank (the greedy pass masks every block with its valid rows).
          const float4 a0 = sbox[b0 + r0], a1 = sbox[b0 + r1];
          const float4* cbox = sbox + b0 + col0;
          const float* carea = sarea + b0 + col0;
          unsigned lo0, lo1 = 0u;
          if constexpr (ZERO_SUP) {
            lo0 = mask_row_bits_part<true, kPart>(a0, sarea[b0 + r0], cbox, carea, thr_up);
            if (need1) lo1 = mask_row_bits_part<true, kPart>(a1, sarea[b0 + r1], cbox, carea, thr_up);
          } else {
            if (need1) mask_rows2_part<kPart, true>(a0, a1, cbox, carea, thr_up, lo0, lo1);
            else mask_rows2_part<kPart, false>(a0, a1, cbox, carea, thr_up, lo0, lo1);
          }
          unsigned long long d0 = r0 < size ? (unsigned long long)lo0 << (part * kPart) : 0ull;
          unsigned long long d1 = r1 < size ? (unsigned long long)lo1 << (part * kPart) : 0ull;
          if (diag) {                                            // only later boxes (column > row)
            d0 &= ~((2ull << lane) - 1ull);
            d1 &= ~((2ull << (lane + 32)) - 1ull);
          }
          // the four parts of a word own 16 bits each: native 32-bit atomics on its halves
          unsigned* w0 = reinterpret_cast<unsigned*>(tmask + (size_t)tile_g * kNmsTile + lane) + (part >> 1);
          const unsigned h0 = (unsigned)(d0 >> ((part >> 1) * 32)), h1 = (unsigned)(d1 >> ((part >> 1) * 32));
          if (h0) atomicOr(w0, h0);
          if (h1) atomicOr(w0 + 64, h1);
        }
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();                                 // the unit's bits before its count
          atomicAdd(&s_done[(int)s_roff[j] + rb], 1);
        }
        continue;
      }
      // (c) nothing to test: leave once every class has been taken, else wait for the units in flight
      int left = 0;
      if (lane == 0) left = queue[1] < cnt ? 1 : 0;
      if (!__shfl_sync(0xffffffffu, left, 0)) break;
      __nanosleep(100);
    }
    __syncthreads();                                             // the buffers and the batch tables are free again
    if (j0 == 0) { B200DET_STAMP_NOSYNC(11); B200DET_NOTE_IF(blockIdx.x == 0 && blockIdx.y == 0, 12, tiles); B200DET_NOTE_IF(blockIdx.x == 0 && blockIdx.y == 0, 13, cnt); }
    j0 += cnt;
  }
  B200DET_STAMP_NOSYNC(14);
  cluster.sync();                                                // every kept rank has landed in every CTA
  B200DET_STAMP(4);

  // ---- 3. kept candidates in score (rank) order: every CTA writes an eighth of the ranks ------------------------------
  const int w64 = (n + 63) / 64;
  int run = 0;
  for (int base = 0; base < w64; base += kClassThreads) {        // exclusive prefix of kept counts per 64-rank word
    const int wi = base + tid;
    const int cnt = wi < w64 ? __popc(keepbits[2 * wi]) + (2 * wi + 1 < kwords ? __popc(keepbits[2 * wi + 1]) : 0) : 0;
    int total;
    const int excl = block_exclusive_scan(cnt, s_scan, &total);
    if (wi < w64) s_pre[wi] = run + excl;
    run += total;
  }
  __syncthreads();
  const int slice = (w64 + kClassCluster - 1) / kClassCluster * 64;
  const int q_end = min(n, (rank + 1) * slice);
  for (int q = rank * slice + tid; q < q_end; q += kClassThreads) {
    if (!((keepbits[q >> 5] >> (q & 31)) & 1u)) continue;
    const unsigned lo = keepbits[(q >> 6) * 2];
    const unsigned hi = (q >> 6) * 2 + 1 < kwords ? keepbits[(q >> 6) * 2 + 1] : 0u;
    const unsigned long long word = ((unsigned long long)hi << 32) | lo;
    const int o = s_pre[q >> 6] + __popcll(word & ((1ull << (q & 63)) - 1ull));
    store_kept(set, out, o0, q0, q, o, clip_h, clip_w, reinterpret_cast<const float4*>(set.box)[o0 + q],
               set.score[o0 + q], set.cls[o0 + q], set.src[o0 + q]);
  }
  B200DET_STAMP(5);
#ifdef B200DET_TRACE
  if (tid == 0 && blockIdx.x == 0) g_trace[8] = n_seg;
#endif
  if (tid == 0 && rank == 0) {
    out.count[b] = run;
    set.mode[b] = kModeDone;                                     // the dense kernels skip this image
  }
}

size_t class_smem_bytes(int cap) {
  int n2 = 1;
  while (n2 < cap) n2 <<= 1;
  const size_t head = (((size_t)n2 * 6 + (size_t)((n2 + 31) / 32) * 4 + 15) / 16) * 16;
  const size_t max_tiles = (size_t)(kClassMaxSeg / kNmsTile) * (kClassMaxSeg / kNmsTile + 1) / 2;
  return head + (size_t)(kBoxBudget + kNmsTile) * (sizeof(float4) + sizeof(float)) + max_tiles * kNmsTile * 8;
}

}  // namespace

int launch_nms_class(const CandSet& set, int batch, float thr_up, bool zero_sup, int clip_h, int clip_w,
                     const NmsOut& out, cudaStream_t stream) {
  const size_t smem = class_smem_bytes(set.cap);
  cudaError_t e = cudaFuncSetAttribute(zero_sup ? nms_class_kernel<true> : nms_class_kernel<false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
  const dim3 grid(kClassCluster, batch);
  if (zero_sup) nms_class_kernel<true><<<grid, kClassThreads, smem, stream>>>(set, thr_up, clip_h, clip_w, out);
  else nms_class_kernel<false><<<grid, kClassThreads, smem, stream>>>(set, thr_up, clip_h, clip_w, out);
  return check_launch();
}

}  // namespace b200det
