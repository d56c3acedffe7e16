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
  __shared__ unsigned short s_big[B200DET_MAX_BOX / 64 + 1];     // segments longer than 64 boxes: fewer than n / 64
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
  // shared memory: keys [n2] u32 | segment starts [n2] u16 | keep bitmap [n2 / 32] u32 | per-warp column staging
  // | tile bits of the class being resolved cooperatively
  unsigned* keys = reinterpret_cast<unsigned*>(smem_raw);
  unsigned short* seg = reinterpret_cast<unsigned short*>(keys + n2);
  unsigned* keepbits = reinterpret_cast<unsigned*>(seg + n2);
  const int kwords = (n2 + 31) / 32;
  float4* cbox_all = reinterpret_cast<float4*>(smem_raw + (((size_t)n2 * 6 + (size_t)kwords * 4 + 15) / 16) * 16);
  float* carea_all = reinterpret_cast<float*>(cbox_all + kClassWarps * kNmsTile);
  unsigned long long* tmask = reinterpret_cast<unsigned long long*>(carea_all + kClassWarps * kNmsTile);   // [136 tiles][64]
  const size_t o0 = (size_t)b * set.cap;
  const size_t q0 = (size_t)b * out.stride;

  // ---- 1. (class, rank) keys, sorted: one contiguous score-ordered segment per class -----------------------
  // Counting sort on the class (ids below kHistBins; the histogram and the scratch copy of the keys live where the
  // column staging and the tile bits will be): the scan gives every class its slot range, the keys take a slot of
  // their class with one atomic and are then ranked inside the class by their score rank.  5 000 candidates: ~5 us
  // instead of 45 us for the 8 192-key bitonic network, which stays for class ids it cannot bin.
  static_assert(kClassThreads == kHistThreads, "the class histogram is scanned by 1024 threads");
  unsigned* chist = reinterpret_cast<unsigned*>(tmask);                 // [kHistBins] (tmask holds 136 x 64 x 8 bytes)
  unsigned* tmpk = reinterpret_cast<unsigned*>(cbox_all);               // [n] (cbox_all holds 32 x 64 x 16 bytes)
  static_assert((size_t)(kClassMaxSeg / kNmsTile) * (kClassMaxSeg / kNmsTile + 1) / 2 * kNmsTile * 8 >= (size_t)kHistBins * 4,
                "class histogram aliases the tile bits");
  static_assert((size_t)kClassWarps * kNmsTile * sizeof(float4) >= (size_t)B200DET_MAX_BOX * 4, "key scratch aliases the column staging");
  {
    uint4* h4 = reinterpret_cast<uint4*>(chist);
#pragma unroll
    for (int q = 0; q < kHistPerThread / 4; ++q) h4[tid + q * kClassThreads] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  float cmax = 0.f, cbig = 0.f;
  for (int i = tid; i < n2; i += kClassThreads) {
    unsigned key = 0xffffffffu;
    if (i < n) {
      const int c = set.cls[o0 + i];
      cmax = fmaxf(cmax, (c < 0 || c >= (1 << (32 - kRankBits - 1))) ? 1.f : 0.f);
      cbig = fmaxf(cbig, (c < 0 || c >= kHistBins) ? 1.f : 0.f);
      key = ((unsigned)c << kRankBits) | (unsigned)i;
      if (c >= 0 && c < kHistBins) atomicAdd(&chist[hist_slot(c)], 1u);
    }
    keys[i] = key;
  }
  for (int i = tid; i < kwords; i += kClassThreads) keepbits[i] = 0u;
  const bool bad_class = block_max(cmax, s_fmax) > 0.f;          // also the barrier before the sort
  const bool unbinned = block_max(cbig, s_fmax) > 0.f;
  B200DET_STAMP(1);
  if (!unbinned) {
    unsigned hb[kHistPerThread];                                  // thread t: classes [16 t, 16 t + 16), ascending
    int mine = 0;
#pragma unroll
    for (int q = 0; q < kHistPerThread; ++q) {
      hb[q] = chist[q * kClassThreads + tid];
      mine += (int)hb[q];
    }
    int total;
    unsigned start = (unsigned)block_exclusive_scan(mine, s_scan, &total);
#pragma unroll
    for (int q = 0; q < kHistPerThread; ++q) {                   // counts -> first slot of the class
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
      const unsigned first = c == 0 ? 0u : chist[hist_slot(c - 1)];     // = end of the classes below
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
  int n_seg = 0;
  float longest = 0.f;
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
  if (bad_class || longest > (float)kClassMaxSeg) return;        // left to the dense path (mode stays vanilla)
  int n_big = 0;                                                 // classes of more than one 64-box block
  for (int base = 0; base < n_seg; base += kClassThreads) {
    const int sg = base + tid;
    const bool big = sg < n_seg && ((sg + 1 < n_seg ? (int)seg[sg + 1] : n) - (int)seg[sg]) > kNmsTile;
    int total;
    const int pos = n_big + block_exclusive_scan(big ? 1 : 0, s_scan, &total);
    if (big) s_big[pos] = (unsigned short)sg;
    n_big += total;
  }
  cluster.sync();                                                // CTA 0's keep bitmap is clear; every CTA runs
  unsigned* image_keep = cluster.map_shared_rank(keepbits, 0);   // kept ranks meet in CTA 0 (32-bit atomicOr, DSMEM)

  B200DET_STAMP(3);
  // ---- 2. one warp per class segment ---------------------------------------------------------------------------
  float4* cbox = cbox_all + warp * kNmsTile;
  float* carea = carea_all + warp * kNmsTile;
  const float4 kNoBox = make_float4(CUDART_INF_F, CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);   // overlaps nothing
  const float kNoArea = __int_as_float(0x7fc00000);                                             // NaN: never suppresses
  auto load_box = [&](const int pos, const int end) {
    return pos < end ? reinterpret_cast<const float4*>(set.box)[o0 + (keys[pos] & ((1u << kRankBits) - 1u))] : kNoBox;
  };
  auto stage_cols = [&](const float4 c0, const float4 c1, const bool ok0, const bool ok1) {
    __syncwarp();
    cbox[lane] = c0;
    cbox[lane + 32] = c1;
    carea[lane] = ok0 ? __fmul_rn(__fsub_rn(c0.z, c0.x), __fsub_rn(c0.w, c0.y)) : kNoArea;
    carea[lane + 32] = ok1 ? __fmul_rn(__fsub_rn(c1.z, c1.x), __fsub_rn(c1.w, c1.y)) : kNoArea;
    __syncwarp();
  };
  auto rank_of = [&](const int pos) { return keys[pos] & ((1u << kRankBits) - 1u); };
  auto keep_rank = [&](const unsigned r) { atomicOr(image_keep + (r >> 5), 1u << (r & 31)); };

  // 2a. classes of more than 64 boxes, one at a time per CTA (dealt over the cluster): all warps evaluate the
  //     class's tiles into shared memory, then one warp runs the greedy pass over the stored bits.
  for (int k = rank; k < n_big; k += kClassCluster) {
    const int sg = s_big[k];
    const int s0 = seg[sg], s1 = sg + 1 < n_seg ? (int)seg[sg + 1] : n;
    const int W = (s1 - s0 + kNmsTile - 1) / kNmsTile;           // 2 .. 16 blocks
    const int tiles = W * (W + 1) / 2;
    // work unit = (tile, 16-column part): a class of 100-400 boxes has 3-28 tiles, too few for 32 warps
    constexpr int kPart = 16, kParts = kNmsTile / kPart;
    for (int i = tid; i < tiles * kNmsTile; i += kClassThreads) tmask[i] = 0ull;
    __syncthreads();
    for (int unit = warp; unit < tiles * kParts; unit += kClassWarps) {
      const int tile = unit / kParts, part = unit - tile * kParts;
      int rb = 0, rem = tile;
      while (rem >= W - rb) { rem -= W - rb; ++rb; }             // row-major upper triangle
      const int cb = rb + rem;
      const int r0 = s0 + rb * kNmsTile + lane, r1 = r0 + 32;
      const int c0 = s0 + cb * kNmsTile + part * kPart + lane;   // (lanes 0-15 stage the part's columns)
      const float4 a0 = load_box(r0, s1), a1 = load_box(r1, s1);
      const float4 cc = lane < kPart ? load_box(c0, s1) : kNoBox;
      __syncwarp();
      if (lane < kPart) {
        cbox[lane] = cc;
        carea[lane] = c0 < s1 ? __fmul_rn(__fsub_rn(cc.z, cc.x), __fsub_rn(cc.w, cc.y)) : kNoArea;
      }
      __syncwarp();
      unsigned long long d0 = 0ull, d1 = 0ull;
      if (r0 < s1) d0 = (unsigned long long)mask_row_bits_part<ZERO_SUP, kPart>(
                            a0, __fmul_rn(__fsub_rn(a0.z, a0.x), __fsub_rn(a0.w, a0.y)), cbox, carea, thr_up) << (part * kPart);
      if (r1 < s1) d1 = (unsigned long long)mask_row_bits_part<ZERO_SUP, kPart>(
                            a1, __fmul_rn(__fsub_rn(a1.z, a1.x), __fsub_rn(a1.w, a1.y)), cbox, carea, thr_up) << (part * kPart);
      if (cb == rb) {                                            // diagonal tile: only later boxes (j > row)
        d0 &= ~((2ull << lane) - 1ull);
        d1 &= ~((2ull << (lane + 32)) - 1ull);
      }
      if (d0) atomicOr(&tmask[(size_t)tile * kNmsTile + lane], d0);
      if (d1) atomicOr(&tmask[(size_t)tile * kNmsTile + lane + 32], d1);
    }
    __syncthreads();
    if (k == rank) { B200DET_STAMP_NOSYNC(10); B200DET_NOTE_IF(blockIdx.x == 0 && blockIdx.y == 0, 12, W); B200DET_NOTE_IF(blockIdx.x == 0 && blockIdx.y == 0, 13, n_big); }
    if (warp == 0) {
      unsigned long long myrem = 0ull;                           // lane w: removed bits of the class's block w
      int tile = 0;
      for (int rb = 0; rb < W; ++rb) {
        const unsigned long long* diag = tmask + (size_t)tile * kNmsTile;
        const int rows = min(kNmsTile, s1 - s0 - rb * kNmsTile);
        const unsigned long long valid = rows == kNmsTile ? ~0ull : ((1ull << rows) - 1ull);
        const unsigned long long keep = resolve_block(shfl64(myrem, rb), valid, diag[lane], diag[lane + 32], lane);
        const bool k0 = (keep >> lane) & 1ull, k1 = (keep >> (lane + 32)) & 1ull;
        if (k0) keep_rank(rank_of(s0 + rb * kNmsTile + lane));
        if (k1) keep_rank(rank_of(s0 + rb * kNmsTile + lane + 32));
        for (int cb = rb + 1; cb < W; ++cb) {
          const unsigned long long* col = tmask + (size_t)(tile + cb - rb) * kNmsTile;
          const unsigned long long v = warp_or64((k0 ? col[lane] : 0ull) | (k1 ? col[lane + 32] : 0ull));
          if (lane == cb) myrem |= v;
        }
        tile += W - rb;
      }
    }
    __syncthreads();
    if (k == rank) B200DET_STAMP_NOSYNC(11);
  }
  B200DET_STAMP_NOSYNC(14);

  // 2b. classes of at most 64 boxes: one warp each, a single diagonal tile, nothing stored
  for (int sg = rank * kClassWarps + warp; sg < n_seg; sg += kClassCluster * kClassWarps) {
    const int s0 = seg[sg], s1 = sg + 1 < n_seg ? (int)seg[sg + 1] : n;
    if (s1 - s0 > kNmsTile) continue;
    const int r0 = s0 + lane, r1 = r0 + 32;
    const bool ok0 = r0 < s1, ok1 = r1 < s1;
    const float4 a0 = load_box(r0, s1), a1 = load_box(r1, s1);
    stage_cols(a0, a1, ok0, ok1);
    unsigned long long d0 = 0ull, d1 = 0ull;
    if (ok0) d0 = mask_row_bits<ZERO_SUP>(a0, carea[lane], 0, cbox, carea, nullptr, thr_up, false);
    if (ok1) d1 = mask_row_bits<ZERO_SUP>(a1, carea[lane + 32], 0, cbox, carea, nullptr, thr_up, false);
    d0 &= ~((2ull << lane) - 1ull);
    d1 &= ~((2ull << (lane + 32)) - 1ull);
    const int rows = s1 - s0;
    const unsigned long long valid = rows == kNmsTile ? ~0ull : ((1ull << rows) - 1ull);
    const unsigned long long keep = resolve_block(0ull, valid, d0, d1, lane);
    if ((keep >> lane) & 1ull) keep_rank(rank_of(r0));
    if ((keep >> (lane + 32)) & 1ull) keep_rank(rank_of(r1));
  }
  cluster.sync();                                                // every kept rank has landed in CTA 0
  B200DET_STAMP(4);
  if (rank != 0) return;

  // ---- 3. kept candidates in score (rank) order -------------------------------------------------------------------
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
  for (int q = tid; q < n; q += kClassThreads) {
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
  if (tid == 0) {
    out.count[b] = run;
    set.mode[b] = kModeDone;                                     // the dense kernels skip this image
  }
}

size_t class_smem_bytes(int cap) {
  int n2 = 1;
  while (n2 < cap) n2 <<= 1;
  const size_t head = (((size_t)n2 * 6 + (size_t)((n2 + 31) / 32) * 4 + 15) / 16) * 16;
  const size_t max_tiles = (size_t)(kClassMaxSeg / kNmsTile) * (kClassMaxSeg / kNmsTile + 1) / 2;
  return head + (size_t)kClassWarps * kNmsTile * (sizeof(float4) + sizeof(float)) + max_tiles * kNmsTile * 8;
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
