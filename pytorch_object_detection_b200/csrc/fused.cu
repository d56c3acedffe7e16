// K2 + K3 fused for max_detection_box <= 1024: ONE CTA per image does select -> mask -> greedy.
//
// The three-kernel chain select -> mask -> scan spends most of its time on per-image single-CTA
// latency, two launch boundaries and a global round trip of a mask that is almost entirely zero.
// The dense mask evaluates all n^2/2 pairs, yet with torchvision's coordinate trick two boxes of
// DIFFERENT classes can only overlap when the higher-class box reaches far enough into negative
// coordinates to bridge the class offset:
//     nms_box = box + cls * S,  S = max_coord + 1  (all in fp32)
//     for cls_i < cls_j:  overlap in x needs  x2_i + cls_i*S > x1_j + cls_j*S,  and x2_i <= max_coord,
//     so x1_j < -1 + eps  (eps bounds the fp32 rounding of the offsets and sums); same for y.
// Hence only (a) same-class pairs and (b) pairs involving a "wildcard" box with x1 < theta and
// y1 < theta (theta = -1 + eps) can ever set a mask bit; every other pair has zero intersection and
// the reference's IoU is 0 (never > thr >= 0).  On the per-class (vanilla) branch only (a) exists.
// The CTA therefore buckets the sorted candidates by class in shared memory and evaluates the
// reference's exact IoU expression only for those pairs (~n^2 / (2 * classes) instead of n^2 / 2),
// writing the bits into a shared-memory mask that never leaves the SM; the greedy pass and the
// output writer are the ones of nms_scan_smem_kernel.  Results are bit-identical to the dense path
// (tests run both).  thr < 0 (zero IoU suppresses) is left to the dense path.
#include "common.cuh"

B200DET_TRACE_BUFFER(fused)

#include "nms_body.cuh"
#include "select_body.cuh"

namespace b200det {
namespace {

constexpr int kFusedMaxBlocks = 16;                           // candidates <= 1024
constexpr int kFusedMaxBox = kFusedMaxBlocks * kNmsTile;
constexpr int kBuckets = 1024;                                // class id & 1023
constexpr size_t kSortBytes = 2 * kSelThreads * sizeof(unsigned long long);          // 16 KB (K2 scratch)
constexpr size_t kMaskWords = (size_t)kNmsTile * (kFusedMaxBlocks * (kFusedMaxBlocks + 1) / 2);   // 8704 words, 68 KB
constexpr size_t kFusedSmem = kSortBytes + kMaskWords * 8 + kFusedMaxBox * (2 * (sizeof(float4) + 4 + 4) + 2 + 2) +
                              (kBuckets + 1) * 4 + kBuckets * 4;
static_assert(kBuckets == kSelThreads, "one bucket counter per thread");

__device__ __forceinline__ float ulp_of(float v) { return __uint_as_float(__float_as_uint(fabsf(v)) + 1u) - fabsf(v); }

// the reference's IoU test on NMS boxes a (earlier, higher score) and c: torchvision nms_kernel_impl
__device__ __forceinline__ bool suppresses(const float4 a, const float aarea, const float4 c, const float carea,
                                           const float thr_up) {
  const float xx1 = fmaxf(a.x, c.x), yy1 = fmaxf(a.y, c.y), xx2 = fminf(a.z, c.z), yy2 = fminf(a.w, c.w);
  if (!(xx2 > xx1 && yy2 > yy1)) return false;                // zero intersection: IoU is 0 (or NaN), never > thr >= 0
  const float w = fmaxf(0.f, __fsub_rn(xx2, xx1));
  const float h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  return iou_reaches(inter, __fsub_rn(__fadd_rn(aarea, carea), inter), iou_midpoint(thr_up));   // == (double)ovr > thr
}

// FROM_SET: the candidate rows come from the candidate set that nms_prepare_kernel wrote (the stand-alone batched_nms
// entry, <= 1024 candidates) instead of from the select; everything after the select is the same code.
template <bool REG, bool FROM_SET = false>
__global__ void __launch_bounds__(kSelThreads, 1)
fused_select_nms_kernel(const LevelTable lt, const float* __restrict__ score, const int16_t* __restrict__ cls0,
                        const float thr, const int max_box, const CandSet set, const float thr_up,
                        const int clip_h, const int clip_w, const NmsOut out) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned long long* sortbuf = reinterpret_cast<unsigned long long*>(smem);
  unsigned long long* maskT = reinterpret_cast<unsigned long long*>(smem + kSortBytes);
  float4* sbox = reinterpret_cast<float4*>(maskT + kMaskWords);          // NMS boxes, candidate order
  float* sarea = reinterpret_cast<float*>(sbox + kFusedMaxBox);
  int* scls = reinterpret_cast<int*>(sarea + kFusedMaxBox);
  float4* obox = reinterpret_cast<float4*>(scls + kFusedMaxBox);         // the same three arrays in bucket order
  float* oarea = reinterpret_cast<float*>(obox + kFusedMaxBox);
  int* ocls = reinterpret_cast<int*>(oarea + kFusedMaxBox);
  unsigned short* order = reinterpret_cast<unsigned short*>(ocls + kFusedMaxBox);   // candidates grouped by bucket
  unsigned short* wild = order + kFusedMaxBox;                                      // wildcard candidates
  int* bstart = reinterpret_cast<int*>(wild + kFusedMaxBox);             // [kBuckets + 1] bucket offsets
  int* bfill = bstart + kBuckets + 1;                                    // [kBuckets] counters
  __shared__ unsigned long long keepw[32];
  __shared__ unsigned nz[32];
  __shared__ int s_pre[33];
  __shared__ int s_scan[33];
  __shared__ float s_f[32];
  __shared__ int s_nwild;

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t o0 = (size_t)b * set.cap;
  const size_t q0 = (size_t)b * out.stride;

  // ---- K2: select, sort, decode.  The candidate set goes to global memory for the writer; this
  // thread keeps candidate row `tid` (raw box, class) in registers. -----------------------------------
  CandSet sel_out = set;
  sel_out.nms_box = nullptr;                                  // NMS boxes stay in registers / shared memory here
  // (the score histogram of the select lives where the suppression mask will be: 64 KB of the 68 KB)
  static_assert(kMaskWords * 8 >= (size_t)kHistBins * sizeof(unsigned), "histogram aliases the mask");
  SelectResult sel;
  int mysrc = tid;                                            // what the keep output names: the candidate's row
  if constexpr (FROM_SET) {
    sel.count = min(set.count[b], kFusedMaxBox);
    sel.box = make_float4(0.f, 0.f, 0.f, 0.f);
    sel.cls = -1;
    sel.score = 0.f;
    float vmax = -CUDART_INF_F;
    if (tid < sel.count) {
      sel.box = reinterpret_cast<const float4*>(set.box)[o0 + tid];
      sel.cls = set.cls[o0 + tid];
      sel.score = set.score[o0 + tid];
      mysrc = set.src[o0 + tid];                              // ... in the caller's thresholded list
      vmax = fmaxf(fmaxf(sel.box.x, sel.box.y), fmaxf(sel.box.z, sel.box.w));
    }
    sel.max_coord = block_max(vmax, s_f);
  } else {
    sel = select_topk_cta<REG>(lt, score, cls0, thr, max_box, sel_out, nullptr, b, sortbuf, true,
                               REG ? reinterpret_cast<unsigned*>(maskT) : nullptr, false);
  }
  const int n = sel.count;
  const int W = (n + kNmsTile - 1) / kNmsTile;
  B200DET_STAMP(8);

  // ---- sparse mask in shared memory ---------------------------------------------------------------
  for (int i = tid; i < (int)kMaskWords; i += kSelThreads) maskT[i] = 0ull;
  bfill[tid] = 0;
  if (tid < 32) { keepw[tid] = 0ull; nz[tid] = 0u; }
  if (tid == 0) s_nwild = 0;
  const bool trick = (n * 4 <= kTrickMaxNumel);               // torchvision batched_nms, CPU rule (nms.cuh)
  const float S = __fadd_rn(sel.max_coord, 1.0f);             // coordinate-trick span, as nms_prepare_boxes
  const float4 raw = sel.box;
  const int mycls = (tid < n) ? sel.cls : -1;
  float4 mybox = raw;
  if (trick) {
    const float off = __fmul_rn((float)mycls, S);
    mybox.x = __fadd_rn(raw.x, off);
    mybox.y = __fadd_rn(raw.y, off);
    mybox.z = __fadd_rn(raw.z, off);
    mybox.w = __fadd_rn(raw.w, off);
  }
  const float myarea = __fmul_rn(__fsub_rn(mybox.z, mybox.x), __fsub_rn(mybox.w, mybox.y));
  if (tid < n) {
    sbox[tid] = mybox;
    sarea[tid] = myarea;
    scls[tid] = mycls;
  }
  if (tid == 0) set.mode[b] = trick ? kModeTrick : kModeVanilla;
  B200DET_STAMP(12);
  const float cmax = block_max((float)mycls, s_f);            // its barriers also order the zero-fill before the atomics
  bool mywild = false;
  if (tid < n) {
    atomicAdd(&bfill[mycls & (kBuckets - 1)], 1);
    if (trick) {
      // wildcard: raw x1 and y1 both below theta = -1 + eps (see header; NaN counts as below).  eps
      // covers the rounding of S = max + 1, of cls * S and of the coordinate sums, with a factor 2 to spare.
      const float eps = 4.f * ulp_of(cmax * S + S) + ulp_of(S);
      const float theta = -1.f + 2.f * eps;
      mywild = !(raw.x >= theta) && !(raw.y >= theta);
      if (mywild) wild[atomicAdd(&s_nwild, 1)] = (unsigned short)tid;
    }
  }
  __syncthreads();
  {
    int total;
    const int start = block_exclusive_scan(bfill[tid], s_scan, &total);
    bstart[tid] = start;
    if (tid == 0) bstart[kBuckets] = total;
    __syncthreads();
    bfill[tid] = 0;
  }
  __syncthreads();
  if (tid < n) {
    const int bk = mycls & (kBuckets - 1);
    const int slot = bstart[bk] + atomicAdd(&bfill[bk], 1);
    order[slot] = (unsigned short)tid;
    obox[slot] = mybox;
    oarea[slot] = myarea;
    ocls[slot] = mycls;
  }
  __syncthreads();
  B200DET_STAMP(13);
  {
    // A set bit is rare, so all mask updates are shared-memory atomics (rows are not thread-private).
    auto mark = [&](const int a, const int c) {
      const int lo = min(a, c), hi = max(a, c);
      atomicOr(&maskT[col_off(hi >> 6) + lo], 1ull << (hi & 63));
      atomicOr(&nz[lo >> 6], 1u << (hi >> 6));
    };
    // same-class pairs: thread t takes the candidate in bucket-ordered SLOT t and tests it against the later slots
    // of its bucket — the lanes of a warp then sit in the same one or two buckets and run the same number of
    // iterations (with one thread per candidate ROW every warp ran as long as its largest class)
    if (tid < n) {
      const int me = order[tid], c_me = ocls[tid];
      const float4 b_me = obox[tid];
      const float a_me = oarea[tid];
      for (int e = tid + 1, e1 = bstart[(c_me & (kBuckets - 1)) + 1]; e < e1; ++e)
        if (ocls[e] == c_me && suppresses(b_me, a_me, obox[e], oarea[e], thr_up)) mark(me, order[e]);   // symmetric
    }
    // wildcard pairs (trick branch): every candidate against every wildcard box of another class
    if (trick && tid < n) {
      const int nw = s_nwild;
      for (int e = 0; e < nw; ++e) {
        const int j = wild[e];
        // a pair of two wildcards is found from both sides: the bit is the same
        if (j != tid && scls[j] != mycls && suppresses(mybox, myarea, sbox[j], sarea[j], thr_up)) mark(tid, j);
      }
    }
  }
  __syncthreads();
  B200DET_STAMP(9);

  // ---- greedy pass (one warp) + outputs --------------------------------------------------------------
  if (warp == 0) {
    unsigned long long myrem = 0ull;                          // lane w: removed-bitmap word w
    for (int rb = 0; rb < W; ++rb) {
      const int rows = min(kNmsTile, n - rb * kNmsTile);
      const unsigned long long valid = rows == kNmsTile ? ~0ull : ((1ull << rows) - 1ull);
      const unsigned long long cur = shfl64(myrem, rb);
      const unsigned flags = nz[rb];
      unsigned long long keep;
      if (!((flags >> rb) & 1u)) {
        keep = ~cur & valid;                                  // empty diagonal tile
      } else {
        const unsigned long long* diag = maskT + col_off(rb) + rb * kNmsTile;
        keep = resolve_block(cur, valid, diag[lane], diag[lane + 32], lane);
      }
      if (lane == 0) keepw[rb] = keep;
      const bool k0 = (keep >> lane) & 1ull, k1 = (keep >> (lane + 32)) & 1ull;
      for (unsigned m = flags & ~((2u << rb) - 1u); m;) {     // flagged column blocks ahead, four independent
        int w[4];                                             //   chains of loads and warp reductions at a time
        unsigned long long v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          w[q] = m ? __ffs((int)m) - 1 : -1;
          m &= m - 1u;                                        // (0 stays 0)
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          v[q] = 0ull;
          if (w[q] >= 0) {
            const unsigned long long* col = maskT + col_off(w[q]) + rb * kNmsTile;
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
    __syncwarp();
    const int cnt = (lane < W) ? __popcll(keepw[lane]) : 0;   // exclusive prefix of kept counts per block
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
  B200DET_STAMP(10);
  if (tid < n) {                                              // one candidate per thread (n <= 1024)
    const unsigned long long kw = keepw[tid >> 6];
    if ((kw >> (tid & 63)) & 1ull) {
      const int o = s_pre[tid >> 6] + __popcll(kw & ((1ull << (tid & 63)) - 1ull));
      store_kept(set, out, o0, q0, tid, o, clip_h, clip_w, raw, sel.score, mycls, mysrc);
    }
  }
  B200DET_STAMP(11);
  if (tid == 0) out.count[b] = s_pre[32];
}

}  // namespace

// the NMS half of the fused kernel on a candidate set made by nms_prepare_kernel (cap <= 1024, nms_thr >= 0)
bool fused_nms_supported(int cap, double nms_thr) { return cap <= kFusedMaxBox && nms_thr >= 0.0; }

int launch_fused_nms_from_set(const CandSet& set, int batch, double nms_thr, int clip_h, int clip_w, const NmsOut& out,
                              cudaStream_t stream) {
  float thr_up;
  bool zero_sup;
  nms_threshold_params(nms_thr, &thr_up, &zero_sup);
  if (zero_sup || set.cap > kFusedMaxBox) return B200DET_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(fused_select_nms_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kFusedSmem);
  if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
  LevelTable none = {};
  fused_select_nms_kernel<true, true><<<batch, kSelThreads, kFusedSmem, stream>>>(none, nullptr, nullptr, 0.f, 0, set, thr_up,
                                                                                 clip_h, clip_w, out);
  return check_launch();
}

bool fused_supported(const LevelTable& lt, int max_box, double nms_thr) {
  const int k = max_box < lt.num_points ? max_box : lt.num_points;
  return k <= kFusedMaxBox && nms_thr >= 0.0;
}

int launch_fused_select_nms(const LevelTable& lt, int batch, const float* score, const int16_t* cls0, float thr,
                            int max_box, const CandSet& set, double nms_thr, int clip_h, int clip_w,
                            const NmsOut& out, cudaStream_t stream) {
  float thr_up;
  bool zero_sup;
  nms_threshold_params(nms_thr, &thr_up, &zero_sup);
  if (zero_sup) return B200DET_ERR_UNSUPPORTED;
  cudaError_t e;
  if (lt.num_points <= kSelItems * kSelThreads) {
    e = cudaFuncSetAttribute(fused_select_nms_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem);
    if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
    fused_select_nms_kernel<true><<<batch, kSelThreads, kFusedSmem, stream>>>(lt, score, cls0, thr, max_box, set, thr_up,
                                                                             clip_h, clip_w, out);
  } else {
    e = cudaFuncSetAttribute(fused_select_nms_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem);
    if (e != cudaSuccess) { set_cuda_error(e); return B200DET_ERR_CUDA; }
    fused_select_nms_kernel<false><<<batch, kSelThreads, kFusedSmem, stream>>>(lt, score, cls0, thr, max_box, set, thr_up,
                                                                              clip_h, clip_w, out);
  }
  return check_launch();
}

}  // namespace b200det
