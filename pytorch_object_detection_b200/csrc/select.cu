// K2 — per-image global top-k over all levels, sorted, with box decode.
//
// Replaces torch.topk + the per-image gathers of FCOSHead.forward (model/modules/head.py:66-80)
// and the score threshold of post_process (head.py:90).  One CTA of 1024 threads per image:
//   1. the image's P scores are read once (coalesced) into registers as order-preserving keys,
//      sub-threshold points become key 0 (threshold-then-top-k selects the same set as the
//      reference's top-k-then-threshold);
//   2. the k-th largest key is found by a bitwise search over the key bits below the common
//      prefix of min/max, one block-wide count per bit, stopping as soon as a count equals k;
//      ties on the k-th key are broken by the lowest point index (a second bitwise search);
//   3. the selected (key, ~index) pairs are bitonic-sorted in shared memory: score descending,
//      point index ascending on equal scores (torch.topk leaves tie order unspecified);
//   4. each selected point is decoded: box = (x - l, y - t, x + r, y + b) with
//      (x, y) = (j*s + s/2, i*s + s/2) computed from the index (head.py:29-38, utills.py:58-73),
//      class = argmax + 1, and — for K3 — the batched_nms coordinate-trick boxes are prepared.
// The working set (4 bytes per point) is L2-resident right after K1; this kernel is latency-bound.
#include "common.cuh"

B200DET_TRACE_BUFFER(select)

#include "select_body.cuh"

namespace b200det {
namespace {

template <bool REG>
__global__ void __launch_bounds__(kSelThreads, 1)
select_topk_kernel(const LevelTable lt, const float* __restrict__ score, const int16_t* __restrict__ cls0,
                   const float thr, const int max_box, const CandSet out, int32_t* __restrict__ cand_point,
                   const unsigned sort_bytes) {
  extern __shared__ __align__(16) unsigned long long sortbuf[];
  unsigned* hist = REG ? reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(sortbuf) + sort_bytes) : nullptr;
  select_topk_cta<REG>(lt, score, cls0, thr, max_box, out, cand_point, blockIdx.x, sortbuf, false, hist);
}

}  // namespace

int launch_select_topk(const LevelTable& lt, int batch, const float* score, const int16_t* cls0, float thr,
                       int max_box, const CandSet& out, int32_t* cand_point, cudaStream_t stream) {
  const int k = max_box < lt.num_points ? max_box : lt.num_points;
  const unsigned sort_bytes = (unsigned)select_sort_bytes(k);
  if (lt.num_points <= kSelItems * kSelThreads) {
    const size_t smem = select_smem_bytes(k, true);
    if (smem > 40 * 1024)
      cudaFuncSetAttribute(select_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    select_topk_kernel<true><<<batch, kSelThreads, smem, stream>>>(lt, score, cls0, thr, max_box, out, cand_point, sort_bytes);
  } else {
    const size_t smem = select_smem_bytes(k, false);
    if (smem > 40 * 1024)
      cudaFuncSetAttribute(select_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    select_topk_kernel<false><<<batch, kSelThreads, smem, stream>>>(lt, score, cls0, thr, max_box, out, cand_point, sort_bytes);
  }
  return check_launch();
}

}  // namespace b200det
