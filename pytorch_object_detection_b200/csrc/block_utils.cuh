// CTA-wide primitives used by the select / NMS kernels: scans, reductions, bitonic sort.
#pragma once
#include "common.cuh"

namespace b200det {

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Exclusive scan of one int per thread over the whole CTA (blockDim.x multiple of 32, <= 1024).
// `warp_sums` is 33 ints of shared scratch.  Returns the exclusive prefix; *total gets the sum.
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  __syncthreads();  // protect warp_sums from a previous use
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int s = lane < nwarps ? warp_sums[lane] : 0;
    int si = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, si, d);
      if (lane >= d) si += o;
    }
    warp_sums[lane] = si - s;            // exclusive warp offsets
    if (lane == 31) warp_sums[32] = si;  // grand total
  }
  __syncthreads();
  *total = warp_sums[32];
  return warp_sums[warp] + incl - v;
}

__device__ __forceinline__ float block_max(float v, float* scratch /*32 floats*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = scratch[0];
  for (int i = 1; i < nwarps; ++i) r = fmaxf(r, scratch[i]);
  return r;
}

// In-place DESCENDING bitonic sort of n (power of two) 64-bit keys in shared memory.
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* buf, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        // t-th compare-exchange pair of this stage: insert a 0 bit at position log2(j)
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const unsigned long long a = buf[i], b = buf[p];
        const bool desc = (i & k) == 0;
        if (desc ? (a < b) : (a > b)) {
          buf[i] = b;
          buf[p] = a;
        }
      }
      __syncthreads();
    }
  }
}

// DESCENDING bitonic sort of N (power of two, <= blockDim.x) 64-bit keys held ONE PER THREAD
// (thread t holds element t; threads >= N must pass 0 and still call).  Exchanges at distance
// < 32 are warp shuffles (no barrier); larger distances go through `scratch` (2 * blockDim.x
// keys, double-buffered so each such stage costs one barrier).  For N = 1024: 15 barriers instead
// of the 55 of the all-shared-memory network.  Fully unrolled: every stage's masks are constants.
template <int N>
__device__ __forceinline__ unsigned long long bitonic_sort_desc_regs(unsigned long long v,
                                                                     unsigned long long* scratch) {
  const int t = threadIdx.x;
  int flip = 0;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      unsigned long long pv;
      if (j >= 32) {
        unsigned long long* buf = scratch + (flip & 1) * blockDim.x;
        buf[t] = v;
        __syncthreads();
        pv = buf[t ^ j];
        ++flip;
      } else {
        const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, j);
        const unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), j);
        pv = ((unsigned long long)hi << 32) | lo;
      }
      // lower index of the pair keeps the max in a descending run, the min in an ascending one
      const bool take_max = ((t & k) == 0) == ((t & j) == 0);
      const bool pv_gt = pv > v;
      v = (take_max == pv_gt) ? pv : v;
    }
  }
  return v;
}

// ---- 16 384-bin shared-memory histogram shared by the select (select_body.cuh) and the NMS sorts (nms.cu,
// nms_class.cu); CTAs of 1024 threads, thread t scans 16 consecutive bins -----------------------------------------
constexpr int kHistBins = 16384;
constexpr int kHistThreads = 1024;
constexpr int kHistPerThread = kHistBins / kHistThreads;
static_assert(kHistPerThread == 16, "a thread scans 16 bins");
// Linear score bin of an order_key: bin = trunc(score * kHistBins) — exact in fp32 (a power of two), so "bin >= b" is
// "score >= b / kHistBins" and a bin found by a scan IS a key threshold.
__device__ __forceinline__ int hist_bin(uint32_t key) {
  return min(kHistBins - 1, max(0, (int)(key_to_float(key) * (float)kHistBins)));
}
// Where bin b lives in shared memory: TRANSPOSED, so that a thread walking its 16 consecutive bins and the lanes of a
// warp walking theirs touch 32 different banks (bins stored linearly: a 16-word stride between lanes, every access a
// 16-way bank conflict — measured in the select: scan 1.5 us, slot ranges 2.2 us instead of 0.6 / 1.0 us).
__device__ __forceinline__ int hist_slot(int bin) { return (bin & (kHistPerThread - 1)) * kHistThreads + (bin >> 4); }

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace b200det
