// Shared device/host helpers for the b200det kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/b200det.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "b200det is written for sm_100a (B200) only"
#endif

namespace b200det {

constexpr int kTileThreads = 128;        // streaming kernels: 128 threads x 4 points
constexpr int kTilePts = 4;
constexpr int kTile = kTileThreads * kTilePts;   // 512 points of one level of one image per CTA
constexpr int kNmsTile = 64;             // boxes per NMS mask word

// Kernel-side level table (passed by value in kernel params).
struct LevelTable {
  const float* cls[B200DET_MAX_LEVELS];
  const float* cnt[B200DET_MAX_LEVELS];
  const float* reg[B200DET_MAX_LEVELS];
  const float* reg_scale[B200DET_MAX_LEVELS];   // ScaleExp folded in: distances = exp(reg * *reg_scale); NULL = reg is final
  int h[B200DET_MAX_LEVELS];
  int w[B200DET_MAX_LEVELS];
  int stride[B200DET_MAX_LEVELS];
  int hw[B200DET_MAX_LEVELS];
  int vec_ok[B200DET_MAX_LEVELS];          // hw % 4 == 0 and every map pointer 16-byte aligned
  int point_off[B200DET_MAX_LEVELS + 1];   // prefix sum of hw (level-major point index)
  int tile_off[B200DET_MAX_LEVELS + 1];    // prefix sum of ceil(hw / kTile)
  int n_levels;
  int num_points;
  int cls_dtype;                           // b200det_dtype of the cls AND cnt maps (b200det_level.dtypes, all levels agree)
  int reg_dtype;                           // b200det_dtype of the reg maps
};

// Gradient destinations, one per level (same NCHW shape as the map they belong to).
struct GradTable {
  float* g[B200DET_MAX_LEVELS];
};

// Upstream gradient dL/d(loss[b]) of image b.  grad NULL: 1/B (the gradient of the batch mean for an upstream
// of 1); grad_mode 0: grad[b]; grad_mode 1: grad[0] is the upstream gradient of the batch MEAN -> grad[0] / B.
__device__ __forceinline__ float upstream_of(const float* grad, const int grad_mode, const int b, const float inv_batch) {
  if (!grad) return inv_batch;
  return grad_mode ? __ldcg(grad) * inv_batch : grad[b];
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Returns false on bad arguments.
inline bool make_level_table(const b200det_level* levels, int n_levels, LevelTable* t) {
  if (!levels || n_levels <= 0 || n_levels > B200DET_MAX_LEVELS) return false;
  t->cls_dtype = levels[0].dtypes & 15;
  t->reg_dtype = (levels[0].dtypes >> 4) & 15;
  if (t->cls_dtype > B200DET_BF16 || t->reg_dtype > B200DET_BF16 || (levels[0].dtypes >> 8)) return false;
  long long off = 0;
  int toff = 0;
  for (int l = 0; l < n_levels; ++l) {
    if (levels[l].h <= 0 || levels[l].w <= 0 || levels[l].stride <= 0 || levels[l].dtypes != levels[0].dtypes) return false;
    t->cls[l] = static_cast<const float*>(levels[l].cls);
    t->cnt[l] = static_cast<const float*>(levels[l].cnt);
    t->reg[l] = static_cast<const float*>(levels[l].reg);
    t->reg_scale[l] = static_cast<const float*>(levels[l].reg_scale);
    t->h[l] = levels[l].h;
    t->w[l] = levels[l].w;
    t->stride[l] = levels[l].stride;
    t->hw[l] = levels[l].h * levels[l].w;
    // 4 consecutive elements per access: 16 bytes of fp32, 8 bytes of fp16 / bf16 (16-byte aligned bases keep every
    // plane aligned when hw % 4 == 0)
    t->vec_ok[l] = (t->hw[l] % 4 == 0) && aligned16(levels[l].cls) && aligned16(levels[l].cnt) &&
                   aligned16(levels[l].reg);
    t->point_off[l] = static_cast<int>(off);
    t->tile_off[l] = toff;
    off += t->hw[l];
    toff += (t->hw[l] + kTile - 1) / kTile;
    if (off > (1ll << 30)) return false;
  }
  for (int l = n_levels; l <= B200DET_MAX_LEVELS; ++l) {
    t->point_off[l] = static_cast<int>(off);
    t->tile_off[l] = toff;
  }
  for (int l = n_levels; l < B200DET_MAX_LEVELS; ++l) {
    t->cls[l] = t->cnt[l] = t->reg[l] = t->reg_scale[l] = nullptr;
    t->h[l] = t->w[l] = t->stride[l] = t->hw[l] = t->vec_ok[l] = 0;
  }
  t->n_levels = n_levels;
  t->num_points = static_cast<int>(off);
  return true;
}

// entry points without half-precision kernels: every level must declare fp32 maps
inline bool fp32_levels(const b200det_level* levels, int n_levels) {
  for (int l = 0; levels && l < n_levels; ++l)
    if (levels[l].dtypes != 0) return false;
  return true;
}

void set_cuda_error(cudaError_t e);

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_cuda_error(e);
    return B200DET_ERR_CUDA;
  }
  return B200DET_OK;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- optional phase tracing (build with -DB200DET_TRACE; scripts/trace_phases.py reads it) --------
#ifdef B200DET_TRACE
#define B200DET_TRACE_BUFFER(name)                                                   \
  namespace b200det { static __device__ long long g_trace[64]; static __device__ long long g_cta[4096 * 2]; } \
  extern "C" int b200det_debug_read_cta_trace_##name(long long* host_out, int n) {   \
    return (int)cudaMemcpyFromSymbol(host_out, b200det::g_cta, sizeof(long long) * (n > 8192 ? 8192 : n)); \
  }                                                                                  \
  extern "C" int b200det_debug_read_trace_##name(long long* host_out, int n) {       \
    return (int)cudaMemcpyFromSymbol(host_out, b200det::g_trace, sizeof(long long) * (n > 64 ? 64 : n)); \
  }                                                                                  \
  extern "C" int b200det_debug_reset_trace_##name(void) {                            \
    long long z_[64];                                                                \
    for (int i = 0; i < 64; ++i) z_[i] = 0;                                          \
    z_[60] = 0x7fffffffffffffffll;                                                   \
    return (int)cudaMemcpyToSymbol(b200det::g_trace, z_, sizeof(z_));                \
  }
#define B200DET_STAMP(slot)                                                          \
  do {                                                                               \
    __syncthreads();                                                                 \
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_trace[slot] = clock64(); \
  } while (0)
#define B200DET_STAMP_NOSYNC(slot)                                                   \
  do {                                                                               \
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_trace[slot] = clock64(); \
  } while (0)
#define B200DET_STAMP_IF(cond, slot)                                                 \
  do {                                                                               \
    if ((cond) && threadIdx.x == 0) g_trace[slot] = clock64();                       \
  } while (0)
#define B200DET_NOTE_IF(cond, slot, v)                                               \
  do {                                                                               \
    if ((cond) && threadIdx.x == 0) g_trace[slot] = (long long)(v);                  \
  } while (0)
#define B200DET_STAMP_ANY(cond, slot)                                                \
  do {                                                                               \
    if (cond) g_trace[slot] = clock64();                                             \
  } while (0)
/* kernel span in ns of %globaltimer: slot 60 = first CTA start (min), 61 = last CTA end (max), 62 = CTAs that   \
   recounted; reset by reading the trace */                                          \
#define B200DET_SPAN_BEGIN()                                                         \
  do {                                                                               \
    if (threadIdx.x == 0) {                                                          \
      unsigned long long t_;                                                         \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                         \
      atomicMin(reinterpret_cast<unsigned long long*>(&g_trace[60]), t_);            \
      const unsigned c_ = blockIdx.x + gridDim.x * blockIdx.y;                       \
      if (c_ < 4096) g_cta[2 * c_] = (long long)t_;                                  \
    }                                                                                \
  } while (0)
#define B200DET_SPAN_END()                                                           \
  do {                                                                               \
    if (threadIdx.x == 0) {                                                          \
      unsigned long long t_;                                                         \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                         \
      atomicMax(reinterpret_cast<unsigned long long*>(&g_trace[61]), t_);            \
      const unsigned c_ = blockIdx.x + gridDim.x * blockIdx.y;                       \
      if (c_ < 4096) g_cta[2 * c_ + 1] = (long long)t_;                              \
    }                                                                                \
  } while (0)
#else
#define B200DET_NOTE_IF(cond, slot, v) do {} while (0)
#define B200DET_STAMP_ANY(cond, slot) do {} while (0)
#define B200DET_SPAN_BEGIN() do {} while (0)
#define B200DET_SPAN_END() do {} while (0)
#define B200DET_STAMP_IF(cond, slot) do {} while (0)
#define B200DET_TRACE_BUFFER(name)
#define B200DET_STAMP(slot) do {} while (0)
#define B200DET_STAMP_NOSYNC(slot) do {} while (0)
#endif

// ---- device helpers ---------------------------------------------------------------------
// sigmoid as torch computes it: 1 / (1 + exp(-x)), IEEE division, full-precision expf.
// (__frcp_rn is the correctly rounded reciprocal, i.e. bit-identical to 1.0f / y.)
__device__ __forceinline__ float sigmoid_f32(float x) { return __frcp_rn(__fadd_rn(1.0f, expf(-x))); }

// ScaleExp (modules.py:170-176): exp(x * scale), each op rounded as torch rounds it
__device__ __forceinline__ float scale_exp_f32(float x, float scale) { return expf(__fmul_rn(x, scale)); }

// streaming (read-once) loads that do not allocate in L1
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 r;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream_f1(const float* p) {
  float r;
  asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
// 4 consecutive 16-bit elements (8 bytes) / one element, streaming, as raw bits
__device__ __forceinline__ uint2 ldg_stream_b64(const void* p) {
  uint2 r;
  asm("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ unsigned short ldg_stream_b16(const void* p) {
  unsigned short r;
  asm("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
  return r;
}

// Element access of a map that may be fp32, fp16 or bf16 (fp32 arithmetic everywhere: a half value is up-cast
// exactly, so a kernel reading half maps equals the fp32 kernel on up-cast maps bit for bit).
template <typename T> struct MapElem;
template <> struct MapElem<float> {
  static __device__ __forceinline__ float4 load4(const void* base, size_t i) { return ldg_stream_f4(static_cast<const float*>(base) + i); }
  static __device__ __forceinline__ float load1(const void* base, size_t i) { return ldg_stream_f1(static_cast<const float*>(base) + i); }
};
template <> struct MapElem<__half> {
  static __device__ __forceinline__ float4 load4(const void* base, size_t i) {
    const uint2 r = ldg_stream_b64(static_cast<const __half*>(base) + i);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ float load1(const void* base, size_t i) {
    const unsigned short r = ldg_stream_b16(static_cast<const __half*>(base) + i);
    return __half2float(*reinterpret_cast<const __half*>(&r));
  }
};
template <> struct MapElem<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load4(const void* base, size_t i) {
    const uint2 r = ldg_stream_b64(static_cast<const __nv_bfloat16*>(base) + i);
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                       __uint_as_float(r.y & 0xffff0000u));
  }
  static __device__ __forceinline__ float load1(const void* base, size_t i) {
    return __uint_as_float((unsigned)ldg_stream_b16(static_cast<const __nv_bfloat16*>(base) + i) << 16);
  }
};
// one element of a map whose type is only known at run time (gathers of a few points)
__device__ __forceinline__ float load_map_elem(const void* base, const int dtype, const size_t i) {
  if (dtype == B200DET_F16) return __half2float(static_cast<const __half*>(base)[i]);
  if (dtype == B200DET_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(base)[i]);
  return static_cast<const float*>(base)[i];
}

// streaming (write-once) stores
__device__ __forceinline__ void stg_stream_f4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream_f1(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// level of a level-major point index / of a tile index (branch-free over the small table)
__device__ __forceinline__ int level_of_point(const LevelTable& t, int p) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < B200DET_MAX_LEVELS; ++i) l += (i < t.n_levels && p >= t.point_off[i]) ? 1 : 0;
  return l;
}
__device__ __forceinline__ int level_of_tile(const LevelTable& t, int tile) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < B200DET_MAX_LEVELS; ++i) l += (i < t.n_levels && tile >= t.tile_off[i]) ? 1 : 0;
  return l;
}

// order-preserving map float -> uint32 (larger float <=> larger key); -0 is folded onto +0
__device__ __forceinline__ uint32_t order_key(float f) {
  const uint32_t u = __float_as_uint(__fadd_rn(f, 0.0f));
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

}  // namespace b200det
