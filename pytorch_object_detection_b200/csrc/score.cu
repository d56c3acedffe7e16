// K1 — per-point detection score straight from the NCHW head outputs.
//
// Replaces reshape_cat_out + sigmoid + max + sqrt of FCOSHead.forward (model/modules/head.py:8-26,
// 53-64) without the NHWC copy the reference makes: a CTA owns 512 consecutive positions of one
// level of one image and walks the C class planes, which are contiguous along hw, so every warp
// load is a full 512-byte run.  sigmoid is monotone, so max_c sigmoid(x_c) = sigmoid(max_c x_c):
// only the running max of the logits is kept (strict '>' in ascending class order = torch.max's
// first-index rule) and ONE sigmoid per point is evaluated — two with the check that no other logit shares
// the winner's fp32 sigmoid (see upd).  fp16 / bf16 maps (autocast) are read as they are.  HBM-bound:
// (C+1)*4 bytes read and 6 bytes written per point.
#include "common.cuh"

namespace b200det {
namespace {


// CTA shape: threads x 4 points.  The kernel is ONE wave (config 2: 784 CTAs of 512 points = 5.3 per SM), but tile
// size is not what limits it: 128-, 256- and 512-point tiles (32 / 64 / 128 threads) all measure 21.4-21.7 us.
#ifndef B200DET_SCORE_THREADS
#define B200DET_SCORE_THREADS 128
#endif
#ifndef B200DET_SCORE_PIPELINE
#define B200DET_SCORE_PIPELINE 0      // measured: two alternating batches (loads of the next one issued before the current one
#endif                                // is reduced) are SLOWER — 25.5 us (2 x 4 planes), 27.1 us (2 x 8), 31.7 us (2 x 6) vs 21.7 us
#ifndef B200DET_SCORE_UNROLL
#define B200DET_SCORE_UNROLL 8
#endif
#ifndef B200DET_SCORE_MINBLOCKS
#define B200DET_SCORE_MINBLOCKS 4
#endif
constexpr int kUnroll = B200DET_SCORE_UNROLL;   // class planes per batch of independent 16-byte (fp32) / 8-byte (fp16, bf16) loads
constexpr int kScoreThreads = B200DET_SCORE_THREADS;
constexpr int kScoreTile = kScoreThreads * 4;
constexpr int kScoreCtasPerSm = B200DET_SCORE_MINBLOCKS * 128 / kScoreThreads;     // register budget

// number of kScoreTile tiles before level l (levels are few: a short loop per CTA)
__device__ __forceinline__ int score_level_of_tile(const LevelTable& lt, const int tile, int* first_tile) {
  int l = 0, first = 0, end = 0;                 // end = tiles of levels 0..i
#pragma unroll
  for (int i = 0; i < B200DET_MAX_LEVELS - 1; ++i) {
    end += (lt.hw[i] + kScoreTile - 1) / kScoreTile;
    if (i + 1 < lt.n_levels && tile >= end) {
      l = i + 1;
      first = end;
    }
  }
  *first_tile = first;
  return l;
}

// Running maximum of the LOGITS with torch.max's first index (strict '>' in ascending class order), plus the second
// largest value (equal values count): the epilogue needs it to tell whether ANOTHER logit shares the winner's fp32
// sigmoid, in which case the reference's argmax over sigmoid(cls) (head.py:57-62) may be an earlier class.
__device__ __forceinline__ void upd(float& best, float& second, int& arg, float v, int c) {
  second = fmaxf(second, fminf(v, best));
  if (v > best) {
    best = v;
    arg = c;
  }
}

template <typename T>
__global__ void __launch_bounds__(kScoreThreads, kScoreCtasPerSm)
score_points_kernel(const LevelTable lt, const int C, float* __restrict__ score, int16_t* __restrict__ cls0) {
  using E = MapElem<T>;
  const int b = blockIdx.y;
  int first_tile;
  const int l = score_level_of_tile(lt, blockIdx.x, &first_tile);
  const int hw = lt.hw[l];
  const int t0 = (blockIdx.x - first_tile) * kScoreTile;
  const void* __restrict__ cls = static_cast<const T*>(static_cast<const void*>(lt.cls[l])) + (size_t)b * C * hw;
  const void* __restrict__ cnt = static_cast<const T*>(static_cast<const void*>(lt.cnt[l])) + (size_t)b * hw;
  const size_t out0 = (size_t)b * lt.num_points + lt.point_off[l];

  float best[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  float second[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  int arg[4] = {0, 0, 0, 0};
  int pos[4];
  float cn[4];

  if (lt.vec_ok[l]) {
    const int p0 = t0 + threadIdx.x * 4;
    if (p0 >= hw) return;                    // hw % 4 == 0: a 4-group is all in or all out
#pragma unroll
    for (int q = 0; q < 4; ++q) pos[q] = p0 + q;
    // (B200DET_SCORE_PIPELINE: two alternating batches, the loads of the next one issued before the current one is
    // reduced — kept as a build option because it was measured and lost, see above)
    auto load = [&](float4* v, const int c0) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) v[u] = E::load4(cls, (size_t)(c0 + u) * hw + p0);
    };
    auto reduce = [&](const float4* v, const int c0) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        upd(best[0], second[0], arg[0], v[u].x, c0 + u);
        upd(best[1], second[1], arg[1], v[u].y, c0 + u);
        upd(best[2], second[2], arg[2], v[u].z, c0 + u);
        upd(best[3], second[3], arg[3], v[u].w, c0 + u);
      }
    };
    int c = 0;
#if B200DET_SCORE_PIPELINE
    float4 va[kUnroll], vb[kUnroll];
    if (C >= kUnroll) load(va, 0);
    while (c + kUnroll <= C) {
      const bool more_b = c + 2 * kUnroll <= C;
      if (more_b) load(vb, c + kUnroll);
      reduce(va, c);
      c += kUnroll;
      if (!more_b) break;
      const bool more_a = c + 2 * kUnroll <= C;
      if (more_a) load(va, c + kUnroll);
      reduce(vb, c);
      c += kUnroll;
      if (!more_a) break;
    }
#else
    for (; c + kUnroll <= C; c += kUnroll) {
      float4 v[kUnroll];
      load(v, c);
      reduce(v, c);
    }
#endif
    for (; c < C; ++c) {
      const float4 v = E::load4(cls, (size_t)c * hw + p0);
      upd(best[0], second[0], arg[0], v.x, c);
      upd(best[1], second[1], arg[1], v.y, c);
      upd(best[2], second[2], arg[2], v.z, c);
      upd(best[3], second[3], arg[3], v.w, c);
    }
    const float4 cv = E::load4(cnt, p0);
    cn[0] = cv.x; cn[1] = cv.y; cn[2] = cv.z; cn[3] = cv.w;
  } else {
    // planes not 16-byte aligned (e.g. 13x21 = 273): coalesced scalar loads, 4 strided points
    bool in[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      pos[q] = t0 + threadIdx.x + q * kScoreThreads;
      in[q] = pos[q] < hw;
    }
    if (!in[0]) return;
    int c = 0;
    for (; c + kUnroll <= C; c += kUnroll) {       // same batching as the vector path: loads first
      float v[kUnroll][4];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          v[u][q] = in[q] ? E::load1(cls, (size_t)(c + u) * hw + pos[q]) : -CUDART_INF_F;
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
#pragma unroll
        for (int q = 0; q < 4; ++q) upd(best[q], second[q], arg[q], v[u][q], c + u);
    }
    for (; c < C; ++c) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (in[q]) upd(best[q], second[q], arg[q], E::load1(cls, (size_t)c * hw + pos[q]), c);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      cn[q] = in[q] ? E::load1(cnt, pos[q]) : 0.f;
      if (!in[q]) pos[q] = -1;
    }
  }

#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (pos[q] < 0) continue;
    // head.py:57-63: sqrt(max_c sigmoid(cls) * sigmoid(cnt)), one rounding per operation.  sigmoid is monotone,
    // so max_c sigmoid(x_c) = sigmoid(max_c x_c); the class is torch.max's FIRST index among equal sigmoid values.
    const float s_best = sigmoid_f32(best[q]);
    int a = arg[q];
    if (second[q] != best[q] && sigmoid_f32(second[q]) == s_best) {
      // (rare: saturated logits, or logits a few ulps apart) another, different logit rounds to the same sigmoid:
      // walk the classes again for the first one that reaches it
      for (int c = 0; c < a; ++c)
        if (sigmoid_f32(E::load1(cls, (size_t)c * hw + pos[q])) == s_best) {
          a = c;
          break;
        }
    }
    score[out0 + pos[q]] = __fsqrt_rn(__fmul_rn(s_best, sigmoid_f32(cn[q])));
    cls0[out0 + pos[q]] = (int16_t)a;
  }
}

}  // namespace

int launch_score_points(const LevelTable& lt, int batch, int num_classes, float* score, int16_t* cls0,
                        cudaStream_t stream) {
  int tiles = 0;
  for (int l = 0; l < lt.n_levels; ++l) tiles += (lt.hw[l] + kScoreTile - 1) / kScoreTile;
  const dim3 grid(tiles, batch);
  if (lt.cls_dtype == B200DET_F16)
    score_points_kernel<__half><<<grid, kScoreThreads, 0, stream>>>(lt, num_classes, score, cls0);
  else if (lt.cls_dtype == B200DET_BF16)
    score_points_kernel<__nv_bfloat16><<<grid, kScoreThreads, 0, stream>>>(lt, num_classes, score, cls0);
  else
    score_points_kernel<float><<<grid, kScoreThreads, 0, stream>>>(lt, num_classes, score, cls0);
  return check_launch();
}

}  // namespace b200det

extern "C" int b200det_score_points(const b200det_level* levels, int n_levels, int batch, int num_classes,
                                    float* score, int16_t* cls0, void* stream) {
  using namespace b200det;
  LevelTable lt;
  if (!make_level_table(levels, n_levels, &lt) || batch <= 0 || batch > 65535 || num_classes <= 0 ||
      num_classes > 32767 || !score || !cls0)
    return B200DET_ERR_ARG;
  for (int l = 0; l < n_levels; ++l)
    if (!levels[l].cls || !levels[l].cnt) return B200DET_ERR_ARG;
  return launch_score_points(lt, batch, num_classes, score, cls0, static_cast<cudaStream_t>(stream));
}
