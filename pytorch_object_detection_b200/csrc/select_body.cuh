// K2 body shared by the stand-alone select kernel (select.cu) and the fused select+NMS cluster
// kernel (fused.cu).  See select.cu for the algorithm description.
#pragma once
#include "block_utils.cuh"
#include "nms.cuh"

namespace b200det {

constexpr int kSelThreads = 1024;
constexpr int kSelItems = 24;            // register-resident keys per thread: P <= 24576
constexpr int kFinishMax = 512;          // undecided keys handed to the single-warp finishing phase
static_assert(kSelThreads == kHistThreads, "the score histogram is scanned by 1024 threads");
__device__ __forceinline__ uint32_t load_key(const float* sc, int i, float thr) {
  const float s = sc[i];
  return (s >= thr) ? order_key(s) : 0u;
}

// Sum over the CTA with ONE barrier per call (double-buffered scratch, 2 x 32 ints).
__device__ __forceinline__ int block_sum(int v, int* buf, int& phase) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = __reduce_add_sync(0xffffffffu, v);
  int* bp = buf + (phase & 1) * 32;
  if (lane == 0) bp[warp] = v;
  __syncthreads();
  int r = (lane < (kSelThreads >> 5)) ? bp[lane] : 0;
  r = __reduce_add_sync(0xffffffffu, r);
  ++phase;
  return r;
}

#define FOR_KEYS(BODY)                                                  \
  if constexpr (REG) {                                                  \
    _Pragma("unroll") for (int j_ = 0; j_ < kSelItems; ++j_) {          \
      const uint32_t kx = key[j_];                                      \
      const int ix = tid + j_ * kSelThreads;                            \
      BODY                                                              \
    }                                                                   \
  } else {                                                              \
    for (int ix = tid; ix < P; ix += kSelThreads) {                     \
      const uint32_t kx = load_key(sc, ix, thr);                        \
      BODY                                                              \
    }                                                                   \
  }

// One CTA (kSelThreads threads) selects, sorts and decodes image `b`.  `sortbuf` is dynamic shared
// memory of at least max(next_pow2(k), 2 * kSelThreads) 64-bit words.
struct SelectResult {
  int count;          // candidates selected for the image (uniform over the CTA)
  float max_coord;    // max box coordinate over them (valid when out.nms_box != nullptr or want_max)
  // candidate row `threadIdx.x` as decoded by this thread (valid for threadIdx.x < count <= kSelThreads)
  float4 box;
  int cls;
  float score;
};

template <bool REG>
__device__ __forceinline__ SelectResult select_topk_cta(const LevelTable& lt, const float* __restrict__ score,
                                                const int16_t* __restrict__ cls0, const float thr, const int max_box,
                                                const CandSet& out, int32_t* __restrict__ cand_point, const int b,
                                                unsigned long long* sortbuf, const bool want_max = false,
                                                unsigned* hist = nullptr, const bool write_set = true) {
  // `write_set` false (the fused kernel): the candidate rows stay in the registers of the threads that decoded them
  // (SelectResult), nothing but count / mode goes to the global candidate set.
  // `hist` (REG path only; kHistBins words of shared memory, contents irrelevant on entry): replaces the ~17
  // block-wide bit passes of the threshold search by one pass of shared-memory atomics and one scan.
  __shared__ int s_red[64];
  __shared__ int s_scan[33];
  __shared__ uint32_t s_mm[64];
  __shared__ float s_fmax[32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = lt.num_points;
  const float* sc = score + (size_t)b * P;
  int phase = 0;

  B200DET_STAMP_NOSYNC(0);
  uint32_t key[REG ? kSelItems : 1];
  int nv = 0;
  uint32_t kmin = 0xffffffffu, kmax = 0u;
  if constexpr (REG) {
    float sv[kSelItems];                 // all loads issued before the first use: one memory round trip
#pragma unroll
    for (int j = 0; j < kSelItems; ++j) {
      const int i = tid + j * kSelThreads;
      sv[j] = (i < P) ? ldg_stream_f1(sc + i) : -CUDART_INF_F;
    }
    if (hist) {                          // (while the loads are in flight) an empty histogram
      uint4* h4 = reinterpret_cast<uint4*>(hist);
#pragma unroll
      for (int q = 0; q < kHistPerThread / 4; ++q) h4[tid + q * kSelThreads] = make_uint4(0u, 0u, 0u, 0u);
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < kSelItems; ++j) {
      const int i = tid + j * kSelThreads;
      const bool valid = i < P && sv[j] >= thr;
      key[j] = valid ? order_key(sv[j]) : 0u;
      // one pass over the keys builds them AND counts them (every pass over 24 keys x 1024 threads costs 0.7-2 us of
      // issue slots): bin = trunc(score * kHistBins), the same value hist_bin() recovers from the key
      if (hist && valid) atomicAdd(&hist[hist_slot(min(kHistBins - 1, max(0, (int)(sv[j] * (float)kHistBins))))], 1u);
    }
  }
  // ---- fast path: threshold from a histogram of the scores -------------------------------------------------
  bool have_T = false;
  bool radix_sorted = false;  // the selection already sits in sortbuf[kSelThreads, kSelThreads + cntT), sorted
  uint32_t T = 0u;
  int cntT = 0, n_valid = 0;
  if constexpr (REG) {
    if (hist) {
      __syncthreads();                              // every key has been counted
      B200DET_STAMP_NOSYNC(21);
      // thread t owns the bins [kHistBins - 16 (t + 1), kHistBins - 16 t): thread 0 the highest scores, so that an
      // exclusive scan over the threads counts the keys in HIGHER bins
      const int owner = kSelThreads - 1 - tid;                  // bins [16 owner, 16 owner + 16)
      const int lo = kHistPerThread * owner;
      unsigned hb[kHistPerThread];
#pragma unroll
      for (int q = 0; q < kHistPerThread; ++q) hb[q] = hist[q * kSelThreads + owner];
      int mine_h = 0;
#pragma unroll
      for (int q = 0; q < kHistPerThread; ++q) mine_h += (int)hb[q];
      const int higher = block_exclusive_scan(mine_h, s_scan, &n_valid);
      B200DET_STAMP_NOSYNC(22);
      const int kk_h = min(min(max_box, P), n_valid);
      if (kk_h > 0) {
        if (n_valid <= next_pow2(kk_h)) {          // everything valid fits the sort: no threshold needed
          T = 1u;                                   // (valid keys are non-zero)
          cntT = n_valid;
          have_T = true;
        } else {
          if (higher < kk_h && kk_h <= higher + mine_h) {   // the k-th score falls into one of this thread's bins
            int c = higher, bsel = lo;
#pragma unroll
            for (int q = kHistPerThread - 1; q >= 0; --q) {
              const int before = c;
              c += (int)hb[q];
              if (before < kk_h && kk_h <= c) {
                bsel = lo + q;
                s_mm[1] = (uint32_t)c;
              }
            }
            s_mm[0] = (uint32_t)bsel;
          }
          __syncthreads();
          const int bsel = (int)s_mm[0];
          cntT = (int)s_mm[1];
          __syncthreads();                          // s_mm is reused below
          if (cntT <= next_pow2(kk_h)) {
            // score >= bsel / kHistBins  <=>  bin >= bsel (the product is exact); bin 0 also holds scores <= 0
            T = bsel > 0 ? order_key((float)bsel / (float)kHistBins) : 1u;
            have_T = true;
          }
        }
        // The histogram also SORTS the selection (a one-pass radix sort on the bin): every bin's slot range in
        // descending bin order is known from the scan, selected keys take a slot of their bin with one shared-memory
        // atomic, and the few keys that share a bin are ranked against each other (key desc, point index asc).
        // Replaces the compaction pass and the 55-stage bitonic network when the selection fits one key per thread.
        if (have_T && cntT <= kSelThreads) {
          unsigned start = (unsigned)higher;
#pragma unroll
          for (int q = kHistPerThread - 1; q >= 0; --q) {       // counts -> first slot of the bin
            const unsigned c = hb[q];
            hist[q * kSelThreads + owner] = start;
            start += c;
          }
          __syncthreads();
          B200DET_STAMP_NOSYNC(23);
          unsigned long long* tmp = sortbuf;                    // unordered inside a bin
          unsigned long long* fin = sortbuf + kSelThreads;
          // (at most kSelThreads of the ~24 x kSelThreads keys are selected — one per thread on average — so a branch
          // per key beats keeping four atomics in flight: two instructions for a key that is not taken instead of ten)
#pragma unroll
          for (int j = 0; j < kSelItems; ++j) {
            const uint32_t k_ = key[j];
            if (k_ >= T) {                                      // (T >= 1 and the invalid keys are 0)
              const unsigned at = atomicAdd(&hist[hist_slot(hist_bin(k_))], 1u);
              tmp[at] = ((unsigned long long)k_ << 32) | (unsigned long long)(0xffffffffu - (uint32_t)(tid + j * kSelThreads));
            }
          }
          __syncthreads();
          B200DET_STAMP_NOSYNC(24);
          if (tid < cntT) {
            const unsigned long long e = tmp[tid];
            const int bin = hist_bin((uint32_t)(e >> 32));
            const unsigned first = bin == kHistBins - 1 ? 0u : hist[hist_slot(bin + 1)];   // = slots of all higher bins
            const unsigned end = hist[hist_slot(bin)];
            unsigned rank = 0;
            for (unsigned i = first; i < end; ++i) rank += tmp[i] > e ? 1u : 0u;
            fin[first + rank] = e;
          }
          radix_sorted = true;                                  // (the barrier before the decode publishes fin)
        }
      }
    }
  }
  if (!have_T) {
  FOR_KEYS(if (kx) { ++nv; kmin = min(kmin, kx); kmax = max(kmax, kx); })

  // block-wide n_valid / min / max
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  if (lane == 0) { s_mm[warp] = kmin; s_mm[32 + warp] = kmax; }
  n_valid = block_sum(nv, s_red, phase);             // barrier inside also publishes s_mm
  kmin = __reduce_min_sync(0xffffffffu, s_mm[lane]);
  kmax = __reduce_max_sync(0xffffffffu, s_mm[32 + lane]);
  }

  B200DET_STAMP(1);
  const int kk = min(min(max_box, P), n_valid);
  if (kk == 0) {
    if (tid == 0) { out.count[b] = 0; out.mode[b] = 0; }
    return SelectResult{0, 0.f, make_float4(0.f, 0.f, 0.f, 0.f), -1, 0.f};
  }

  // ---- k-th largest key -----------------------------------------------------------------
  const int n2 = next_pow2(kk);
  const int sort_cap = n2;    // the sort below handles up to n2 keys and orders ties by point index
  if (!have_T) {
    T = kmin;                 // count(key >= kmin) = n_valid
    cntT = n_valid;
  }
  const int list_cap = max(n2, 2 * kSelThreads);   // entries the dynamic shared buffer holds
  bool listed = false;        // selected entries already compacted into sortbuf[0, total)
  int total = 0;
  if (!have_T && n_valid > sort_cap) {
    const uint32_t diff = kmin ^ kmax;              // non-zero here unless all keys are equal
    const int top = diff ? 31 - __clz(diff) : -1;
    T = (top >= 0) ? (kmax & ~((2u << top) - 1u)) : kmax;
    // Phase A: block-wide passes over all keys.  `above` = count(key >= T + 2^(bit+1)) are already
    // known to be selected; the cntT - above keys inside [T, T + 2^(bit+1)) are still undecided.
    int above = 0, bit = top;
    bool done = false;
    // (k <= 2048: stop as soon as everything >= T fits the 2048-entry list of phase B; larger k: stop
    // when few keys are undecided and let the single-warp fallback finish.)
    const bool list_mode = kk <= 2 * kSelThreads;
    for (; bit >= 0 && (list_mode ? cntT > 2 * kSelThreads : cntT - above > kFinishMax); --bit) {
      const uint32_t cand = T | (1u << bit);
      int c = 0;
      FOR_KEYS(c += (kx >= cand) ? 1 : 0;)
      c = block_sum(c, s_red, phase);
      if (c >= kk) {
        T = cand;
        cntT = c;
        if (c <= sort_cap) { done = true; break; }  // everything >= T fits the sort: it finishes the selection
      } else {
        above = c;
      }
    }
    B200DET_STAMP(2);
    // Phase B (common case): few keys are still undecided and everything >= T fits the shared buffer.
    // ONE compaction pass over the registers moves those entries (key, ~index) to shared memory; the
    // remaining bits, the tie rule and the final selection then work on <= 2 entries per thread.
    if (!done && bit >= 0 && cntT <= min(list_cap, 2 * kSelThreads)) {
      int mine_l = 0;
      FOR_KEYS(mine_l += (kx >= T) ? 1 : 0;)
      int n_l;
      int at_l = block_exclusive_scan(mine_l, s_scan, &n_l);              // n_l == cntT
      FOR_KEYS(if (kx >= T) {
        sortbuf[at_l++] = ((unsigned long long)kx << 32) | (unsigned long long)(0xffffffffu - (uint32_t)ix);
      })
      __syncthreads();
      const unsigned long long eA = (tid < n_l) ? sortbuf[tid] : 0ull;
      const unsigned long long eB = (tid + kSelThreads < n_l) ? sortbuf[tid + kSelThreads] : 0ull;
      const uint32_t kA = (uint32_t)(eA >> 32), kB = (uint32_t)(eB >> 32);
      for (; bit >= 0; --bit) {
        const uint32_t cand = T | (1u << bit);
        const int c = block_sum(((kA >= cand) ? 1 : 0) + ((kB >= cand) ? 1 : 0), s_red, phase);
        if (c >= kk) {
          T = cand;
          cntT = c;
          if (c <= sort_cap) break;
        }
      }
      int idx_lim_l = 0x7fffffff;
      if (cntT > sort_cap) {          // every bit decided and too many ties on T: lowest point indices win
        const int iA = (int)(0xffffffffu - (uint32_t)eA), iB = (int)(0xffffffffu - (uint32_t)eB);
        const int c_gt = block_sum(((kA > T) ? 1 : 0) + ((kB > T) ? 1 : 0), s_red, phase);
        const int r = kk - c_gt;
        int L = 0;
        for (int ib = 31 - __clz(P); ib >= 0; --ib) {
          const int cand = L | (1 << ib);
          const int e = block_sum(((kA == T && iA < cand) ? 1 : 0) + ((kB == T && iB < cand) ? 1 : 0), s_red, phase);
          if (e <= r) L = cand;
        }
        idx_lim_l = L;
      }
      const int iA = (int)(0xffffffffu - (uint32_t)eA), iB = (int)(0xffffffffu - (uint32_t)eB);
      const bool sA = kA && (kA > T || (kA == T && iA < idx_lim_l));
      const bool sB = kB && (kB > T || (kB == T && iB < idx_lim_l));
      int at_s = block_exclusive_scan((sA ? 1 : 0) + (sB ? 1 : 0), s_scan, &total);   // its barriers order the reads above
      if (sA) sortbuf[at_s++] = eA;
      if (sB) sortbuf[at_s++] = eB;
      listed = true;
      done = true;
    }
    // Phase B (fallback when the list does not fit): the undecided keys go to shared memory and ONE
    // warp decides the remaining bits with warp reductions only.
    if (!done && bit >= 0) {
      const uint32_t span_hi = (bit >= 31) ? 0u : (T >> (bit + 1));       // keys sharing T's bits above `bit`
      int mine_u = 0;
      FOR_KEYS(mine_u += (kx >= T && ((bit >= 31) || (kx >> (bit + 1)) == span_hi)) ? 1 : 0;)
      int n_u;
      int at_u = block_exclusive_scan(mine_u, s_scan, &n_u);              // n_u == cntT - above <= kFinishMax
      uint32_t* ulist = reinterpret_cast<uint32_t*>(sortbuf);
      FOR_KEYS(if (kx >= T && ((bit >= 31) || (kx >> (bit + 1)) == span_hi)) ulist[at_u++] = kx;)
      __syncthreads();
      if (warp == 0) {
        uint32_t uk[kFinishMax / 32];
#pragma unroll
        for (int j = 0; j < kFinishMax / 32; ++j) uk[j] = (lane + 32 * j < n_u) ? ulist[lane + 32 * j] : 0u;
        for (; bit >= 0; --bit) {
          const uint32_t cand = T | (1u << bit);
          int c = 0;
#pragma unroll
          for (int j = 0; j < kFinishMax / 32; ++j) c += (uk[j] >= cand) ? 1 : 0;
          c = above + __reduce_add_sync(0xffffffffu, c);   // `above` stays fixed: every key outside the list
          if (c >= kk) {
            T = cand;
            cntT = c;
            if (c <= sort_cap) break;
          }
        }
        if (lane == 0) { s_mm[0] = T; s_mm[1] = (uint32_t)cntT; }
      }
      __syncthreads();
      T = s_mm[0];
      cntT = (int)s_mm[1];
      __syncthreads();                                                    // ulist (sortbuf) is reused below
    }
  }
  B200DET_STAMP(3);
  // cntT = count(key >= T) >= kk.  If it exceeds the sort capacity every bit was decided: T is
  // exactly the k-th key and more ties sit on it than fit -> keep the lowest point indices.
  int idx_lim = 0x7fffffff;
  if (!listed && cntT > sort_cap) {
    int c = 0;
    FOR_KEYS(c += (kx > T) ? 1 : 0;)
    const int c_gt = block_sum(c, s_red, phase);
    const int r = kk - c_gt;                        // how many == T to take, lowest index first
    int L = 0;                                      // largest L with count(key==T && idx<L) <= r
    for (int bit = 31 - __clz(P); bit >= 0; --bit) {
      const int cand = L | (1 << bit);
      int e = 0;
      FOR_KEYS(e += (kx == T && ix < cand) ? 1 : 0;)
      e = block_sum(e, s_red, phase);
      if (e <= r) L = cand;
    }
    idx_lim = L;
  }

  // ---- compaction (order irrelevant: sorted next) ---------------------------------------
  if (!listed && !radix_sorted) {
    int mine = 0;
    FOR_KEYS(mine += (kx > T || (kx == T && ix < idx_lim)) ? 1 : 0;)
    int at = block_exclusive_scan(mine, s_scan, &total);    // kk <= total <= n2
    FOR_KEYS(if (kx > T || (kx == T && ix < idx_lim)) {
      sortbuf[at++] = ((unsigned long long)kx << 32) | (unsigned long long)(0xffffffffu - (uint32_t)ix);
    })
  }
  B200DET_STAMP(4);
  const size_t o0 = (size_t)b * out.cap;
  float vmax = -CUDART_INF_F;

  float4 my_box = make_float4(0.f, 0.f, 0.f, 0.f);
  int my_cls = -1;
  float my_score = 0.f;
  // decode one selected point: box (head.py:29-38), class, score -> candidate row i
  auto emit = [&](const unsigned long long e, const int i) {
    const int p = (int)(0xffffffffu - (uint32_t)(e & 0xffffffffull));
    const int l = level_of_point(lt, p);
    const int pos = p - lt.point_off[l];
    const int hw = lt.hw[l], w = lt.w[l], s = lt.stride[l];
    const float x = (float)((pos % w) * s + s / 2);
    const float y = (float)((pos / w) * s + s / 2);
    const size_t r0 = (size_t)b * 4 * hw + pos;
    float4 d;
    if (lt.reg_dtype == B200DET_F32) {
      const float* rg = lt.reg[l] + r0;
      d = make_float4(rg[0], rg[hw], rg[2 * hw], rg[3 * hw]);
    } else {                                                  // fp16 / bf16 maps of an autocast forward, read as they are
      const void* rg = lt.reg[l];
      d = make_float4(load_map_elem(rg, lt.reg_dtype, r0), load_map_elem(rg, lt.reg_dtype, r0 + hw),
                      load_map_elem(rg, lt.reg_dtype, r0 + 2 * (size_t)hw), load_map_elem(rg, lt.reg_dtype, r0 + 3 * (size_t)hw));
    }
    if (lt.reg_scale[l]) {                                    // raw regression output: ScaleExp folded in
      const float sc = *lt.reg_scale[l];
      d = make_float4(scale_exp_f32(d.x, sc), scale_exp_f32(d.y, sc), scale_exp_f32(d.z, sc), scale_exp_f32(d.w, sc));
    }
    float4 bx;
    bx.x = __fsub_rn(x, d.x);
    bx.y = __fsub_rn(y, d.y);
    bx.z = __fadd_rn(x, d.z);
    bx.w = __fadd_rn(y, d.w);
    my_box = bx;
    my_cls = (int)cls0[(size_t)b * P + p] + 1;
    my_score = key_to_float((uint32_t)(e >> 32));
    if (write_set) {
      out.score[o0 + i] = my_score;
      out.cls[o0 + i] = my_cls;
      out.src[o0 + i] = i;
      if (cand_point) cand_point[o0 + i] = p;
      reinterpret_cast<float4*>(out.box)[o0 + i] = bx;
    }
    vmax = fmaxf(vmax, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
  };

  __syncthreads();
  if (radix_sorted) {
    B200DET_STAMP(5);
    if (tid < kk) emit(sortbuf[kSelThreads + tid], tid);
  } else if (n2 <= kSelThreads) {
    // one key per thread: shuffle network inside warps, shared memory only for distances >= 32
    unsigned long long v = (tid < total) ? sortbuf[tid] : 0ull;
    __syncthreads();
    if (n2 <= 64) v = bitonic_sort_desc_regs<64>(v, sortbuf);
    else if (n2 <= 256) v = bitonic_sort_desc_regs<256>(v, sortbuf);
    else v = bitonic_sort_desc_regs<1024>(v, sortbuf);
    B200DET_STAMP(5);
    if (tid < kk) emit(v, tid);
  } else {
    for (int i = total + tid; i < n2; i += kSelThreads) sortbuf[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc(sortbuf, n2);
    for (int i = tid; i < kk; i += kSelThreads) emit(sortbuf[i], i);
  }
  B200DET_STAMP(6);
  if (out.nms_box || want_max) vmax = block_max(vmax, s_fmax);
  if (out.nms_box) nms_prepare_boxes(out, b, kk, vmax, tid, kSelThreads);
  B200DET_STAMP(7);
  if (tid == 0) out.count[b] = kk;
  return SelectResult{kk, vmax, my_box, my_cls, my_score};
}


#undef FOR_KEYS

// dynamic shared memory (bytes) select_topk_cta needs for a capacity of k candidates
inline size_t select_sort_bytes(int k) {
  int n2 = 1;
  while (n2 < k) n2 <<= 1;
  return (size_t)(n2 > 2 * kSelThreads ? n2 : 2 * kSelThreads) * sizeof(unsigned long long);
}
// ... plus the score histogram of the register path behind the sort buffer
inline size_t select_smem_bytes(int k, bool with_hist) {
  return select_sort_bytes(k) + (with_hist ? (size_t)kHistBins * sizeof(unsigned) : 0);
}

}  // namespace b200det
